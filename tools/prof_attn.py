"""Not a test: drive the grouped mma.sync attention kernels (and the grouped LayerNorm) on the
shapes of BASELINE cfg 1a (9 chains, B=192, H=6, L=50, hd=16, 2 layers) and cfg 4 (18 problems,
B=256, H=8, L in {40,76,275}, hd=16, lite) for ncu / event timing.
    python tools/prof_attn.py [cfg1a|cfg4] [reps]
``cfg3c`` / ``cfg2`` time the UNGROUPED attention op on the text-encoder shape (B=128, H=8, L=256,
hd=64: tiled tcgen05 kernels; MMEMO_ATTN_TC2=0 routes it to the mma.sync kernels for an A/B) and on
the headline shape (B=64, H=8, L=128), three layers each (first / middle / last of a chain)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmemo_b200 import group_ops, ops  # noqa: E402

DEV = "cuda"
BF = torch.bfloat16
which = sys.argv[1] if len(sys.argv) > 1 else "cfg1a"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
g = torch.Generator().manual_seed(0)


def rnd(*s):
    return torch.randn(*s, generator=g).to(DEV).bfloat16()


if which in ("cfg3c", "cfg2"):
    B, H, d, L = (128, 8, 512, 256) if which == "cfg3c" else (64, 8, 512, 128)
    q, k, v, do = rnd(B, L, d), rnd(B, L, d), rnd(B, L, d), rnd(B, L, d)
    mask = (torch.arange(L)[None] < torch.randint(1, L + 1, (B, 1), generator=g)).float().to(DEV)
    c = torch.tensor([0.3], device=DEV)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def chain():
        """three layers: first (no prev, S out), middle (prev, S out), last (prev, recompute)"""
        times = []
        sp, saved = None, []
        for i in range(3):
            a, b_ = ev(), ev()
            a.record()
            o, s, st = ops._attn_fwd(True, q, k, v, mask, sp, c, H, i < 2)
            b_.record()
            times.append((f"fwd{i}", a, b_))
            saved.append((o, s, st, sp))
            sp = s
        dsn = None
        for i in (2, 1, 0):
            o, s, st, sp = saved[i]
            dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
            a, b_ = ev(), ev()
            a.record()
            dsn, _ = ops._attn_bwd(True, do, q, k, v, mask, s, sp, c, dsn, o, st, H, dq, dk, dv, True)
            b_.record()
            times.append((f"bwd{i}", a, b_))
        return times

    for _ in range(2):
        chain()
    acc = {}
    for _ in range(reps):
        ts = chain()
        torch.cuda.synchronize()
        for n, a, b_ in ts:
            acc.setdefault(n, []).append(a.elapsed_time(b_) * 1e3)
    print(which, "TC2 env", os.environ.get("MMEMO_ATTN_TC2", "on"),
          {n: round(sorted(v)[len(v) // 2], 1) for n, v in acc.items()}, "us")
    sys.exit(0)

if which == "cfg1a":
    B, H, d = 192, 6, 96
    shapes = [(50, 50)] * 9
    same_kv = False
else:
    B, H, d = 256, 8, 128
    Ls = (40, 76, 275)
    shapes = [(a, b) for a in Ls for b in Ls] * 2
    same_kv = True

qs = [rnd(B, lq, d) for lq, _ in shapes]
ks = [rnd(B, lk, d) for _, lk in shapes]
vs = ks if same_kv else [rnd(B, lk, d) for _, lk in shapes]
masks = [(torch.arange(lk)[None] < torch.randint(1, lk + 1, (B, 1), generator=g)).float().to(DEV)
         for _, lk in shapes]
cs = [torch.tensor([0.3], device=DEV) for _ in shapes]
none = [None] * len(shapes)


def run(layers=2 if which == "cfg1a" else 1):
    sp = none
    saved = []
    for i in range(layers):
        emit = i + 1 < layers
        o, s, st = group_ops.attn_fwd_group(True, qs, ks, vs, masks, sp, cs, H, emit, True)
        saved.append((o, s, st, sp))
        sp = s if emit else none
    dsn = none
    for i in reversed(range(layers)):
        o, s, st, sp = saved[i]
        dqs = [torch.empty_like(q) for q in qs]
        dks = [torch.empty_like(k) for k in ks]
        dvs = dks if same_kv else [torch.empty_like(k) for k in ks]
        dcs = [torch.zeros(1, device=DEV) for _ in shapes]
        dsp = group_ops.attn_bwd_group(True, o, qs, ks, vs, masks, s, sp, cs, dsn, o, st, H, dqs,
                                       dks, dvs, True, dcs, True)
        dsn = dsp


for _ in range(2):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    run()
e1.record()
torch.cuda.synchronize()
print(f"{which}: {e0.elapsed_time(e1) / reps * 1e3:.1f} us per fwd+bwd of all layers "
      f"(KB env {os.environ.get('MMEMO_ATTN_KB', 'auto')})")
