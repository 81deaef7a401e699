"""The warp-level tensor-core attention kernels (csrc/resattn_mma.cu) through the GROUPED C-ABI
entry points, against the CPU oracle (fp32 math on the same bf16-rounded inputs).

Covers what the fused trunk uses and the ungrouped op tests cannot reach: several problems of
different shapes in one launch, scores with a padded row stride, S stored vs recomputed in the
backward, K = V aliasing (lite blocks), dS_next / S_prev / dc, ragged lengths."""
import pytest
import torch

from mmemo_b200 import ops
from oracle import mmemo_oracle as O
from tests.cases import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"
BF = torch.bfloat16


def rb(t):
    return t.bfloat16().float()


def _problem(seed, B, H, Lq, Lk, hd, prev, same_kv):
    g = torch.Generator().manual_seed(seed)
    d = H * hd
    q, k = rb(torch.randn(B, Lq, d, generator=g)), rb(torch.randn(B, Lk, d, generator=g))
    v = k if same_kv else rb(torch.randn(B, Lk, d, generator=g))
    lens = torch.randint(1, Lk + 1, (B,), generator=g)
    mask = (torch.arange(Lk)[None] < lens[:, None]).float()
    sp = None
    if prev:
        sp = rb(torch.randn(B, H, Lq, Lk, generator=g) - 1.0e8 * (1.0 - mask[:, None, None, :]))
    c = torch.tensor([0.37])
    do = rb(torch.randn(B, Lq, d, generator=g))
    dsn = rb(torch.randn(B, H, Lq, Lk, generator=g) * 0.1 * mask[:, None, None, :])
    return dict(q=q, k=k, v=v, mask=mask, sp=sp, c=c, do=do, dsn=dsn, B=B, H=H, Lq=Lq, Lk=Lk, hd=hd,
                same_kv=same_kv)


def _oracle(p, use_dsn):
    q = p["q"].clone().requires_grad_(True)
    k = p["k"].clone().requires_grad_(True)
    v = k if p["same_kv"] else p["v"].clone().requires_grad_(True)
    c = p["c"].clone().requires_grad_(True)
    sp = p["sp"].clone().requires_grad_(True) if p["sp"] is not None else None
    o, s = O.resattn_core(q, k, v, p["mask"], p["H"], c, sp)
    loss = (o * p["do"]).sum()
    if use_dsn:
        loss = loss + (s * p["dsn"]).sum()
    loss.backward()
    return dict(o=o.detach(), s=s.detach(), dq=q.grad, dk=k.grad, dv=None if p["same_kv"] else v.grad,
                dsp=None if sp is None else sp.grad, dc=None if sp is None else c.grad)


def _padded(t, lds):
    """(B,H,Lq,Lk) -> device bf16 tensor with row stride lds (returns the padded buffer)."""
    B, H, Lq, Lk = t.shape
    buf = torch.zeros(B, H, Lq, lds, dtype=BF, device=DEV)
    buf[..., :Lk] = t.to(DEV).bfloat16()
    return buf


SETS = [
    # cfg 1a: nine equal chains -> here three, full blocks (K != V), two layers (S written / read)
    [(2, 6, 50, 50, 16, False), (2, 6, 50, 50, 16, False), (3, 6, 50, 50, 16, False)],
    # cfg 1b / 4: lite blocks (K == V), every (Lq, Lk) combination differs
    [(2, 8, 40, 275, 16, True), (2, 8, 275, 40, 16, True), (1, 8, 275, 275, 16, True),
     (2, 8, 76, 76, 16, True)],
    # cfg 5 (hd 32) and ragged lengths
    [(2, 6, 25, 100, 32, False), (2, 6, 100, 25, 32, False), (1, 6, 100, 100, 32, False)],
    [(1, 3, 7, 9, 16, False), (2, 2, 33, 65, 64, False), (1, 2, 130, 70, 64, True)],
    # cfg 3 composite shape (hd 64, L 256): K / V staged one key block at a time in the backward;
    # the small problem rides along in the same (blocked) launch
    [(1, 4, 256, 256, 64, False), (1, 2, 40, 72, 64, False)],
]


@pytest.mark.parametrize("si", range(len(SETS)))
@pytest.mark.parametrize("prev", [False, True])
@pytest.mark.parametrize("stored", [False, True])
def test_grouped_attention_matches_oracle(si, prev, stored):
    shapes = SETS[si]
    hds = {s[4] for s in shapes}
    groups = [[s for s in shapes if s[4] == hd] for hd in sorted(hds)]   # one hd per launch
    for grp in groups:
        ps = [_problem(100 * si + i, *s[:5], prev, s[5]) for i, s in enumerate(grp)]
        refs = [_oracle(p, use_dsn=stored) for p in ps]
        dev, fw, bw = [], [], []
        for p in ps:
            B, H, Lq, Lk, hd = p["B"], p["H"], p["Lq"], p["Lk"], p["hd"]
            lds = ops.score_stride(Lk)
            t = dict(q=p["q"].to(DEV).bfloat16(), k=p["k"].to(DEV).bfloat16())
            t["v"] = t["k"] if p["same_kv"] else p["v"].to(DEV).bfloat16()
            t["mask"] = p["mask"].to(DEV)
            t["sp"] = _padded(p["sp"], lds) if prev else None
            t["c"] = p["c"].to(DEV)
            t["s"] = torch.full((B, H, Lq, lds), 7.0, dtype=BF, device=DEV)
            t["o"] = torch.empty(B, Lq, H * hd, dtype=BF, device=DEV)
            t["stat"] = torch.empty(B, H, Lq, 2, device=DEV)
            t["do"] = p["do"].to(DEV).bfloat16()
            t["dsn"] = _padded(p["dsn"], lds) if stored else None
            t["dq"], t["dk"] = torch.empty_like(t["q"]), torch.empty_like(t["k"])
            t["dv"] = torch.empty_like(t["k"])
            t["dsp"] = torch.zeros(B, H, Lq, lds, dtype=BF, device=DEV) if prev else None
            t["dc"] = torch.zeros(1, device=DEV) if prev else None
            dev.append(t)
            fw.append(ops._attn_problem(t["q"], t["k"], t["v"], t["mask"], t["sp"], t["c"], t["s"],
                                        t["o"], t["stat"], H, lds))
            bw.append(ops._attn_problem(t["q"], t["k"], t["v"], t["mask"], t["sp"], t["c"], None,
                                        t["o"], t["stat"], H, lds, d_o=t["do"],
                                        s=t["s"] if stored else None, ds_next=t["dsn"], dq=t["dq"],
                                        dk=t["dk"], dv=t["dv"], ds_prev=t["dsp"], dc=t["dc"]))
        assert ops._attn_group_call("mmemo_resattn_fwd_grouped_bf16", fw)
        assert ops._attn_group_call("mmemo_resattn_bwd_grouped_bf16", bw)
        torch.cuda.synchronize()
        for p, r, t in zip(ps, refs, dev):
            Lk = p["Lk"]
            tag = (p["Lq"], Lk, p["hd"])
            assert rel_err(t["o"].float(), r["o"]) < 2e-2, tag
            valid = p["mask"][:, None, None, :].expand_as(r["s"]) > 0
            s_dev = t["s"][..., :Lk].float().cpu()
            assert rel_err(s_dev[valid], r["s"][valid]) < 2e-2, tag
            assert bool((s_dev[~valid] < -5e7).all()), tag
            assert rel_err(t["dq"].float(), r["dq"]) < 3e-2, tag
            if p["same_kv"]:
                assert rel_err(t["dk"].float() + t["dv"].float(), r["dk"]) < 3e-2, tag
            else:
                assert rel_err(t["dk"].float(), r["dk"]) < 3e-2, tag
                assert rel_err(t["dv"].float(), r["dv"]) < 3e-2, tag
            if prev:
                dsp = t["dsp"][..., :Lk].float().cpu()
                assert rel_err(dsp[valid], r["dsp"][valid]) < 3e-2, tag
                # dc = sum(dS * S_prev) is ONE cancelling sum over B*H*Lq*Lk signed bf16-noisy terms:
                # its absolute error grows like sqrt(n) (0.05 at n = 6e5), not with |dc|
                n_terms = p["B"] * p["H"] * p["Lq"] * Lk
                tol = 3e-2 * max(1.0, abs(r["dc"].item()), n_terms ** 0.5 / 250.0)
                assert abs(t["dc"].item() - r["dc"].item()) < tol, tag


def test_fully_masked_rows_are_uniform_on_the_mma_path():
    """mask all zero -> exactly uniform 1/Lk attention, no NaN.  In bf16 mode EVERY score of such a
    row is bf16(-1e8) (the ulp there is 2^19, far above any QK^T), so the output is the plain mean
    of V - what the reference's own .bfloat16() run computes.  (The fp32 oracle is not the yardstick
    for these rows: in fp32 a |QK^T/sqrt(hd)| >= 4 survives the subtraction quantised to 8.)"""
    p = _problem(5, 2, 6, 50, 50, 16, False, False)
    p["mask"][1] = 0
    r = _oracle(p, use_dsn=False)
    o, s, _ = ops.resattn_op(p["q"].to(DEV).bfloat16(), p["k"].to(DEV).bfloat16(),
                             p["v"].to(DEV).bfloat16(), p["mask"].to(DEV), None, None, 6)
    assert torch.isfinite(o.float()).all()
    assert rel_err(o[0].float(), r["o"][0]) < 2e-2                 # the normally masked sample
    expect = p["v"][1].mean(0, keepdim=True).expand(50, -1)        # uniform attention = mean of V
    assert rel_err(o[1].float(), expect) < 2e-2
    assert bool((s[1].float() == torch.tensor(-1.0e8).bfloat16().float()).all())


def test_mma_path_is_the_one_that_runs():
    from mmemo_b200 import _lib
    lib = _lib.load()
    for Lq, Lk, hd in ((50, 50, 16), (275, 275, 16), (100, 25, 32), (20, 200, 16)):
        assert lib.mmemo_resattn_uses_mma(Lq, Lk, hd, 8 * hd, 0, 0) == 1
        assert lib.mmemo_resattn_uses_mma(Lq, Lk, hd, 8 * hd, 1, 1) == 1
