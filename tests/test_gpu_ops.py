"""GPU parity tests of the individual C-ABI ops (through the torch custom ops) against the CPU
oracle / plain torch fp32 on identical seeded inputs.

Tolerances (BASELINE.json north_star): float32 mode <= 1e-4 relative (max|a-b| / max|b|) on outputs
and gradients; bf16 mode <= 2e-2 relative on outputs.
"""
import math

import pytest
import torch
import torch.nn.functional as F

import mmemo_b200
from mmemo_b200 import ops
from oracle import mmemo_oracle as O
from tests.cases import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL32 = 1e-4
TOLBF = 2e-2


def g(seed):
    return torch.Generator().manual_seed(seed)


def rnd(gen, *shape, scale=1.0):
    return torch.randn(*shape, generator=gen) * scale


def lib_loaded():
    from mmemo_b200 import _lib
    return _lib.load().mmemo_version() >= 100


def test_library_loads():
    assert lib_loaded()


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(300, 96, 300), (9600, 96, 35), (77, 7, 576), (1, 12, 96),
                                   (513, 130, 67), (2048, 256, 128)])
@pytest.mark.parametrize("bias,pos,relu", [(False, False, False), (True, True, True)])
def test_linear_fp32(M, N, K, bias, pos, relu):
    gen = g(M + N + K)
    L = 25 if M % 25 == 0 else 1
    if pos and M % L:
        pytest.skip("no period")
    x = rnd(gen, M // L, L, K).requires_grad_(True)
    w = rnd(gen, N, K, scale=1 / math.sqrt(K)).requires_grad_(True)
    b = rnd(gen, N).requires_grad_(True) if bias else None
    p = rnd(gen, L, N).requires_grad_(True) if pos else None
    y = F.linear(x, w, b)
    if pos:
        y = y + p[None]
    if relu:
        y = torch.relu(y)
    dy = rnd(gen, *y.shape)
    y.backward(dy)
    xc, wc = x.detach().to(DEV).requires_grad_(True), w.detach().to(DEV).requires_grad_(True)
    bc = b.detach().to(DEV).requires_grad_(True) if bias else None
    pc = p.detach().to(DEV).requires_grad_(True) if pos else None
    yc = ops.linear(xc, wc, bc, pc, relu=relu)
    yc.backward(dy.to(DEV))
    assert rel_err(yc, y) < TOL32
    assert rel_err(xc.grad, x.grad) < TOL32
    assert rel_err(wc.grad, w.grad) < TOL32
    if bias:
        assert rel_err(bc.grad, b.grad) < TOL32
    if pos:
        assert rel_err(pc.grad, p.grad) < TOL32


@pytest.mark.parametrize("M,N,K", [(1024, 512, 512), (8192, 1024, 512), (640, 96, 304),
                                   (256, 128, 768)])
def test_linear_bf16(M, N, K):
    gen = g(7)
    x = rnd(gen, M, K).bfloat16()
    w = rnd(gen, N, K, scale=1 / math.sqrt(K))
    dy = rnd(gen, M, N).bfloat16()
    ref = x.float() @ w.bfloat16().float().t()
    xc = x.to(DEV).requires_grad_(True)
    wc = w.to(DEV).requires_grad_(True)
    ops.clear_shadow_cache()
    yc = ops.linear(xc, wc, bf16=True)
    yc.backward(dy.to(DEV))
    assert yc.dtype == torch.bfloat16 and wc.grad.dtype == torch.float32
    assert rel_err(yc.float(), ref) < TOLBF
    assert rel_err(xc.grad.float(), dy.float() @ w.bfloat16().float()) < TOLBF
    assert rel_err(wc.grad, dy.float().t() @ x.float()) < TOLBF


def test_split_k_workspace_is_per_stream():
    """The split-K scratch of the tcgen05 GEMM is an attribute of the stream (mmemo_stream_set_*):
    the same bf16-output split-K GEMM issued back to back on two concurrent streams must not mix
    partial tiles (it did when the workspace was one process-wide buffer)."""
    M, N, K = 512, 512, 4096        # 4 pair tiles, K long: splits into 16 slices through the scratch
    gen = g(11)
    xs = [rnd(gen, M, K).bfloat16().to(DEV) for _ in range(2)]
    ws = [rnd(gen, N, K, scale=1 / math.sqrt(K)).to(DEV) for _ in range(2)]
    refs = [x.float() @ w.bfloat16().float().t() for x, w in zip(xs, ws)]
    ops.clear_shadow_cache()
    with torch.no_grad():
        for x, w in zip(xs, ws):
            ops.linear(x, w, bf16=True)                         # shadows + default-stream warm-up
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = [[], []]
    with torch.no_grad():
        for it in range(20):
            for i, st in enumerate(streams):
                with torch.cuda.stream(st):
                    outs[i].append(ops.linear(xs[i], ws[i], bf16=True))
    torch.cuda.synchronize()
    keys = {k[1] for k in ops._stream_ws}
    assert all(st.cuda_stream in keys for st in streams)        # each stream got its own scratch
    bufs = {ops._stream_ws[k].data_ptr() for k in ops._stream_ws}
    assert len(bufs) == len(ops._stream_ws)
    for i in range(2):
        for y in outs[i]:
            assert rel_err(y.float(), refs[i]) < TOLBF


# ------------------------------------------------------------------------------------------------
ATTN_SHAPES = [  # B, H, Lq, Lk, hd
    (3, 6, 50, 50, 16), (2, 6, 20, 200, 16), (2, 8, 40, 275, 16), (2, 6, 25, 100, 32),
    (2, 8, 128, 128, 64), (1, 2, 7, 9, 12), (2, 2, 33, 65, 64), (1, 4, 256, 256, 64),
]


def _attn_inputs(B, H, Lq, Lk, hd, seed, prev):
    gen = g(seed)
    d = H * hd
    q, k, v = rnd(gen, B, Lq, d), rnd(gen, B, Lk, d), rnd(gen, B, Lk, d)
    lens = torch.randint(1, Lk + 1, (B,), generator=gen)
    mask = (torch.arange(Lk)[None] < lens[:, None]).float()
    sp = None
    if prev:  # a plausible previous-layer score tensor (already carries the mask term)
        sp = rnd(gen, B, H, Lq, Lk) - 1.0e8 * (1.0 - mask[:, None, None, :])
    c = torch.tensor([0.37])
    return q, k, v, mask, sp, c


@pytest.mark.parametrize("shape", ATTN_SHAPES)
@pytest.mark.parametrize("prev", [False, True])
def test_resattn_fp32(shape, prev):
    B, H, Lq, Lk, hd = shape
    q, k, v, mask, sp, c = _attn_inputs(*shape, seed=sum(shape), prev=prev)
    leaf = [t.clone().requires_grad_(True) for t in (q, k, v)]
    cl = c.clone().requires_grad_(True)
    spl = sp.clone().requires_grad_(True) if prev else None
    o, s = O.resattn_core(*leaf, mask, H, cl, spl)
    gen = g(5)
    do, ds = rnd(gen, *o.shape), rnd(gen, *s.shape) * 0.1
    ds = ds * mask[:, None, None, :]  # next layer's dS is exactly 0 on masked keys
    (o * do).sum().backward(retain_graph=True)
    (s * ds).sum().backward()
    dl = [t.to(DEV).requires_grad_(True) for t in (q, k, v)]
    cd = c.to(DEV).requires_grad_(True)
    spd = sp.to(DEV).requires_grad_(True) if prev else None
    od, sd, _ = ops.resattn_op(*dl, mask.to(DEV), spd, cd if prev else None, H)
    ((od * do.to(DEV)).sum() + (sd * ds.to(DEV)).sum()).backward()
    assert rel_err(od, o) < TOL32
    valid = mask[:, None, None, :].expand_as(s) > 0
    assert rel_err(sd.cpu()[valid], s[valid]) < TOL32
    assert torch.equal(sd.cpu()[~valid] < -5e7, torch.ones_like(s[~valid], dtype=torch.bool))
    for a, b in zip(dl, leaf):
        assert rel_err(a.grad, b.grad) < TOL32
    if prev:
        assert rel_err(spd.grad.cpu()[valid], spl.grad[valid]) < TOL32
        assert abs(cd.grad.item() - cl.grad.item()) < TOL32 * max(1.0, abs(cl.grad.item())) * 10


def test_resattn_fully_masked_rows_are_uniform_fp32():
    """SURVEY §8a note 1: mask all zero -> exactly uniform 1/Lk attention, no NaN."""
    B, H, Lq, Lk, hd = 2, 3, 9, 11, 16
    q, k, v, mask, _, _ = _attn_inputs(B, H, Lq, Lk, hd, seed=3, prev=False)
    mask[1] = 0
    o, s = O.resattn_core(q, k, v, mask, H)
    od, sd, _ = ops.resattn_op(q.to(DEV), k.to(DEV), v.to(DEV), mask.to(DEV), None, None, H)
    assert torch.isfinite(od).all()
    assert rel_err(od, o) < TOL32
    assert torch.equal(sd[1].cpu(), s[1])  # -1e8 exactly (qk absorbed by rounding)


def test_resattn_3d_mask_fp32():
    B, H, Lq, Lk, hd = 2, 2, 5, 6, 8
    q, k, v, _, _, _ = _attn_inputs(B, H, Lq, Lk, hd, seed=4, prev=False)
    m3 = (torch.rand(B, Lq, Lk, generator=g(1)) > 0.3).float()
    m3[..., 0] = 1
    o, s = O.resattn_core(q, k, v, m3, H)
    od, sd, _ = ops.resattn_op(q.to(DEV), k.to(DEV), v.to(DEV), m3.to(DEV), None, None, H)
    assert rel_err(od, o) < TOL32


# the shapes the reference's real models run in bf16 mode: cfg 1a (50,50,16), cfg 1b (20..200, 16),
# cfg 4 (40/76/275, 16; Lk > 128 = the tiled backward), cfg 5 (25/100, 32), cfg 2 (128,128,64),
# cfg 3 composite (256,256,64), plus ragged lengths
BF16_SHAPES = [(2, 6, 50, 50, 16), (2, 8, 128, 128, 64), (2, 8, 40, 275, 16), (2, 8, 275, 275, 16),
               (2, 6, 200, 100, 16), (2, 6, 20, 200, 16), (2, 8, 76, 40, 16), (2, 6, 25, 100, 32),
               (2, 6, 100, 100, 32), (2, 8, 256, 256, 64), (1, 3, 7, 9, 16), (1, 2, 33, 65, 64)]


@pytest.mark.parametrize("shape", BF16_SHAPES)
@pytest.mark.parametrize("prev", [False, True])
def test_resattn_bf16(shape, prev):
    B, H, Lq, Lk, hd = shape
    q, k, v, mask, sp, c = _attn_inputs(*shape, seed=1 + sum(shape), prev=prev)
    qb, kb, vb = (t.bfloat16() for t in (q, k, v))
    spb = sp.bfloat16() if prev else None
    o, s = O.resattn_core(qb.float(), kb.float(), vb.float(), mask, H, c, spb.float() if prev else None)
    od, sd, _ = ops.resattn_op(qb.to(DEV), kb.to(DEV), vb.to(DEV), mask.to(DEV),
                               spb.to(DEV) if prev else None, c.to(DEV) if prev else None, H)
    assert od.dtype == torch.bfloat16
    assert rel_err(od.float(), o) < TOLBF
    valid = mask[:, None, None, :].expand_as(s) > 0
    assert rel_err(sd.float().cpu()[valid], s[valid]) < TOLBF


@pytest.mark.parametrize("shape", BF16_SHAPES + [(3, 8, 128, 128, 64)])
@pytest.mark.parametrize("prev", [False, True])
def test_resattn_bf16_backward(shape, prev):
    """bf16 backward (tcgen05 kernel for hd=64/L=128, SIMT otherwise) against the fp32 oracle run
    on the same bf16-rounded inputs.  No stated criterion for bf16 gradients; 3e-2 relative."""
    B, H, Lq, Lk, hd = shape
    q, k, v, mask, sp, c = _attn_inputs(*shape, seed=2 + sum(shape), prev=prev)
    rb = lambda t: t.bfloat16().float()
    leaf = [rb(t).requires_grad_(True) for t in (q, k, v)]
    cl = c.clone().requires_grad_(True)
    spl = rb(sp).requires_grad_(True) if prev else None
    o, s = O.resattn_core(*leaf, mask, H, cl, spl)
    gen = g(6)
    do = rb(rnd(gen, *o.shape))
    ds = rb(rnd(gen, *s.shape) * 0.1 * mask[:, None, None, :])
    (o * do).sum().backward(retain_graph=True)
    (s * ds).sum().backward()
    dl = [t.detach().bfloat16().to(DEV).requires_grad_(True) for t in leaf]
    cd = c.to(DEV).requires_grad_(True)
    spd = sp.bfloat16().to(DEV).requires_grad_(True) if prev else None
    od, sd, _ = ops.resattn_op(*dl, mask.to(DEV), spd, cd if prev else None, H)
    ((od.float() * do.to(DEV)).sum() + (sd.float() * ds.to(DEV)).sum()).backward()
    for a, b_ in zip(dl, leaf):
        assert rel_err(a.grad.float(), b_.grad) < 3e-2
    if prev:
        valid = mask[:, None, None, :].expand_as(s) > 0
        assert rel_err(spd.grad.float().cpu()[valid], spl.grad[valid]) < 3e-2
        assert abs(cd.grad.item() - cl.grad.item()) < 3e-2 * max(1.0, abs(cl.grad.item()))


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,d", [(150, 96), (64, 512), (33, 128), (10, 16), (5, 192)])
@pytest.mark.parametrize("res,gate,relu", [(True, True, False), (False, False, True)])
def test_add_ln_fp32(M, d, res, gate, relu):
    gen = g(M + d)
    x = rnd(gen, M, d).requires_grad_(True)
    r = rnd(gen, M, d).requires_grad_(True) if res else None
    gt = torch.tensor([0.41]).requires_grad_(True) if gate else None
    w = (1 + 0.1 * rnd(gen, d)).requires_grad_(True)
    b = (0.1 * rnd(gen, d)).requires_grad_(True)
    z = (gt * x if gate else x)
    z = z + r if res else z
    y = F.layer_norm(z, (d,), w, b, 1e-5)
    y = torch.relu(y) if relu else y
    dy = rnd(gen, M, d)
    y.backward(dy)
    cu = lambda t: None if t is None else t.detach().to(DEV).requires_grad_(True)
    xc, rc, gc, wc, bc = cu(x), cu(r), cu(gt), cu(w), cu(b)
    yc = ops.add_ln(rc, xc, gc, wc, bc, relu=relu)
    yc.backward(dy.to(DEV))
    assert rel_err(yc, y) < TOL32
    assert rel_err(xc.grad, x.grad) < TOL32
    assert rel_err(wc.grad, w.grad) < TOL32 and rel_err(bc.grad, b.grad) < TOL32
    if res:
        assert rel_err(rc.grad, r.grad) < TOL32
    if gate:
        assert rel_err(gc.grad, gt.grad) < TOL32


def test_pool_matches_concat_mean_max():
    gen = g(9)
    B, d, n_slots, lens = 3, 24, 6, (5, 9, 7)   # groups in the order (l, a, v)
    segs = [[rnd(gen, B, L, d).requires_grad_(True) for _ in range(n_slots)] for L in lens]
    x = torch.cat([torch.cat(gs, 2) for gs in segs], 1)
    ref = torch.cat([x.mean(1), x.max(1)[0]], 1)
    dy = rnd(gen, *ref.shape)
    ref.backward(dy)
    csegs = [t.detach().to(DEV).requires_grad_(True) for gs in segs for t in gs]
    out = ops.pool(csegs, 3)
    out.backward(dy.to(DEV))
    assert rel_err(out, ref) < 1e-6
    flat = [t for gs in segs for t in gs]
    for a, b in zip(csegs, flat):
        assert rel_err(a.grad, b.grad) < 1e-6


def test_pool_ties_take_first_index():
    """Padded query rows of the lite models are identical -> ties in max (SURVEY §8a note 3)."""
    B, d = 2, 8
    a = torch.zeros(B, 4, d)
    b_ = torch.zeros(B, 3, d)
    c = torch.zeros(B, 2, d)
    segs = [t.to(DEV).requires_grad_(True) for t in (a, b_, c)]
    out = ops.pool(segs, 3)
    out[:, d:].sum().backward()   # only the max half
    assert torch.all(segs[0].grad[:, 0] == 1) and segs[0].grad[:, 1:].abs().sum() == 0
    assert segs[1].grad.abs().sum() == 0 and segs[2].grad.abs().sum() == 0


# ------------------------------------------------------------------------------------------------
def test_state_transfer_head():
    gen = g(21)
    B, P = 7, 6
    f = rnd(gen, B, P, 12).requires_grad_(True)
    T = torch.rand(6, 6, generator=gen).requires_grad_(True)
    ref = O.state_transfer_head(f, T)
    dy = rnd(gen, B, P, 6)
    ref.backward(dy)
    fc, Tc = f.detach().to(DEV).requires_grad_(True), T.detach().to(DEV).requires_grad_(True)
    out = ops.state_transfer_op(fc, Tc)
    out.backward(dy.to(DEV))
    assert rel_err(out, ref) < TOL32
    assert rel_err(fc.grad, f.grad) < TOL32 and rel_err(Tc.grad, T.grad) < TOL32


@pytest.mark.parametrize("B,C", [(5, 7), (64, 9), (1, 9)])
def test_bilinear_head(B, C):
    gen = g(B + C)
    leaves = [rnd(gen, B, C), rnd(gen, B, C), torch.rand(C, C, C, generator=gen),
              1 + 0.1 * rnd(gen, C), 0.1 * rnd(gen, C), rnd(gen, C, 2 * C) * 0.3, rnd(gen, C) * 0.1]
    cpu = [t.clone().requires_grad_(True) for t in leaves]
    ref = O.bilinear_head(*cpu)
    dy = rnd(gen, B, C)
    ref.backward(dy)
    cu = [t.clone().to(DEV).requires_grad_(True) for t in leaves]
    out = ops.bilinear_head(*cu)
    out.backward(dy.to(DEV))
    assert rel_err(out, ref) < TOL32
    for a, b in zip(cu, cpu):
        assert rel_err(a.grad, b.grad) < TOL32


def test_circle_loss_and_rdrop():
    gen = g(31)
    s = (rnd(gen, 6, 4, 9) * 3).requires_grad_(True)
    y = (torch.rand(6, 4, 9, generator=gen) < 0.3).long()
    y[0] = 0
    y[1] = 1
    ref = O.multi_circle_loss(s, y)
    w = torch.rand(6, 4, generator=gen)
    (ref * w).mean().backward()
    sc = s.detach().to(DEV).requires_grad_(True)
    out = ops.circle_loss_op(sc, y.to(DEV))
    (out * w.to(DEV)).mean().backward()
    assert rel_err(out, ref) < 1e-5 and rel_err(sc.grad, s.grad) < 1e-5

    lg = (rnd(gen, 8, 9) * 2).requires_grad_(True)
    ref = O.rdrop_kl(lg)
    ref.backward()
    lc = lg.detach().to(DEV).requires_grad_(True)
    out = ops.rdrop_kl_op(lc)
    out.backward()
    assert abs(out.item() - ref.item()) < 1e-5 * max(1, abs(ref.item()))
    assert rel_err(lc.grad, lg.grad) < TOL32


def test_dropout_mask_is_consistent_between_forward_and_backward():
    x = torch.ones(10000, device=DEV, requires_grad=True)
    y = ops.dropout_op(x, 0.25, 12345)
    y.sum().backward()
    keep = (y > 0).float().mean().item()
    assert abs(keep - 0.75) < 0.03
    assert torch.equal(x.grad, y.detach())      # same mask, same 1/(1-p) scale
    assert torch.allclose(y[y > 0], torch.tensor(1 / 0.75, device=DEV))


def test_unsupported_shape_fails_loudly():
    q = torch.zeros(1, 4, 2 * 200, device=DEV)
    with pytest.raises(RuntimeError):
        ops.resattn_op(q, q, q, None, None, None, 2)   # hd = 200 > 128


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(64, 128, 512), (3, 7, 5), (1,)])
def test_sq_mean_loss(shape, dt):
    """mean(x^2) and its gradient (bench.py's synthetic loss) vs torch."""
    x = rnd(g(9), *shape).to(dt)
    xr = x.float().clone().requires_grad_(True)
    ref = (xr ** 2).mean()
    (ref * 1.7).backward()
    xc = x.detach().clone().to(DEV).requires_grad_(True)
    out = ops.sq_mean_op(xc)
    (out * 1.7).backward()
    tol = 1e-5 if dt == torch.float32 else 1e-2
    assert abs(out.item() - ref.item()) <= tol * max(1.0, abs(ref.item()))
    assert xc.grad.dtype == dt
    assert rel_err(xc.grad.float(), xr.grad) < tol
