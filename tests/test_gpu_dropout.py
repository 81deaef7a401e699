"""Training dropout inside the grouped fusion trunk (Ren-MME/run.py:36 and robot_demo.py:40 train
with p = 0.1; their blocks drop the projected attention output and the block / FFN output).

Dropout is random, so parity with the reference is statistical + structural:
  * the grouped kernel keeps a fraction 1-p of the elements, scales them by 1/(1-p), draws the
    same mask for the same seed (forward == backward) and a new one when the seed or the device
    step counter changes;
  * one grouped layer (``group_ops.trunk_lite_op`` / ``trunk_full_op``, float32) with dropout equals
    the oracle's block with the SAME masks multiplied in at the reference's two dropout sites —
    outputs and every gradient (inputs, weights, biases, LayerNorm parameters, gates) to 1e-4:
    the backward applies exactly the forward's masks at the right places (before the bias column
    sums, before the weight gradients);
  * a whole model in train() mode runs its dropout on the grouped path, eval() switches it off."""
import itertools

import pytest
import torch

import mmemo_b200
from mmemo_b200 import group_ops, ops, synth
from oracle import mmemo_oracle as O
from tests.cases import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_dropout_multi_kernel(dt):
    g = torch.Generator().manual_seed(0)
    xs = [torch.randn(n, generator=g).abs().add(0.5).to(dt).to(DEV) for n in (100_003, 4096, 7)]
    p = 0.25
    seeds = ops.site_seeds(1234, len(xs), 0)
    ys = [torch.empty_like(x) for x in xs]
    ops._dropout_group(xs, ys, p, seeds)
    for x, y in zip(xs, ys):
        kept = y != 0
        if x.numel() > 1000:
            assert abs(kept.float().mean().item() - (1 - p)) < 0.01
        ref = (x.float() / (1 - p)).to(dt)
        assert torch.equal(y[kept], ref[kept])
    # same seeds -> same masks, in place too
    zs = [x.clone() for x in xs]
    ops._dropout_group(zs, zs, p, seeds)
    assert all(torch.equal(a, b) for a, b in zip(ys, zs))
    # other seeds / an advanced device step -> other masks
    ws = [torch.empty_like(x) for x in xs]
    ops._dropout_group(xs, ws, p, ops.site_seeds(1234, len(xs), 1))
    assert not torch.equal(ws[0] != 0, ys[0] != 0)
    ops.advance_dropout_step(DEV)
    try:
        ops._dropout_group(xs, ws, p, seeds)
        assert not torch.equal(ws[0] != 0, ys[0] != 0)
    finally:
        ops.dropout_step(DEV).zero_()


def _masks(shapes, p, seeds):
    """mask * 1/(1-p) of every problem of one dropout site, from the kernel itself."""
    ones = [torch.ones(sh, device=DEV) for sh in shapes]
    out = [torch.empty_like(t) for t in ones]
    ops._dropout_group(ones, out, p, seeds)
    return [t.cpu() for t in out]


def _problems(seed, d, shapes):
    g = torch.Generator().manual_seed(seed)
    qs, kvs, masks, sps = [], [], [], []
    for B, Lq, Lk, H in shapes:
        qs.append(torch.randn(B, Lq, d, generator=g))
        kvs.append(torch.randn(B, Lk, d, generator=g))
        lens = torch.randint(1, Lk + 1, (B,), generator=g)
        m = (torch.arange(Lk)[None] < lens[:, None]).float()
        masks.append(m)
        sps.append(torch.randn(B, H, Lq, Lk, generator=g) - 1.0e8 * (1.0 - m[:, None, None, :]))
    return qs, kvs, masks, sps


def _compare(outs_dev, outs_ref, leaves_dev, leaves_ref, weights):
    loss_d = sum((o.float() * w.to(DEV)).sum() for o, w in zip(outs_dev, weights))
    loss_r = sum((o * w).sum() for o, w in zip(outs_ref, weights))
    loss_d.backward()
    loss_r.backward()
    for o, r in zip(outs_dev, outs_ref):
        assert rel_err(o.detach().float(), r.detach()) < 1e-4
    for a, b in zip(leaves_dev, leaves_ref):
        assert a.grad is not None
        assert rel_err(a.grad.float(), b.grad) < 2e-4, (a.shape, rel_err(a.grad.float(), b.grad))


@pytest.mark.parametrize("p", [0.0, 0.3])
def test_lite_layer_with_dropout_matches_oracle_with_the_same_masks(p):
    d, H, seed = 32, 2, 4242
    shapes = [(2, 6, 9, H), (3, 7, 7, H), (2, 9, 6, H)]
    G = len(shapes)
    qs, kvs, masks, sps = _problems(1, d, shapes)
    g = torch.Generator().manual_seed(2)
    params = []
    for _ in range(G):     # proj.weight, minus.weight, norm.weight, norm.bias, c
        params += [torch.randn(d, d, generator=g) * 0.2, torch.randn(d, 2 * d, generator=g) * 0.2,
                   1 + 0.1 * torch.randn(d, generator=g), 0.1 * torch.randn(d, generator=g),
                   torch.tensor([0.3])]
    m0 = _masks([q.shape for q in qs], p, ops.site_seeds(seed, G, 0)) if p else None
    m1 = _masks([q.shape for q in qs], p, ops.site_seeds(seed, G, 1)) if p else None
    # ---- oracle with explicit masks (Ren-MME/run.py:188-214) -----------------------------------------
    rq, rkv, rsp, rp = ([t.clone().requires_grad_(True) for t in ts] for ts in (qs, kvs, sps, params))
    outs_ref = []
    for i in range(G):
        wo, wm, nw, nb, c = rp[5 * i:5 * i + 5]
        o, _ = O.resattn_core(rq[i], rkv[i], rkv[i], masks[i], H, c, rsp[i])
        x = o @ wo.t()
        x = x * m0[i] if p else x
        out = O._ln(torch.cat([rq[i], x], -1) @ wm.t(), nw, nb)
        outs_ref.append(out * m1[i] if p else out)
    # ---- device -----------------------------------------------------------------------------------
    dq, dkv, dsp, dp = ([t.clone().to(DEV).requires_grad_(True) for t in ts]
                        for ts in (qs, kvs, sps, params))
    res = group_ops.trunk_lite_op(dq, dkv, [m.to(DEV) for m in masks], dsp, dp, H, False, False,
                                  p, seed)
    outs_dev = [res[group_ops.LITE_OUT * i] for i in range(G)]
    weights = [torch.randn(o.shape, generator=g) for o in outs_ref]
    _compare(outs_dev, outs_ref, dq + dkv + dsp + dp, rq + rkv + rsp + rp, weights)


@pytest.mark.parametrize("p", [0.0, 0.3])
def test_full_layer_with_dropout_matches_oracle_with_the_same_masks(p):
    d, H, dff, seed = 24, 2, 48, 777
    shapes = [(2, 5, 8, H), (2, 8, 5, H)]
    G = len(shapes)
    qs, kvs, masks, sps = _problems(3, d, shapes)
    g = torch.Generator().manual_seed(4)
    params = []
    for _ in range(G):   # wq wk wv wo n1w n1b n2w n2b f1w f1b f2w f2b a b c
        r = lambda *sh: torch.randn(*sh, generator=g) * 0.2
        params += [r(d, d), r(d, d), r(d, d), r(d, d), 1 + r(d) * 0.5, r(d) * 0.5, 1 + r(d) * 0.5,
                   r(d) * 0.5, r(dff, d), r(dff), r(d, dff), r(d), torch.tensor([0.7]),
                   torch.tensor([0.6]), torch.tensor([0.3])]
    m0 = _masks([q.shape for q in qs], p, ops.site_seeds(seed, G, 0)) if p else None
    m1 = _masks([q.shape for q in qs], p, ops.site_seeds(seed, G, 1)) if p else None
    # ---- oracle with explicit masks (others/realformer.py:163-209, robot_demo.py:333-374) ------------
    rq, rkv, rsp, rp = ([t.clone().requires_grad_(True) for t in ts] for ts in (qs, kvs, sps, params))
    outs_ref = []
    for i in range(G):
        wq, wk, wv, wo, n1w, n1b, n2w, n2b, f1w, f1b, f2w, f2b, a, b, c = rp[15 * i:15 * i + 15]
        o, _ = O.resattn_core(rq[i] @ wq.t(), rkv[i] @ wk.t(), rkv[i] @ wv.t(), masks[i], H, c, rsp[i])
        x = o @ wo.t()
        x = x * m0[i] if p else x
        h1 = O._ln(rq[i] + a * x, n1w, n1b)
        f = torch.relu(h1 @ f1w.t() + f1b) @ f2w.t() + f2b
        f = f * m1[i] if p else f
        outs_ref.append(O._ln(h1 + b * f, n2w, n2b))
    dq, dkv, dsp, dp = ([t.clone().to(DEV).requires_grad_(True) for t in ts]
                        for ts in (qs, kvs, sps, params))
    res = group_ops.trunk_full_op(dq, dkv, [m.to(DEV) for m in masks], dsp, dp, H, False, False,
                                  p, seed)
    outs_dev = [res[group_ops.FULL_OUT * i] for i in range(G)]
    weights = [torch.randn(o.shape, generator=g) for o in outs_ref]
    _compare(outs_dev, outs_ref, dq + dkv + dsp + dp, rq + rkv + rsp + rp, weights)


def test_train_mode_runs_dropout_on_the_grouped_path_and_eval_switches_it_off(monkeypatch):
    monkeypatch.setattr(mmemo_b200.ren_mme, "DROP", 0.3)
    torch.manual_seed(0)
    m = mmemo_b200.ren_mme.Base_model(32, 6, 7, 9, 2, 2, 1, l_dim=12, v_dim=10, a_dim=8)
    m.load_state_dict(synth.randomize_gates({k: v.detach().clone() for k, v in m.state_dict().items()}))
    m = m.to(DEV).train()
    assert m.intensity.multimodal_blocks[0].drop.p == 0.3
    b = synth.renmme_batch(seed=3, B=4, L=(6, 7, 9), D=(12, 10, 8), rdrop_pairs=True)
    inputs = [t.to(DEV) for t in b["inputs"]]
    calls = []
    real = ops._dropout_group
    monkeypatch.setattr(ops, "_dropout_group",
                        lambda xs, ys, p, seeds: (calls.append(p), real(xs, ys, p, seeds))[1])

    def run(pinned=True):
        if pinned:
            counter = itertools.count(1)
            monkeypatch.setattr(ops, "next_dropout_seed", lambda: 1000 + next(counter))
        return m(*inputs)

    for mode in ("fp32", "bf16"):
        with mmemo_b200.precision(mode):
            a, b_ = run(), run()
            assert torch.allclose(a, b_, atol=1e-5 if mode == "fp32" else 1e-2)   # same seeds, same masks
            assert 0.3 in calls
            calls.clear()
            # R-Drop (Ren-MME/run.py:143-146, 332-334): rows 2i / 2i+1 carry the same sample and
            # must see DIFFERENT masks, otherwise the KL term is identically zero
            assert (a[0::2] - a[1::2]).abs().max() > 1e-3
            ops.rdrop_kl_op(a).backward()
            assert m.intensity.multimodal_blocks[0].proj.weight.grad is not None
            m.eval()
            calls.clear()
            c = run()
            assert 0.3 not in calls
            m.train()
            assert (c[0::2] - c[1::2]).abs().max() < (1e-5 if mode == "fp32" else 1e-2)   # no dropout
