"""CPU tests that pin the oracle (oracle/mmemo_oracle.py).

1. against the committed golden fixtures (outputs of the reference's own classes, frozen by
   tests/golden/make_golden.py) — runs everywhere;
2. against the live reference classes extracted from /root/reference — runs only where that tree
   exists (the build container), including extra behaviours the fixtures do not hold: all-zero
   masks (uniform attention, SURVEY §8a note 1), fp64, and the loss functions on edge labels.
Tolerance: fp32 reference-vs-restatement noise; the reference's own fp32-vs-fp64 floor is 3e-7 on
logits / 7e-6 on the worst gradient tensor, so 2e-5 relative leaves margin.
"""
import numpy as np
import pytest
import torch

from oracle import mmemo_oracle as O
from oracle import refload
from tests import cases

TOL = 2e-5
needs_ref = pytest.mark.skipif(not refload.available(), reason="/root/reference not present")


@pytest.mark.parametrize("name", list(cases.CASES))
def test_oracle_matches_golden(name):
    c = cases.CASES[name]
    g = torch.load(cases.golden_path(name))
    logits, loss, grads, igrads = cases.run_with_grads(c.oracle, g["state"], g["batch"], c.loss, O,
                                                       c.grad_inputs)
    assert cases.rel_err(logits, g["logits"]) < TOL
    assert abs(loss.item() - g["loss"].item()) < TOL * max(1.0, abs(g["loss"].item()))
    # every parameter the reference gives a gradient to must get (the same) one from the oracle
    assert set(grads) == set(g["grads"]), set(grads) ^ set(g["grads"])
    for k, v in g["grads"].items():
        assert cases.rel_err(grads[k], v) < TOL * 5, k
    for k, v in g["input_grads"].items():
        assert cases.rel_err(igrads[k], v) < TOL * 5, k


@needs_ref
@pytest.mark.parametrize("name", list(cases.CASES))
def test_golden_is_current_reference_output(name):
    """The fixture really is what the reference computes here (guards against a stale file)."""
    c = cases.CASES[name]
    g = torch.load(cases.golden_path(name))
    ns = refload.load(c.family, **c.overrides)
    model = c.ref_model(ns).float().train()
    model.load_state_dict(g["state"])
    with torch.no_grad():
        out = c.call(model, g["batch"])
    assert cases.rel_err(out, g["logits"]) < 1e-6


@needs_ref
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_block_full_vs_reference_allzero_mask_and_residual(dtype):
    """Full block, two layers deep (live score residual), with one all-zero mask row (forward
    only: SURVEY §8a note 1 — uniform attention in fp32, softmax(qk) in fp64)."""
    ns = refload.load("realformer", DROP=0.0, FFN=2)
    torch.manual_seed(3)
    blks = [ns.Attention_Block(24, 3).to(dtype) for _ in range(2)]
    sd = {}
    for i, b in enumerate(blks):
        st = cases.seeded_state(b, seed=5 + i)
        b.load_state_dict(st)
        sd.update({f"b{i}.{k}": v for k, v in st.items()})
    g = torch.Generator().manual_seed(4)
    q = torch.randn(3, 7, 24, generator=g).to(dtype)
    kv = torch.randn(3, 9, 24, generator=g).to(dtype)
    mask = (torch.arange(9)[None] < torch.tensor([9, 4, 0])[:, None]).to(dtype)
    with torch.no_grad():
        r1, s1 = blks[0](q, kv, kv, mask, None)
        r2, s2 = blks[1](r1, kv, kv, mask, s1)
        o1, t1 = O.block_full(sd, "b0.", q, kv, kv, mask, 3, None)
        o2, t2 = O.block_full(sd, "b1.", o1, kv, kv, mask, 3, t1)
    tol = 1e-5 if dtype == torch.float32 else 1e-12
    assert cases.rel_err(o2, r2) < tol
    assert cases.rel_err(t2, s2) < tol
    if dtype == torch.float32:  # fully masked row -> exactly uniform attention
        att = torch.softmax(t1[2], -1)
        assert torch.allclose(att, torch.full_like(att, 1 / 9), atol=1e-7)


@needs_ref
def test_block_lite_vs_reference_3d_mask():
    """Lite block with the (B, Lq, Lk) mask branch (others/realformer.py:197-199 twin)."""
    ns = refload.load("renmme", DROP=0.0)
    torch.manual_seed(5)
    blk = ns.Attention_Block(16, 2, 1)
    st = cases.seeded_state(blk, seed=2)
    blk.load_state_dict(st)
    g = torch.Generator().manual_seed(6)
    q, kv = torch.randn(2, 5, 16, generator=g), torch.randn(2, 6, 16, generator=g)
    m3 = (torch.rand(2, 5, 6, generator=g) > 0.3).float()
    m3[..., 0] = 1
    with torch.no_grad():
        r, s = blk(q, kv, kv, m3, None)
        o, t = O.block_lite(st, "", q, kv, kv, m3, 2, None, norm="norm2")
    assert cases.rel_err(o, r) < 1e-5 and cases.rel_err(t, s) < 1e-5


@needs_ref
def test_losses_vs_reference():
    ns_r = refload.load("realformer")
    ns_m = refload.load("renmme")
    g = torch.Generator().manual_seed(7)
    s = torch.randn(6, 4, 9, generator=g) * 3
    y = (torch.rand(6, 4, 9, generator=g) < 0.3).long()
    y[0] = 0          # no positive label
    y[1] = 1          # all labels positive
    assert torch.allclose(O.multi_circle_loss(s, y), ns_r.multi_circle_loss(s, y), atol=1e-6)
    s2, y2 = s[:, 0], y[:, 0].float()
    assert torch.allclose(O.multi_loss(s2, y2), ns_m.multi_loss(s2, y2), atol=1e-6)


def test_state_transfer_head_single_window_is_identity():
    f = torch.randn(4, 1, 12)
    out = O.state_transfer_head(f, torch.rand(6, 6))
    assert torch.equal(out[:, 0], f[:, 0, :6])


def test_circle_loss_closed_form():
    """loss = log(1 + sum_{y=0} e^s) + log(1 + sum_{y=1} e^-s)."""
    s = torch.tensor([[0.5, -1.0, 2.0]])
    y = torch.tensor([[1, 0, 0]])
    want = torch.log1p(torch.exp(s[0, 1]) + torch.exp(s[0, 2])) + torch.log1p(torch.exp(-s[0, 0]))
    assert torch.allclose(O.multi_circle_loss(s, y)[0], want, atol=1e-6)


# ---- batch assembly: oracle/batching_oracle.py vs the reference's own helpers -----
def _ragged(seed, D, lens):
    rng = np.random.default_rng(seed)
    seqs = []
    for T in lens:
        a = rng.standard_normal((T, D)).astype(np.float32)
        if T:
            bad = rng.random((T, D)) < 0.02
            a[bad] = rng.choice(np.array([np.nan, np.inf, -np.inf], dtype=np.float32), bad.sum())
        seqs.append(a)
    return seqs


@needs_ref
def test_batching_oracle_tail_matches_reference_masking():
    from oracle import batching_oracle as BO
    ns = refload.load("realformer")
    for a in _ragged(3, 35, [1, 7, 49, 50, 51, 180]):
        ref_m, ref_mask = ns.masking(a.copy()[-50:], 50)      # call site others/realformer.py:100-102
        m, mask = BO.masking_tail(a.copy(), 50)
        assert np.array_equal(m, ref_m) and np.array_equal(mask, ref_mask)


@needs_ref
def test_batching_oracle_stride_matches_reference_features(tmp_path):
    from oracle import batching_oracle as BO
    ns = refload.load("robot")
    for i, a in enumerate(_ragged(4, 40, [0, 3, 99, 100, 101, 250, 1000])):
        a = np.nan_to_num(a, nan=0.5, posinf=1.5, neginf=-1.5)    # the demo does not scrub
        np.save(tmp_path / f"u{i}.npy", a)
        ref_f, ref_mask = ns.audio_features(str(tmp_path) + "/", f"u{i}", 100)
        f, mask = BO.features_stride(a, 100)
        assert np.array_equal(f, ref_f) and np.array_equal(mask, ref_mask)
    b = _ragged(5, 768, [30])[0]
    b = np.nan_to_num(b, nan=0.0, posinf=0.0, neginf=0.0)
    np.save(tmp_path / "t.npy", b)
    ref_f, ref_mask = ns.text_features(str(tmp_path) + "/", "t", 25)
    f, mask = BO.features_stride(b, 25)
    assert np.array_equal(f, ref_f) and np.array_equal(mask, ref_mask)


@needs_ref
def test_batching_oracle_mosei_statistics_rows_match_reference_masking():
    """cmu-mosei/run.py:104-151 (is_bert=False): statistics rows + head / tail views; audio scrub."""
    from oracle import batching_oracle as BO
    ns = refload.load("mosei")
    for is_audio, D, m_len in ((False, 35, 100), (True, 74, 200), (False, 300, 20)):
        for a in _ragged(7 + D, D, [1, 5, m_len - 4, m_len - 3, m_len - 2, m_len + 40, 3 * m_len]):
            a = a.astype(np.float64)
            if not is_audio:        # only the acoustic stream is scrubbed; keep the others finite
                a = np.nan_to_num(a, nan=0.25, posinf=2.0, neginf=-2.0)
            ref_f, ref_m = ns.masking(a.copy(), m_len, is_bert=False, is_audio=is_audio)
            f, m = BO.mosei_masking(a.copy(), m_len, is_audio=is_audio)
            assert len(f) == len(ref_f) == (2 if len(a) >= m_len - 3 else 1)
            for x, y in zip(f + m, ref_f + ref_m):
                assert np.array_equal(x, y)


def test_batching_oracle_literal_cases():
    """Runs everywhere (no reference tree needed): hand-checked small cases."""
    from oracle import batching_oracle as BO
    a = np.arange(10, dtype=np.float32).reshape(5, 2)
    a[4, 1] = np.inf
    m, mask = BO.masking_tail(a, 3)
    assert m.tolist() == [[4, 5], [6, 7], [8, -71]] and mask.tolist() == [1, 1, 1]
    m, mask = BO.masking_tail(a[:2], 3)
    assert m.tolist() == [[0, 1], [2, 3], [0, 0]] and mask.tolist() == [1, 1, 0]
    f, mask = BO.features_stride(np.arange(14, dtype=np.float32).reshape(7, 2), 3)   # gap 2
    assert f.tolist() == [[0, 1], [4, 5], [8, 9]] and mask.tolist() == [1, 1, 1]
    f, mask = BO.features_stride(np.zeros((0, 2), np.float32), 2)
    assert f.tolist() == [[0, 0], [0, 0]] and mask.tolist() == [0, 0]
    f, mask = BO.head(np.arange(8, dtype=np.float32).reshape(4, 2), 3)
    assert f.tolist() == [[0, 1], [2, 3], [4, 5]] and mask.tolist() == [1, 1, 1]
    # cmu-mosei statistics rows: 4 rows, m_len 5 -> two views of 3 stats + 2 body rows
    a = np.array([[1., 8.], [2., 6.], [3., 4.], [6., 2.]])
    f, mask = BO.mosei_masking(a, 5)
    assert f[0].tolist() == [[6, 8], [1, 2], [3, 5], [1, 8], [2, 6]]
    assert f[1].tolist() == [[6, 8], [1, 2], [3, 5], [3, 4], [6, 2]] and mask[1].tolist() == [1] * 5
    f, mask = BO.mosei_masking(a[:1], 6)
    assert f[0].tolist() == [[1, 8], [1, 8], [1, 8], [1, 8], [0, 0], [0, 0]]
    assert len(f) == 1 and mask[0].tolist() == [1, 1, 1, 1, 0, 0]


def test_ragged_batch_pack_host_side():
    from mmemo_b200.batching import RaggedBatch
    seqs = [np.ones((3, 4)), None, np.zeros((0, 4)), 2 * np.ones((2, 4), dtype=np.float64)]
    rb = RaggedBatch.pack(seqs, pin=False)
    assert rb.n_rows.tolist() == [3, 0, 0, 2] and rb.row_start.tolist() == [0, 3, 3, 3]
    assert rb.flat.dtype == torch.float32 and rb.flat.shape == (5, 4) and float(rb.flat[4, 0]) == 2.0
    with pytest.raises(ValueError):
        RaggedBatch.pack([np.ones((1, 3)), np.ones((1, 4))], pin=False)
    with pytest.raises(ValueError):
        RaggedBatch.pack([None, None], pin=False)
    assert len(RaggedBatch.pack([None, None], dim=7, pin=False)) == 2
    with pytest.raises(RuntimeError):           # no CPU fallback
        rb.assemble(5)
