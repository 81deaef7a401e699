"""Device-side batch assembly (csrc/assemble.cu, mmemo_b200/batching.py) vs the numpy oracle of the
reference's host code (oracle/batching_oracle.py; pinned to others/realformer.py:72-82 and
robot_demo.py:119-150 in tests/test_oracle.py).  Bit-exact: it is a gather."""
import numpy as np
import pytest
import torch

from mmemo_b200.batching import RaggedBatch
from oracle import batching_oracle as BO

pytestmark = pytest.mark.gpu


def _ragged(seed, D, lens, bad=True):
    rng = np.random.default_rng(seed)
    seqs = []
    for T in lens:
        a = rng.standard_normal((T, D)).astype(np.float32)
        if T and bad:
            m = rng.random((T, D)) < 0.02
            a[m] = rng.choice(np.array([np.nan, np.inf, -np.inf], dtype=np.float32), m.sum())
        seqs.append(a)
    return seqs


LENS = [0, 1, 7, 49, 50, 51, 180, 0, 1000, 25, 100, 101]


@pytest.mark.parametrize("D", [35, 74, 300, 40, 768])
@pytest.mark.parametrize("mode,m_len", [("tail", 50), ("stride", 100), ("stride", 25), ("head", 50)])
def test_assemble_equals_reference_host_code(D, mode, m_len):
    scrub = mode == "tail"                     # realformer scrubs; the demo does not
    seqs = _ragged(D + m_len, D, LENS, bad=scrub)
    fn = {"tail": BO.masking_tail, "stride": BO.features_stride, "head": BO.head}[mode]
    exp_x, exp_m = [], []
    for a in seqs:
        if len(a) == 0:                        # 'no_name' slot: others/realformer.py:108-113
            x, m = np.zeros((m_len, D)), np.zeros(m_len)
        else:
            x, m = fn(a.copy(), m_len)
        exp_x.append(x)
        exp_m.append(m)
    exp_x = torch.from_numpy(np.stack(exp_x)).float()
    exp_m = torch.from_numpy(np.stack(exp_m)).float()
    x, mask = RaggedBatch.pack(seqs, dim=D).cuda().assemble(m_len, mode, scrub=-71.0 if scrub else None)
    assert x.shape == (len(LENS), m_len, D) and mask.shape == (len(LENS), m_len)
    assert torch.equal(x.cpu(), exp_x)
    assert torch.equal(mask.cpu(), exp_m)


def test_assemble_windows_feed_the_model():
    """realformer batch (B, P) windows: assemble -> State_Transfer forward works on the result and
    all-masked windows behave like the reference's zero-filled 'no_name' slots."""
    import mmemo_b200
    from mmemo_b200 import synth
    B, P, L = 4, 6, 50
    dims = (300, 35, 74)
    rng = np.random.default_rng(0)
    inputs = []
    for D in dims:
        seqs = [None if (i % P) >= 4 else rng.standard_normal((int(rng.integers(1, 90)), D)).astype(np.float32)
                for i in range(B * P)]
        inputs.append(RaggedBatch.pack(seqs, dim=D).cuda().assemble(L, "tail", lead_shape=(B, P)))
    (l, lm), (v, vm), (a, am) = inputs
    assert l.shape == (B, P, L, 300) and lm.shape == (B, P, L)
    assert float(lm[:, 4:].sum()) == 0.0 and float(l[:, 4:].abs().sum()) == 0.0
    torch.manual_seed(0)
    model = mmemo_b200.realformer.State_Transfer(300, 35, 74, 96, L, L, L, 6, 1, 2)
    model.load_state_dict(synth.randomize_gates(model.state_dict(), seed=1))
    model = model.cuda().eval()
    with torch.no_grad():
        out = model(l, v, a, lm, vm, am)
    assert out.shape == (B, P, 6) and bool(torch.isfinite(out).all())


@pytest.mark.parametrize("D,m_len,audio", [(300, 20, False), (35, 100, False), (74, 200, True)])
def test_assemble_stats_equals_cmu_mosei_masking(D, m_len, audio):
    """cmu-mosei/run.py:104-151: three statistics rows (max / min / mean over the whole sample) +
    head / tail views.  max, min and the body rows are bit-exact; the mean is a float64 sum
    rounded to float32 on both sides (numpy sums pairwise, the kernel sequentially: <= 1 ulp)."""
    lens = [0, 1, 5, m_len - 4, m_len - 3, m_len - 2, m_len + 40, 3 * m_len, 2]
    seqs = _ragged(D + m_len, D, lens, bad=audio)
    rb = RaggedBatch.pack(seqs, dim=D)
    two = rb.two_views(m_len).tolist()
    assert two == [len(a) >= m_len - 3 for a in seqs]
    dev = rb.cuda()
    for view in ("head", "tail"):
        x, mask = dev.assemble_stats(m_len, view, scrub=-71.0 if audio else None)
        x, mask = x.cpu(), mask.cpu()
        for i, a in enumerate(seqs):
            if len(a) == 0:
                assert float(x[i].abs().sum()) == 0.0 and float(mask[i].sum()) == 0.0
                continue
            f, m = BO.mosei_masking(a.astype(np.float64), m_len, is_audio=audio)
            ef = torch.from_numpy(f[-1] if view == "tail" else f[0]).float()
            em = torch.from_numpy(m[-1] if view == "tail" else m[0]).float()
            assert torch.equal(mask[i], em), (i, view)
            assert torch.equal(x[i, :2], ef[:2]) and torch.equal(x[i, 3:], ef[3:]), (i, view)
            assert torch.allclose(x[i, 2], ef[2], rtol=2e-7, atol=1e-9), (i, view)
