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
