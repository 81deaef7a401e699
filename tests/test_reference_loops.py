"""The reference's OWN training / validation loops, executed unchanged on the drop-in modules.

``train`` / ``valid`` are ast-extracted from the reference scripts (others/realformer.py:300-335,
cmu-mosei/run.py:354-391, Ren-MME/run.py:307-369) by oracle/refload.py - their source is executed as
it is, with the names the scripts get from their imports / module level supplied by the test
(``tqdm``, ``CLIP``, ``device``, the loss function and - on a box without CUDA - a ``torch``
stand-in whose ``torch.cuda.FloatTensor`` builds CPU tensors).  The model is OUR module, the loss is
OUR ``multi_circle_loss``, the optimizer is the fused ``mmemo_b200.optim`` one: this is the literal
meaning of "drop-in".

Where it runs: the reference tree exists only in the build container, which has no GPU, so there
the kernels are replaced by the launch recorder of test_host_plumbing (the loop's control flow,
tensor plumbing, autograd wiring, clip + optimizer calls are exercised; values are not).  With
both CUDA and the reference tree present the same test runs on the real kernels and checks that the
loss decreases.  On the GPU box (no /root/reference) it skips; numerical parity of the training
step is covered there by tests/test_gpu_train.py against the oracle.
"""
import types

import pytest
import torch

import mmemo_b200
from mmemo_b200 import ops, synth
from oracle import refload

pytestmark = pytest.mark.skipif(not refload.available(), reason="reference tree not present")
ON_GPU = torch.cuda.is_available()
DEV = "cuda" if ON_GPU else "cpu"


class _TorchStandIn(types.ModuleType):
    """``torch`` for a CUDA-less box: ``torch.cuda.FloatTensor/LongTensor`` build CPU tensors."""

    def __init__(self):
        super().__init__("torch")
        self.cuda = types.SimpleNamespace(FloatTensor=torch.FloatTensor, LongTensor=torch.LongTensor)

    def __getattr__(self, name):
        return getattr(torch, name)


@pytest.fixture()
def kernels(monkeypatch):
    """Real kernels on a GPU; the launch recorder otherwise."""
    calls = []
    if not ON_GPU:
        class _FakeLib:
            def mmemo_resattn_uses_tensor_cores(self, *a):
                return 0

            def mmemo_resattn_uses_mma(self, *a):
                return 1
        monkeypatch.setattr(ops, "_call", lambda name, *a: calls.append(name))
        monkeypatch.setattr(ops, "_try_call", lambda name, *a: calls.append(name) or True)
        monkeypatch.setattr(ops, "_need_cuda", lambda *a: None)
        monkeypatch.setattr(ops, "_stream", lambda: 0)
        monkeypatch.setattr(ops._lib, "load", lambda: _FakeLib())
        from mmemo_b200 import optim as mo
        monkeypatch.setattr(mo, "_check", lambda ts, what: None)
    mmemo_b200.ren_mme.DROP = 0.0
    ops.clear_shadow_cache()
    yield calls
    ops.clear_shadow_cache()


def _inject(loss_name, loss_fn):
    inj = {"tqdm": lambda it, desc=None: _Bar(it), "CLIP": 1.0, loss_name: loss_fn}
    if not ON_GPU:
        inj["torch"] = _TorchStandIn()
    return inj


class _Bar:
    """What the loops use of tqdm: iteration + set_description."""

    def __init__(self, it):
        self.it = it

    def __iter__(self):
        return iter(self.it)

    def set_description(self, s):
        self.last = s


def _rows(batch, keys):
    """list-of-samples batches, as the reference's data_loader yields them: tuples of ndarrays."""
    n = batch[keys[0]].shape[0]
    return [tuple(batch[k][i].numpy() for k in keys) for i in range(n)]


def test_realformer_train_and_valid_run_unchanged(kernels):
    from mmemo_b200 import optim as mo
    ns = refload.load("realformer", device=DEV, functions=("train", "valid"),
                      inject=_inject("multi_circle_loss", mmemo_b200.realformer.multi_circle_loss),
                      DROP=0.0)
    # (the injected loss must win over the reference's own def of the same name)
    ns._ns["multi_circle_loss"] = mmemo_b200.realformer.multi_circle_loss
    kw = dict(l_dim=300, v_dim=35, a_dim=74, dim=32, l_len=10, v_len=10, a_len=10, n_heads=2,
              n_layers=2, ffn=2)
    torch.manual_seed(0)
    model = mmemo_b200.realformer.State_Transfer(**kw)
    model.load_state_dict(synth.randomize_gates(model.state_dict(), seed=1))
    model = model.to(DEV)
    opt = mo.Adam(model.parameters(), lr=1e-3)
    keys = ("l", "v", "a", "label", "l_mask", "v_mask", "a_mask", "wmask")
    batches = [_rows(synth.realformer_batch(seed=s, B=4, P=3, L=(10, 10, 10)), keys)
               for s in range(3)]
    l0 = ns._ns["train"](model, batches, opt)
    tot, cnt, avg = ns._ns["valid"](model, batches)
    assert cnt == 3 and isinstance(l0, float)
    assert all(float(opt.state[p]["step"]) == 3 for p in model.parameters() if p in opt.state)
    if ON_GPU:
        l1 = ns._ns["train"](model, batches * 5, opt)
        assert l1 < l0
    else:
        assert any(n.startswith("mmemo_adam_step") for n in kernels)
        assert any(n.startswith("mmemo_resattn_bwd") for n in kernels)


def test_mosei_train_runs_unchanged(kernels):
    from mmemo_b200 import optim as mo
    ns = refload.load("mosei", device=DEV, functions=("train", "valid"),
                      inject=_inject("multi_circle_loss", mmemo_b200.cmu_mosei.multi_circle_loss),
                      DROP=0.0)
    ns._ns["multi_circle_loss"] = mmemo_b200.cmu_mosei.multi_circle_loss
    torch.manual_seed(0)
    model = mmemo_b200.cmu_mosei.Concat_Trans(32, 6, 8, 10, 2, 1, 1).to(DEV)
    opt = mo.AdamW(model.parameters(), lr=1e-3)
    src = ns._ns["train"].__code__.co_varnames
    b = synth.mosei_batch(seed=2, B=4, L=(6, 8, 10))
    # cmu-mosei/run.py:361: linguistic, visual, acoustic, l_mask, v_mask, a_mask, label = zip(*batch)
    keys = ("l", "v", "a", "l_mask", "v_mask", "a_mask", "label")
    batches = [_rows(b, keys)] * 2
    loss = ns._ns["train"](model, batches, opt)
    assert isinstance(loss, float) and "optimizer" in src
    _, cnt, _ = ns._ns["valid"](model, batches)
    assert cnt == 2


def test_renmme_train_with_rdrop_term_runs_unchanged(kernels):
    """Ren-MME/run.py:307-340: 13-field batches, multi_loss + the inline symmetric KL on OUR
    logits (plain torch ops on the drop-in's output), clip, AdamW."""
    from mmemo_b200 import optim as mo
    ns = refload.load("renmme", device=DEV, functions=("train",),
                      inject=_inject("multi_loss", mmemo_b200.ren_mme.multi_loss), DROP=0.0)
    ns._ns["multi_loss"] = mmemo_b200.ren_mme.multi_loss
    torch.manual_seed(0)
    model = mmemo_b200.ren_mme.Base_model(dim=32, l_len=5, v_len=7, a_len=11, n_heads=4).to(DEV)
    opt = mo.AdamW(model.parameters(), lr=1e-3)
    b = synth.renmme_batch(seed=3, B=4, L=(5, 7, 11))
    rows = [tuple(t[i].numpy() for t in b["inputs"]) + (b["label"][i].numpy(),) for i in range(4)]
    loss = ns._ns["train"](model, [rows, rows], opt)
    assert isinstance(loss, float)
