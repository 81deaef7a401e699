"""Host-side epoch driver (mmemo_b200/trainer.py) against the behaviour of the reference's run()
(others/realformer.py:338-389): fold slicing, plateau schedule, best-checkpoint naming, early stop,
log format.  CPU only — the driver is model-agnostic."""
import os

import torch

from mmemo_b200 import trainer


def test_kfold_equals_reference_slicing():
    for n in (5, 10, 13, 101, 2248):
        names = list(range(n))
        folds = trainer.kfold_splits(names, 5)
        # literal restatement of others/realformer.py:366-389
        ref_valid = [names[:int(n * 0.2)], names[int(n * 0.2):int(n * 0.4)],
                     names[int(n * 0.4):int(n * 0.6)], names[int(n * 0.6):int(n * 0.8)],
                     names[int(n * 0.8):]]
        ref_train = [names[int(n * 0.2):], names[:int(n * 0.2)] + names[int(n * 0.4):],
                     names[:int(n * 0.4)] + names[int(n * 0.6):],
                     names[:int(n * 0.6)] + names[int(n * 0.8):], names[:int(n * 0.8)]]
        assert [v for _, v in folds] == ref_valid
        assert [t for t, _ in folds] == ref_train


class _Scripted(torch.nn.Module):
    """A model whose validation loss follows a script, to drive the schedule deterministically."""

    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(3))


def test_fit_plateau_checkpoints_and_early_stop(tmp_path):
    model = _Scripted()
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    script = iter([1.314159, 1.25, 1.30, 1.31, 1.32, 1.33, 1.34, 0.5])
    state = {}

    def train_step(m, batch):
        return (m.w * batch).sum() + 1.0

    def valid_step(m, batch):
        if "cur" not in state or state["fresh"]:
            state["cur"], state["fresh"] = next(script), False
        return torch.tensor(state["cur"], dtype=torch.float64)

    def valid_iter():
        state["fresh"] = True
        return [torch.ones(3)] * 2

    msgs = []
    hist = trainer.fit(model, opt, lambda: [torch.ones(3)] * 3, valid_iter, train_step, valid_step,
                       epochs=20, name="model_1", log_dir=str(tmp_path), sched_patience=2,
                       log=msgs.append)
    # best at epoch 2 (1.25); epochs 3..6 do not improve -> stop after the 4th miss
    assert len(hist["valid"]) == 6 and hist["valid"][:3] == [1.314159, 1.25, 1.30]
    assert sorted(os.listdir(tmp_path)) == ["model_1.txt", "model_1_1.25.pt", "model_1_1.31.pt"]
    assert hist["best"].endswith("model_1_1.25.pt")
    # patience 2: the lr drops by 10x after the third epoch without improvement (epoch 5)
    assert hist["lrs"][:4] == [1e-2] * 4 and abs(hist["lrs"][4] - 1e-3) < 1e-12
    assert any("reducing learning rate" in m for m in msgs)
    lines = open(tmp_path / "model_1.txt").read().split("\n")
    assert lines[0] == "epoch, train_loss, valid_loss" and lines[2] == "1, 0.97, 1.31"
    sd = torch.load(hist["best"])
    assert set(sd) == {"w"}
    # the training steps really ran: 6 epochs x 3 batches of Adam on d/dw = 1
    assert float(model.w.detach().abs().min()) > 0


def test_fit_without_log_dir_and_with_self_clipping_optimizer():
    model = _Scripted()
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    seen = []

    class Writer:
        def add_scalars(self, name, d, epoch):
            seen.append((name, sorted(d), epoch))

    hist = trainer.fit(model, opt, lambda: [torch.ones(3)], lambda: [torch.ones(3)],
                       lambda m, b: ((m.w - b) ** 2).sum(), epochs=3, name="m", clip=None,
                       writer=Writer(), log=lambda s: None)
    assert hist["best"] is None and len(hist["train"]) == 3
    assert hist["valid"][2] < hist["valid"][0]
    assert seen == [("m", ["train_loss", "valid_loss"], e) for e in range(3)]
