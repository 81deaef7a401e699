"""Freeze outputs of the REFERENCE's own classes into tests/golden/*.pt.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
For each case in tests/cases.py: construct the reference model under torch.manual_seed(0),
re-draw the zero-initialised gates (seed 1), run fwd + family loss + bwd in fp32 on CPU and save
{state, batch, logits, loss, grads, input_grads}.  The GPU box has no /root/reference; there the
oracle and the CUDA path are checked against these files.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import refload  # noqa: E402
from tests import cases  # noqa: E402


def main():
    assert refload.available(), "reference tree not found"
    torch.set_num_threads(1)
    for name, c in cases.CASES.items():
        ns = refload.load(c.family, **c.overrides)
        torch.manual_seed(0)
        model = c.ref_model(ns).float().train()
        model.load_state_dict(cases.seeded_state(model, seed=1))
        batch = c.batch()
        logits, loss, grads, igrads = cases.run_module_with_grads(model, c, batch, _RefLoss(ns))
        blob = {
            "state": {k: v.clone() for k, v in model.state_dict().items()},
            "batch": batch, "logits": logits, "loss": loss, "grads": grads, "input_grads": igrads,
            "torch": str(torch.__version__),
        }
        torch.save(blob, cases.golden_path(name))
        n = sum(v.numel() for v in blob["state"].values())
        print(f"{name}: params={n} logits={tuple(logits.shape)} loss={loss.item():.6f} "
              f"grads={len(grads)} -> {os.path.getsize(cases.golden_path(name)) / 1024:.0f} KiB")


class _RefLoss:
    """Loss functions taken from the reference itself where it defines them."""

    def __init__(self, ns):
        from oracle import mmemo_oracle as O
        self.multi_circle_loss = getattr(ns, "multi_circle_loss", None) or \
            refload.load("realformer").multi_circle_loss
        self.multi_loss = getattr(ns, "multi_loss", None) or (lambda p, t: self.multi_circle_loss(p, t).mean())
        # the R-Drop term is inline code in the reference's train() (Ren-MME/run.py:332-334), not a
        # function, so the oracle's restatement is used; it is 3 lines of F.kl_div.
        self.rdrop_kl = O.rdrop_kl


if __name__ == "__main__":
    main()
