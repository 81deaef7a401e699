"""Not a test: prints the margins of tests/test_gpu_train.py (run-to-run spread of the loss-curve
comparisons) — `python tools/loss_curve_margins.py [reps]` on the GPU box.  Besides the distance
to the oracle it prints the distance of every repetition to OUR OWN first repetition: the kernels
sum in a run-dependent order (atomics in the LayerNorm parameter gradients, split-K), and at the
reference learning rate the 200-step trajectory amplifies last-bit differences, so this self-spread
is the noise floor the oracle comparison has to be read against."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import test_gpu_train as T  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ref = {1e-4: T.oracle_curve(1e-4), 1e-3: T.oracle_curve(1e-3)}
for prec in ("fp32", "bf16"):
    first = {}
    for r in range(reps):
        a = T.our_curve(1e-4, prec)
        per_step = ((a - ref[1e-4]).abs() / ref[1e-4].abs()).max().item()
        b = T.our_curve(1e-3, prec)
        sm = ((T.smooth(b) - T.smooth(ref[1e-3])).abs() / T.smooth(ref[1e-3])).max().item()
        first.setdefault("a", a)
        first.setdefault("b", b)
        self_step = ((a - first["a"]).abs() / first["a"].abs()).max().item()
        self_sm = ((T.smooth(b) - T.smooth(first["b"])).abs() / T.smooth(first["b"])).max().item()
        print(f"{prec} run {r}: vs oracle per-step lr1e-4 {per_step:.5f}  smoothed lr1e-3 {sm:.5f}"
              f"   |   vs our run 0: per-step lr1e-4 {self_step:.5f}  smoothed lr1e-3 {self_sm:.5f}",
              flush=True)
