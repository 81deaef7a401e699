"""Not a test: request latency of the 4-model robot_demo ensemble (BASELINE config 5) on one GPU —
members one after the other with a .cpu() each (robot_demo.py:610-614) vs the graphed ensemble."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmemo_b200  # noqa: E402
from mmemo_b200 import synth  # noqa: E402

DEV = "cuda"
kw = dict(dim=192, l_len=25, v_len=100, a_len=100, n_heads=6, n_layers=3, ffn=2)
models = []
for i in range(4):
    torch.manual_seed(i)
    m = mmemo_b200.robot_demo.Multi_class(**kw)
    m.load_state_dict(synth.randomize_gates(m.state_dict(), seed=10 + i))
    models.append(m.to(DEV).eval())


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


for prec in ("fp32", "bf16"):
    with mmemo_b200.precision(prec):
        for B in (1, 32):
            b = {k: v.to(DEV) for k, v in synth.robot_batch(seed=5, B=B).items()}
            args = [b[k] for k in mmemo_b200.robot_demo.Ensemble.NAMES]

            def sequential():
                with torch.no_grad():
                    ps = [m(*args).detach().cpu() for m in models]
                return (ps[0] + ps[1] + ps[2] + ps[3]) / 4

            ens = mmemo_b200.robot_demo.Ensemble(models)
            streams = mmemo_b200.robot_demo.Ensemble(models, use_graph=False)
            t_seq = timeit(sequential)
            t_str = timeit(lambda: streams(*args).cpu())
            t_gr = timeit(lambda: ens(*args).cpu())
            err = float((ens(*args).cpu() - sequential()).abs().max())
            print(f"{prec} B={B:2d}: members in sequence {t_seq:7.3f} ms | 4 streams {t_str:7.3f} ms | "
                  f"CUDA graph {t_gr:7.3f} ms per request  (max |diff| {err:.2e})", flush=True)
