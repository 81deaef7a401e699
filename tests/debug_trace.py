"""Not a test: per-CTA timeline of the tcgen05 GEMM (debug stamps), run on the GPU box."""
import sys, math, ctypes
import torch
sys.path.insert(0, ".")
from mmemo_b200 import ops, _lib
L = _lib.load()
L.mmemo_debug_set_gemm_trace.argtypes = [ctypes.c_void_p]
M, N, K = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (8192, 512, 512)))
x = torch.randn(M, K, device="cuda").bfloat16()
w = torch.randn(N, K, device="cuda") / math.sqrt(K)
for _ in range(3):
    y = ops.linear(x, w, bf16=True)
torch.cuda.synchronize()
ncta = ((M + 127) // 128) * ((N + 127) // 128)
tr = torch.zeros(ncta * 8, dtype=torch.int64, device="cuda")
L.mmemo_debug_set_gemm_trace(tr.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); y = ops.linear(x, w, bf16=True); e1.record()
torch.cuda.synchronize()
L.mmemo_debug_set_gemm_trace(None)
t = tr.view(ncta, 8).cpu().double()
t0 = t[:, 0].min()
rel = (t - t0) / 1e3
names = ["entry", "setup done", "1st TMA issued", "1st stage landed", "MMAs issued", "acc ready", "pass1 done", "epilogue done"]
print(f"M{M} N{N} K{K} ctas {ncta} event time {e0.elapsed_time(e1)*1e3:.1f} us; kernel span {(t[:,7].max()-t0)/1e3:.1f} us")
for i, n in enumerate(names):
    print(f"  {n:18s} min {rel[:, i].min():7.2f}  median {rel[:, i].median():7.2f}  max {rel[:, i].max():7.2f} us")
d = (t[:, 1:] - t[:, :-1]) / 1e3
print("  per-CTA phase durations (median us):", [round(float(v), 2) for v in d.median(0)[0]])
print("  per-CTA total (median/max us):", float((t[:, 7] - t[:, 0]).median() / 1e3), float((t[:, 7] - t[:, 0]).max() / 1e3))
