"""Not a test: GEMM micro-benchmark (graph replay, L2-warm) over shapes/modes. Run on the GPU box."""
import sys, math
import torch
sys.path.insert(0, ".")
from mmemo_b200 import ops

def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); g.replay(); g.replay(); e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / (2 * reps) * 1e3

shapes = [(8192, 512, 512), (8192, 1024, 512), (8192, 512, 1024), (8192, 1024, 1024), (8192, 2048, 512),
          (8192, 512, 2048), (16384, 512, 512), (4096, 512, 1024), (8192, 256, 1024), (8192, 1536, 512)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for (M, N, K) in shapes:
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
    dy = torch.randn(M, N, device="cuda").bfloat16()
    y = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    dx = torch.empty(M, K, device="cuda", dtype=torch.bfloat16)
    dw = torch.empty(N, K, device="cuda", dtype=torch.float32)
    fl = 2.0 * M * N * K
    t_f = timeit(lambda: ops._linear_fwd(True, x, w, None, None, y))
    t_x = timeit(lambda: ops._linear_bwd_x(True, dy, w, dx))
    t_w = timeit(lambda: ops._linear_bwd_w(True, dy, x, dw))
    t_ref = timeit(lambda: torch.matmul(x, w.t(), out=y))
    print(f"M{M} N{N} K{K}: fwd {t_f:6.1f}us {fl/t_f/1e6:6.0f}TF | bwd_x {t_x:6.1f}us {fl/t_x/1e6:6.0f}TF | "
          f"bwd_w {t_w:6.1f}us {fl/t_w/1e6:6.0f}TF | cuBLAS fwd {t_ref:6.1f}us {fl/t_ref/1e6:6.0f}TF", flush=True)
    b = torch.randn(N, device="cuda")
    t_b = timeit(lambda: ops._linear_fwd(True, x, w, b, None, y))
    t_br = timeit(lambda: ops._linear_fwd(True, x, w, b, None, y, relu=True))
    t_acc = timeit(lambda: ops._linear_bwd_x(True, dy, w, dx, accumulate=True))
    t_rs = timeit(lambda: ops._linear_bwd_x(True, dy, w, dx, relu_src=x))
    print(f"      fwd+bias {t_b:6.1f}us  fwd+bias+relu {t_br:6.1f}us  bwd_x+accumulate {t_acc:6.1f}us  bwd_x+relu_mask {t_rs:6.1f}us", flush=True)
