"""Not a test: tcgen05 GEMM diagnostics (run on the GPU box)."""
import sys, math
import torch
sys.path.insert(0, ".")
from mmemo_b200 import ops, _lib
from tests.cases import rel_err

torch.manual_seed(0)
L = _lib.load()
for (M, N, K) in [(128, 128, 64), (256, 128, 128), (1024, 512, 512), (8192, 1024, 512), (640, 96, 304), (8192, 512, 1024)]:
    x = torch.randn(M, K).bfloat16()
    w = torch.randn(N, K) / math.sqrt(K)
    dy = torch.randn(M, N).bfloat16()
    xc = x.cuda().requires_grad_(True); wc = w.cuda().requires_grad_(True)
    ops.clear_shadow_cache()
    uses = [L.mmemo_gemm_uses_tensor_cores(M, N, K, K, K, N, 0), L.mmemo_gemm_uses_tensor_cores(M, N, K, N, K, K, 1),
            L.mmemo_gemm_uses_tensor_cores(M, N, K, N, K, K, 2)]
    try:
        y = ops.linear(xc, wc, bf16=True)
        torch.cuda.synchronize()
        ref = x.float() @ w.bfloat16().float().t()
        e_f = rel_err(y.float(), ref)
        y.backward(dy.cuda())
        torch.cuda.synchronize()
        e_x = rel_err(xc.grad.float(), dy.float() @ w.bfloat16().float())
        e_w = rel_err(wc.grad, dy.float().t() @ x.float())
        print(f"M{M} N{N} K{K} tc={uses} fwd {e_f:.3e} bwd_x {e_x:.3e} bwd_w {e_w:.3e}", flush=True)
        if e_f > 2e-2:
            d = (y.float().cpu() - ref).abs()
            bad = (d > 0.05 * ref.abs().max()).nonzero()
            print("  fwd bad count", len(bad), "first", bad[:5].tolist(), "rows%128 hist", torch.bincount(bad[:, 0] % 128, minlength=128)[:16].tolist(), "cols%128", torch.bincount(bad[:, 1] % 128, minlength=128)[:16].tolist())
    except Exception as ex:
        print(f"M{M} N{N} K{K} tc={uses} EXC {type(ex).__name__}: {str(ex)[:300]}", flush=True)
        break
