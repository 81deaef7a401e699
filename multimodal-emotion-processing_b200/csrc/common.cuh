// Shared device/host helpers for libmmemo (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mmemo.h"

typedef __nv_bfloat16 bf16;

// ---- error plumbing -------------------------------------------------------------------------
void mmemo_set_error(const char* what, const char* file, int line);
#define MM_CUDA_OK(expr)                                   \
  do {                                                     \
    cudaError_t _e = (expr);                               \
    if (_e != cudaSuccess) {                               \
      mmemo_set_error(cudaGetErrorString(_e), __FILE__, __LINE__); \
      return MMEMO_ERR_CUDA;                               \
    }                                                      \
  } while (0)
#define MM_LAUNCH_OK() MM_CUDA_OK(cudaGetLastError())
#define MM_REQUIRE(cond)            \
  do {                              \
    if (!(cond)) return MMEMO_ERR_ARG; \
  } while (0)

static inline cudaStream_t mm_stream(mmemo_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- programmatic dependent launch ----------------------------------------------------------
// The hot kernels are launched with programmaticStreamSerialization: their CTAs may become
// resident and run their prologue (smem carve-up, mbarrier init, descriptor prefetch) while the
// previous kernel on the stream drains, and pdl_wait() then blocks until that kernel has
// completed and its writes are visible.  Every thread of such a kernel calls pdl_wait() before
// its first global-memory access (reads AND writes: buffers are recycled between kernels) and
// before any early return, so completion order stays transitive along the stream.
// Per-stream launch settings (lib.cu; mmemo_stream_set_*): split-K scratch of the tcgen05 GEMM,
// SMs the persistent kernels may occupy (0 = all), programmatic dependent launch on/off.
struct MmStreamCfg {
  float* ws = nullptr;
  size_t ws_bytes = 0;
  int sm_budget = 0;
  int pdl = 1;
};
MmStreamCfg mm_stream_cfg(cudaStream_t st);
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t mm_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                    cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = mm_stream_cfg(st).pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- element conversion ---------------------------------------------------------------------
__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(bf16 x) { return __bfloat162float(x); }
template <typename T>
__device__ __forceinline__ T from_f(float x);
template <>
__device__ __forceinline__ float from_f<float>(float x) { return x; }
template <>
__device__ __forceinline__ bf16 from_f<bf16>(float x) { return __float2bfloat16_rn(x); }
// value after a round trip through T (fp32: identity; bf16: RNE rounding)
template <typename T>
__device__ __forceinline__ float round_to(float x) { return to_f(from_f<T>(x)); }

// ---- warp / block reductions ----------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
