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
