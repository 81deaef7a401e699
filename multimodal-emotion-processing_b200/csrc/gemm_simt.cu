// Generic strided SIMT GEMM (fp32 accumulate) used (i) for the float32 parity mode everywhere and
// (ii) in the bf16 mode for shapes the tcgen05 kernel does not take (tiny N, ragged K, fp32 inputs).
//
//   C[m,n] (+)= epi( sum_k A[m*sAm + k*sAk] * B[n*sBn + k*sBk] )
//
// One kernel covers y = x w^T (NT), dx = dy w (NN) and dw = dy^T x (TN) through the strides.
// 64x64x16 tiles, 256 threads, 4x4 register micro-tile; split-K over gridDim.z with float atomics
// for the tall-skinny weight-gradient shapes (K' = B*L rows, M'xN' = a weight matrix).
#include "common.cuh"
#include "gemm.h"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;

template <typename TA, typename TB, typename TC>
__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const TA* __restrict__ A = static_cast<const TA*>(g.A);
  const TB* __restrict__ B = static_cast<const TB*>(g.B);
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  // split-K range of this CTA
  const int64_t kchunk = ((g.K + gridDim.z - 1) / gridDim.z + BK - 1) / BK * BK;
  const int64_t kbeg = (int64_t)blockIdx.z * kchunk;
  const int64_t kend = min(g.K, kbeg + kchunk);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const bool a_kfast = (g.sAk == 1), b_kfast = (g.sBk == 1);
  for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      int mm, kk;
      if (a_kfast) { mm = idx >> 4; kk = idx & 15; } else { kk = idx >> 6; mm = idx & 63; }
      const int64_t m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < g.M && k < kend) ? to_f(A[m * g.sAm + k * g.sAk]) : 0.f;
      int nn;
      if (b_kfast) { nn = idx >> 4; kk = idx & 15; } else { kk = idx >> 6; nn = idx & 63; }
      const int64_t n = n0 + nn, kb = k0 + kk;
      Bs[kk][nn] = (n < g.N && kb < kend) ? to_f(B[n * g.sBn + kb * g.sBk]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  TC* __restrict__ C = static_cast<TC*>(g.C);
  const bool first_split = (blockIdx.z == 0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (gridDim.z > 1) {  // split-K: linear epilogue terms once, atomics into a float32 C
        if (first_split) {
          if (g.bias) v += g.bias[n];
          if (g.pos) v += g.pos[(m % g.pos_period) * (g.ldpos ? g.ldpos : g.N) + n];
        }
        atomicAdd(reinterpret_cast<float*>(C) + m * g.ldc + n, v);
        continue;
      }
      if (g.bias) v += g.bias[n];
      if (g.pos) v += g.pos[(m % g.pos_period) * (g.ldpos ? g.ldpos : g.N) + n];
      if (g.relu) v = fmaxf(v, 0.f);
      if (g.relu_src) {
        const float r = g.relu_src_bf16
                            ? to_f(static_cast<const bf16*>(g.relu_src)[m * g.ldrelu + n])
                            : static_cast<const float*>(g.relu_src)[m * g.ldrelu + n];
        v = r > 0.f ? v : 0.f;
      }
      if (g.accumulate) v += to_f(C[m * g.ldc + n]);
      C[m * g.ldc + n] = from_f<TC>(v);
    }
  }
}

template <typename TA, typename TB, typename TC>
int launch(const GemmArgs& g, int splitk, cudaStream_t st) {
  dim3 grid((unsigned)cdiv(g.N, BN), (unsigned)cdiv(g.M, BM), (unsigned)splitk);
  gemm_simt_kernel<TA, TB, TC><<<grid, 256, 0, st>>>(g);
  MM_LAUNCH_OK();
  return MMEMO_OK;
}

}  // namespace

int gemm_simt(const GemmArgs& g, int a_bf16, int b_bf16, int c_bf16, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return MMEMO_OK;
  MM_REQUIRE(g.A && g.B && g.C && g.K >= 0);
  // split-K when the output grid cannot fill the GPU and K is long; needs a float32 C and no
  // non-linear epilogue
  int splitk = 1;
  const int64_t tiles = cdiv(g.M, BM) * cdiv(g.N, BN);
  // (also the small float32 head GEMMs: (B, 576..2304) x (9..96 outputs) are 3-8 tiles whose serial
  // K loop took 30-76 us on a handful of CTAs)
  if (!c_bf16 && !g.relu && !g.relu_src && tiles < 148 && g.K >= 128) {
    const int64_t want = cdiv(296, tiles), cap = g.K / 64;
    splitk = (int)(want < cap ? want : cap);
    if (splitk < 1) splitk = 1;
  }
  if (splitk > 1 && !g.accumulate) {
    MM_CUDA_OK(cudaMemset2DAsync(g.C, g.ldc * sizeof(float), 0, g.N * sizeof(float), g.M, st));
  }
  if (!a_bf16 && !b_bf16 && !c_bf16) return launch<float, float, float>(g, splitk, st);
  if (a_bf16 && b_bf16 && c_bf16) return launch<bf16, bf16, bf16>(g, splitk, st);
  if (!a_bf16 && b_bf16 && c_bf16) return launch<float, bf16, bf16>(g, splitk, st);
  if (a_bf16 && b_bf16 && !c_bf16) return launch<bf16, bf16, float>(g, splitk, st);
  if (a_bf16 && !b_bf16 && !c_bf16) return launch<bf16, float, float>(g, splitk, st);
  return MMEMO_ERR_ARG;
}
