// "Skinny" GEMM for the narrow models of the reference (d = 96 / 128 / 192: others/realformer.py,
// cmu-mosei/run.py, Ren-MME/run.py, robot_demo.py):  C[M,N] (+)= epi( A[M,K] * B ),  M = B*L rows
// (thousands), N, K <= 384.  These are HBM-bound (arithmetic intensity ~ N*K/(N+K) = 48..128 flop/B
// against a ridge of ~210): the whole weight fits in shared memory, so the kernel is a stream over
// the rows - a CTA loads W ONCE, then walks 64-row tiles of A (cp.async, double-buffered) through
// warp-level tensor-core MMAs (mma.sync m16n8k16, bf16, fp32 accumulate) and writes C from
// registers.  The 128 x 128-tile tcgen05 kernel (gemm_tc.cu) spends ~3-7 us per tile on these
// shapes (TMA round trip, TMEM epilogue, barriers) for 2.4 MFLOP of math; this one is bounded by
// the row stream.  Epilogues: bias, periodic position table, ReLU, ReLU mask from a saved
// activation, accumulate into C.  B is either K-major (y = x w^T: w as stored) or MN-major
// (dx = dy w: the same w, read transposed through ldmatrix.trans) - no transposes in HBM.
// GROUPED: up to SK_MAXP problems per launch (the nine chains of a trunk layer ...).
#include <stdlib.h>

#include "common.cuh"
#include "gemm.h"

namespace {

constexpr int SK_MAXP = 48;
constexpr int SK_WARPS = 8, SK_BM = 64, SK_KC = 128;     // 4 row groups x 2 column halves
constexpr size_t SK_SMEM_MAX = 200 * 1024;

struct SkProb {
  const bf16* A;
  const bf16* B;
  bf16* C;
  const float* bias;
  const float* pos;
  const bf16* relu_src;
  int lda, ldb, ldc, ldrelu;
  int M, N, K;
  int pos_period, b_mn, relu, accumulate;
  int cta_start, n_ctas;
};
struct SkTable {
  int n;
  int cta_start[SK_MAXP + 1];
  SkProb p[SK_MAXP];
};

__device__ __forceinline__ uint32_t sk_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void sk_cp16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz)
               : "memory");
}
__device__ __forceinline__ void sk_ldsm4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void sk_ldsm4t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void sk_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                       uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
      "{%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t sk_pack(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

// NT8 = number of 8-column MMA tiles one warp owns (its half of N, rounded up to 16 columns)
template <int NT8>
__global__ void __launch_bounds__(SK_WARPS * 32)
gemm_skinny_kernel(const __grid_constant__ SkTable T) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ SkProb P;
  {
    int p = 0;
    while (p + 1 < T.n && (int)blockIdx.x >= T.cta_start[p + 1]) ++p;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&T.p[p]);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&P);
    for (int i = threadIdx.x; i < (int)(sizeof(SkProb) / 4); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const int M = P.M, N = P.N, K = P.K;
  const int cta = (int)blockIdx.x - P.cta_start, n_ctas = P.n_ctas;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int rg = warp & 3, nh = warp >> 2;
  // shared memory: W (padded rows: +16 B keeps the 8 row addresses of an ldmatrix in distinct
  // bank groups for every row length used here) | two A stages [64][KC]
  const int wrows = P.b_mn ? K : N, wcols = P.b_mn ? N : K;
  const uint32_t wstride = (uint32_t)wcols * 2 + 16;
  const int KC = K < SK_KC ? K : SK_KC;
  const uint32_t astride = (uint32_t)KC * 2 + 16;
  const uint32_t sW = sk_u32(smem);
  const uint32_t sA0 = sW + ((wrows * wstride + 127u) & ~127u);
  const uint32_t a_bytes = SK_BM * astride;

  pdl_wait();
  pdl_trigger();
  // ---- W: once per CTA -----------------------------------------------------------------------
  {
    const int cpr = wcols >> 3;
    for (int i = threadIdx.x; i < wrows * cpr; i += SK_WARPS * 32) {
      const int r = i / cpr, c = i - r * cpr;
      sk_cp16(sW + r * wstride + c * 16, P.B + (size_t)r * P.ldb + c * 8, true);
    }
  }
  const int n_tiles = (M + SK_BM - 1) / SK_BM;
  const int n_kc = (K + KC - 1) / KC;
  const int my_tiles = cta < n_tiles ? (n_tiles - cta + n_ctas - 1) / n_ctas : 0;
  const int n_it = my_tiles * n_kc;                  // (row tile, k chunk) steps of this CTA
  auto load_a = [&](int it, int stage) {
    const int tile = cta + (it / n_kc) * n_ctas, kc = it % n_kc;
    const int k0 = kc * KC, kw = min(KC, K - k0);
    const int cpr = kw >> 3;
    const uint32_t dst = sA0 + stage * a_bytes;
    for (int i = threadIdx.x; i < SK_BM * cpr; i += SK_WARPS * 32) {
      const int r = i / cpr, c = i - r * cpr;
      const int row = tile * SK_BM + r;
      sk_cp16(dst + r * astride + c * 16, P.A + (size_t)(row < M ? row : 0) * P.lda + k0 + c * 8,
              row < M);
    }
  };
  if (n_it > 0) load_a(0, 0);
  asm volatile("cp.async.commit_group;" ::: "memory");

  const int ncol0 = nh * (NT8 * 8);                  // first column of this warp's half
  float acc[NT8][4];
  for (int it = 0; it < n_it; ++it) {
    const int stage = it & 1;
    if (it + 1 < n_it) load_a(it + 1, stage ^ 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    const int tile = cta + (it / n_kc) * n_ctas, kc = it % n_kc;
    const int k0 = kc * KC, kw = min(KC, K - k0);
    if (kc == 0) {
#pragma unroll
      for (int n = 0; n < NT8; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[n][e] = 0.f;
    }
    const uint32_t sA = sA0 + stage * a_bytes;
    if (ncol0 < N) {
      for (int ks = 0; ks < kw; ks += 16) {
        uint32_t af[4];
        sk_ldsm4(sA + (rg * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * astride +
                     (ks >> 3) * 16 + (lane >> 4) * 16, af);
        const int kk = k0 + ks;
#pragma unroll
        for (int n2 = 0; n2 < NT8; n2 += 2) {
          const int nc = ncol0 + n2 * 8;
          if (nc >= N) break;
          uint32_t bf[4];
          const int mi = lane >> 3, r = lane & 7;
          if (!P.b_mn)      // W[n][k]: rows = n, 16-byte chunks along k
            sk_ldsm4(sW + (nc + r + 8 * (mi >> 1)) * wstride + ((kk >> 3) + (mi & 1)) * 16, bf);
          else              // W[k][n]: rows = k, chunks along n, transposed on the way in
            sk_ldsm4t(sW + (kk + r + 8 * (mi & 1)) * wstride + ((nc >> 3) + (mi >> 1)) * 16, bf);
          sk_mma(acc[n2], af, bf[0], bf[1]);
          sk_mma(acc[n2 + 1], af, bf[2], bf[3]);
        }
      }
    }
    if (kc == n_kc - 1 && ncol0 < N) {
      // ---- epilogue from registers ---------------------------------------------------------------
      const int rowA = tile * SK_BM + rg * 16 + g, rowB = rowA + 8;
#pragma unroll
      for (int n = 0; n < NT8; ++n) {
        const int c = ncol0 + n * 8 + 2 * t;
        if (c >= N) break;
        float b0 = 0.f, b1 = 0.f;
        if (P.bias) { b0 = P.bias[c]; b1 = P.bias[c + 1]; }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int row = half ? rowB : rowA;
          if (row >= M) continue;
          float v0 = acc[n][half * 2] + b0, v1 = acc[n][half * 2 + 1] + b1;
          if (P.pos) {
            const float* pr = P.pos + (size_t)(row % P.pos_period) * N + c;
            v0 += pr[0]; v1 += pr[1];
          }
          if (P.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
          if (P.relu_src) {
            const uint32_t m = *reinterpret_cast<const uint32_t*>(P.relu_src + (size_t)row * P.ldrelu + c);
            if (!(__uint_as_float(m << 16) > 0.f)) v0 = 0.f;
            if (!(__uint_as_float(m & 0xFFFF0000u) > 0.f)) v1 = 0.f;
          }
          uint32_t* dst = reinterpret_cast<uint32_t*>(P.C + (size_t)row * P.ldc + c);
          if (P.accumulate) {
            const uint32_t old = *dst;
            v0 += __uint_as_float(old << 16);
            v1 += __uint_as_float(old & 0xFFFF0000u);
          }
          *dst = sk_pack(v0, v1);
        }
      }
    }
    __syncthreads();     // the stage just read is refilled by the next iteration's prefetch
  }
}

inline bool sk_a16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

size_t sk_smem(const GemmArgs& g, bool b_mn) {
  const int64_t wrows = b_mn ? g.K : g.N, wcols = b_mn ? g.N : g.K;
  const int64_t KC = g.K < SK_KC ? g.K : SK_KC;
  return (size_t)(((wrows * (wcols * 2 + 16) + 127) & ~127ll) + 2 * SK_BM * (KC * 2 + 16));
}

}  // namespace

// C bf16, A K-major bf16, B bf16 K-major or MN-major, N and K multiples of 16 and <= 384.
bool gemm_skinny_supported(const GemmArgs& g, int c_bf16) {
  static const bool off = getenv("MMEMO_NO_SKINNY") != nullptr;      // A/B switch for measurements
  if (off) return false;
  if (!c_bf16 || g.M < 1 || g.N < 16 || g.K < 16 || g.N > 384 || g.K > 512) return false;
  if (g.N % 16 || g.K % 16) return false;
  if (g.sAk != 1 || g.sAm % 8) return false;
  const bool b_k = (g.sBk == 1 && g.sBn % 8 == 0), b_mn = (g.sBn == 1 && g.sBk % 8 == 0);
  if (!b_k && !b_mn) return false;
  if (!sk_a16(g.A) || !sk_a16(g.B) || (reinterpret_cast<uintptr_t>(g.C) & 3) || g.ldc % 2) return false;
  if (g.relu_src && (!g.relu_src_bf16 || (reinterpret_cast<uintptr_t>(g.relu_src) & 3) ||
                     g.ldrelu % 2))
    return false;
  if (g.M > (1ll << 31) - 512) return false;
  return sk_smem(g, !b_k) <= SK_SMEM_MAX;
}

int gemm_skinny_grouped(const GemmArgs* gs, int n, cudaStream_t st) {
  if (n < 1 || n > SK_MAXP) return MMEMO_ERR_ARG;
  static thread_local SkTable T;
  T.n = n;
  size_t smem = 0;
  int64_t nmax = 0, tiles_total = 0;
  for (int i = 0; i < n; ++i) {
    if (!gemm_skinny_supported(gs[i], 1)) return MMEMO_ERR_SHAPE;
    tiles_total += cdiv(gs[i].M, SK_BM);
    nmax = gs[i].N > nmax ? gs[i].N : nmax;
  }
  // CTAs: about three per SM over the group, shared out by row count (each loads W once, so a
  // CTA should own several row tiles), at least one per problem
  const int64_t budget = 148 * 3;
  int ctas = 0;
  for (int i = 0; i < n; ++i) {
    const GemmArgs& g = gs[i];
    const bool b_mn = !(g.sBk == 1 && g.sBn % 8 == 0);
    SkProb& p = T.p[i];
    p = SkProb{};
    p.A = static_cast<const bf16*>(g.A); p.B = static_cast<const bf16*>(g.B);
    p.C = static_cast<bf16*>(g.C); p.bias = g.bias; p.pos = g.pos;
    p.relu_src = static_cast<const bf16*>(g.relu_src);
    p.lda = (int)g.sAm; p.ldb = (int)(b_mn ? g.sBk : g.sBn); p.ldc = (int)g.ldc;
    p.ldrelu = (int)g.ldrelu; p.M = (int)g.M; p.N = (int)g.N; p.K = (int)g.K;
    p.pos_period = (int)(g.pos ? g.pos_period : 1); p.b_mn = b_mn; p.relu = g.relu;
    p.accumulate = g.accumulate;
    const int64_t tiles = cdiv(g.M, SK_BM);
    int64_t c = (tiles * budget + tiles_total - 1) / tiles_total;
    if (c > tiles) c = tiles;
    if (c < 1) c = 1;
    p.cta_start = ctas; p.n_ctas = (int)c;
    T.cta_start[i] = ctas;
    ctas += (int)c;
    const size_t s = sk_smem(g, b_mn);
    smem = s > smem ? s : smem;
  }
  T.cta_start[n] = ctas;
  const int nt8 = (int)(cdiv(nmax, 32) * 2);           // per-warp half of N in 8-column tiles (even)
#define SK_LAUNCH(NT)                                                                            \
  {                                                                                              \
    MM_CUDA_OK(cudaFuncSetAttribute(gemm_skinny_kernel<NT>,                                      \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
    MM_CUDA_OK(mm_launch(gemm_skinny_kernel<NT>, dim3((unsigned)ctas), dim3(SK_WARPS * 32), smem, \
                         st, T));                                                                \
  }
  if (nt8 <= 6) SK_LAUNCH(6) else if (nt8 <= 8) SK_LAUNCH(8) else if (nt8 <= 12) SK_LAUNCH(12)
  else if (nt8 <= 16) SK_LAUNCH(16) else SK_LAUNCH(24)
#undef SK_LAUNCH
  return MMEMO_OK;
}
