// Fused residual-attention core on tcgen05 / TMA (bf16, head_dim 64, Lq = Lk = 128): kernel (a).
//
// One CTA per (batch, head).  Everything between the projected Q/K/V and the merged-head output
// happens on chip; the only score-sized HBM traffic is the mandatory read of S_prev and write of S:
//
//   TMA  : Q, K, V tiles (128 x 64 bf16, SWIZZLE_128B) and the S_prev tile (128 x 128 bf16)
//   UMMA : S_acc = Q K^T            (128 x 128 x 64, fp32 accumulator in 128 TMEM columns)
//   regs : thread = query row; tcgen05.ld the row, s = acc/sqrt(hd) + c*S_prev - 1e8*(1-mask) in the
//          reference's fp32 op order, round to bf16, write S back into the S_prev tile in shared
//          memory (in place), exact single-pass softmax (row max, exp, sum)
//   TMA  : store S for the next layer / backward
//   UMMA : O_acc = P V              (128 x 64 x 128; P bf16 K-major, V MN-major as loaded)
//   TMA  : store O / rowsum into the (B, Lq, H*hd) merged-head layout
//
// 192 threads: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 = softmax /
// epilogue.  80 KB shared memory and 256 TMEM columns per CTA -> two CTAs per SM so one tile's
// loads and stores overlap the other's math.
#include <math.h>

#include "common.cuh"
#include "resattn.h"
#include "tc_common.cuh"

namespace {

constexpr int L = 128, HD = 64;
constexpr uint32_t TILE_QKV = L * HD * 2;      // 16 KB
constexpr uint32_t TILE_S = L * L * 2;         // 32 KB (two 64-column halves of 16 KB)
constexpr int NTHREADS = 192;
// smem map (offsets from the 1024-aligned base)
constexpr uint32_t OFF_Q = 0, OFF_K = TILE_QKV, OFF_V = 2 * TILE_QKV, OFF_S = 3 * TILE_QKV;
constexpr uint32_t OFF_P = OFF_Q;              // P (32 KB) reuses Q|K once S = QK^T has completed
constexpr uint32_t OFF_O = OFF_V;              // O staging reuses V once O = PV has completed
constexpr uint32_t OFF_MASK = OFF_S + TILE_S;  // 128 floats
constexpr uint32_t OFF_BAR = OFF_MASK + 512;
constexpr uint32_t SMEM_FWD = OFF_BAR + 128 + 1024;

struct FwdArgs {
  const float* mask;   // (B, Lk) or null
  int64_t mask_bs;
  const float* c;      // device scalar or null
  float* stat;         // (B, H, Lq, 2)
  int H;
  int has_prev, write_s;
  int pf_dist;   // L2-prefetch the tile of CTA (linear id + pf_dist); 0 = off
  float sqrt_hd;
  uint32_t idesc_qk, idesc_pv;
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }

__global__ void __launch_bounds__(NTHREADS, 2)
resattn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV,
                      const __grid_constant__ CUtensorMap tmSprev,
                      const __grid_constant__ CUtensorMap tmSout,
                      const __grid_constant__ CUtensorMap tmO, const FwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = tc::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const uint32_t bar_qk = base + OFF_BAR, bar_v = bar_qk + 8, bar_sp = bar_qk + 16,
                 bar_s = bar_qk + 24, bar_p = bar_qk + 32, bar_o = bar_qk + 40;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + OFF_BAR + 64);
  float* mask_s = reinterpret_cast<float*>(gbase + OFF_MASK);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, b = blockIdx.y;
  const int row0 = b * L;                       // first row of this batch in the (B*L, ld) views
  const int srow0 = (b * a.H + h) * L;          // first row in the (B*H*L, L) score views

  if (threadIdx.x == 0) {
    tc::mbar_init(bar_qk, 1);
    tc::mbar_init(bar_v, 1);
    tc::mbar_init(bar_sp, 1);
    tc::mbar_init(bar_s, 1);
    tc::mbar_init(bar_p, 128);
    tc::mbar_init(bar_o, 1);
    tc::fence_barrier_init();
  }
  pdl_trigger();
  if (warp == 1) {        // TMEM allocation overlaps the previous kernel's drain
    tc::tmem_alloc(base + OFF_BAR + 64, 256);
    tc::tmem_relinquish();
  }
  pdl_wait();             // before the first global access (mask row, TMA loads)
  if (threadIdx.x >= 64) {  // softmax threads stage the mask row (same for every query)
    const int j = threadIdx.x - 64;
    // additive mask term 1e8*(1-m), computed once per key (same fp32 ops as the reference)
    mask_s[j] = a.mask ? __fmul_rn(1.0e8f, __fsub_rn(1.0f, a.mask[(int64_t)b * a.mask_bs + j])) : 0.0f;
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_S = tmem, tmem_O = tmem + 128;

  if (warp == 0) {
    if (lane == 0) {
      tc::tma_prefetch_desc(&tmQ);
      tc::tma_prefetch_desc(&tmK);
      tc::tma_prefetch_desc(&tmV);
      tc::mbar_expect_tx(bar_qk, 2 * TILE_QKV);
      tc::tma_load_2d(base + OFF_Q, &tmQ, h * HD, row0, bar_qk);
      tc::tma_load_2d(base + OFF_K, &tmK, h * HD, row0, bar_qk);
      if (a.has_prev) {
        tc::mbar_expect_tx(bar_sp, TILE_S);
        tc::tma_load_2d(base + OFF_S, &tmSprev, 0, srow0, bar_sp);
        tc::tma_load_2d(base + OFF_S + TILE_S / 2, &tmSprev, 64, srow0, bar_sp);
      }
      tc::mbar_expect_tx(bar_v, TILE_QKV);
      tc::tma_load_2d(base + OFF_V, &tmV, h * HD, row0, bar_v);
      // pull the tile of the CTA that follows on this SM (about one wave ahead) into L2
      const int nxt = blockIdx.y * gridDim.x + blockIdx.x + a.pf_dist;
      if (a.pf_dist > 0 && nxt < (int)(gridDim.x * gridDim.y)) {
        const int h2 = nxt % gridDim.x, b2 = nxt / gridDim.x;
        if (a.has_prev) {
          tc::tma_prefetch_2d(&tmSprev, 0, (b2 * a.H + h2) * L);
          tc::tma_prefetch_2d(&tmSprev, 64, (b2 * a.H + h2) * L);
        }
        tc::tma_prefetch_2d(&tmQ, h2 * HD, b2 * L);
        tc::tma_prefetch_2d(&tmK, h2 * HD, b2 * L);
        tc::tma_prefetch_2d(&tmV, h2 * HD, b2 * L);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---- S = Q K^T ----
      tc::mbar_wait(bar_qk, 0);
      tc::tc_fence_after();
#pragma unroll
      for (int k = 0; k < HD / 16; ++k) {
        const uint64_t ad = tc::smem_desc_sw128(base + OFF_Q + k * 32, 16, 1024);
        const uint64_t bd = tc::smem_desc_sw128(base + OFF_K + k * 32, 16, 1024);
        tc::umma_bf16(tmem_S, ad, bd, a.idesc_qk, k > 0 ? 1u : 0u);
      }
      tc::umma_commit(bar_s);
      // ---- O = P V ----
      tc::mbar_wait(bar_p, 0);
      tc::mbar_wait(bar_v, 0);
      tc::tc_fence_after();
#pragma unroll
      for (int k = 0; k < L / 16; ++k) {
        const uint64_t ad =
            tc::smem_desc_sw128(base + OFF_P + (k >> 2) * (TILE_S / 2) + (k & 3) * 32, 16, 1024);
        const uint64_t bd = tc::smem_desc_sw128(base + OFF_V + k * 2048, TILE_QKV, 1024);
        tc::umma_bf16(tmem_O, ad, bd, a.idesc_pv, k > 0 ? 1u : 0u);
      }
      tc::umma_commit(bar_o);
    }
  } else {
    // ================= softmax / epilogue: thread = query row =================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const float cval = (a.has_prev && a.c) ? a.c[0] : 0.f;
    tc::mbar_wait(bar_s, 0);
    tc::tc_fence_after();
    if (a.has_prev) tc::mbar_wait(bar_sp, 0);
    // pass 1: finish the scores, write them (bf16) into the S tile, track the row max
    float mx = -INFINITY;
#pragma unroll 1
    for (int ch = 0; ch < 4; ++ch) {
      uint32_t r[32];
      tc::tmem_ld32(tmem_S + lane_off + ch * 32, r);
      tc::tmem_ld_wait();
      const uint32_t half = base + OFF_S + (ch >> 1) * (TILE_S / 2);
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const uint32_t addr = half + tc::sw128_offset(row, (ch & 1) * 4 + q4);
        uint32_t pv[4] = {0u, 0u, 0u, 0u};
        if (a.has_prev)
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(pv[0]), "=r"(pv[1]), "=r"(pv[2]), "=r"(pv[3])
                       : "r"(addr));
        uint32_t out[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = ch * 32 + q4 * 8 + e * 2;
          const float2 mk = *reinterpret_cast<const float2*>(mask_s + j);
          float s0 = __uint_as_float(r[q4 * 8 + e * 2]) * 0.125f;   // == / sqrt(64), exact
          float s1 = __uint_as_float(r[q4 * 8 + e * 2 + 1]) * 0.125f;   // == / sqrt(64), exact
          if (a.has_prev) {
            s0 = __fadd_rn(s0, __fmul_rn(cval, bf16_lo(pv[e])));
            s1 = __fadd_rn(s1, __fmul_rn(cval, bf16_hi(pv[e])));
          }
          s0 = __fsub_rn(s0, mk.x);
          s1 = __fsub_rn(s1, mk.y);
          out[e] = pack_bf16(s0, s1);
          mx = fmaxf(mx, fmaxf(bf16_lo(out[e]), bf16_hi(out[e])));
        }
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(out[0]),
                     "r"(out[1]), "r"(out[2]), "r"(out[3])
                     : "memory");
      }
    }
    // pass 2: e = exp(s - max) rounded to bf16 -> P tile (unnormalised); sum of the rounded values
    float sum = 0.f;
    const float kLog2e = 1.4426950408889634f;
#pragma unroll 1
    for (int c16 = 0; c16 < 16; ++c16) {
      const uint32_t off = (c16 >> 3) * (TILE_S / 2) + tc::sw128_offset(row, c16 & 7);
      uint32_t sv[4], pe[4];
      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(sv[0]), "=r"(sv[1]), "=r"(sv[2]), "=r"(sv[3])
                   : "r"(base + OFF_S + off));
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float e0 = tc::ex2((bf16_lo(sv[e]) - mx) * kLog2e);
        const float e1 = tc::ex2((bf16_hi(sv[e]) - mx) * kLog2e);
        pe[e] = pack_bf16(e0, e1);
        sum += bf16_lo(pe[e]) + bf16_hi(pe[e]);
      }
      asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(base + OFF_P + off), "r"(pe[0]),
                   "r"(pe[1]), "r"(pe[2]), "r"(pe[3])
                   : "memory");
    }
    float* st2 = a.stat + 2 * ((int64_t)srow0 + row);
    st2[0] = mx;
    st2[1] = sum;
    // publish S (TMA store) and P (UMMA operand): generic-proxy writes -> async proxy
    tc::fence_proxy_async();
    tc::mbar_arrive(bar_p);
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (a.write_s && threadIdx.x == 64) {
      tc::tma_store_2d(&tmSout, base + OFF_S, 0, srow0);
      tc::tma_store_2d(&tmSout, base + OFF_S + TILE_S / 2, 64, srow0);
      tc::tma_store_commit();
    }
    // O epilogue
    tc::mbar_wait(bar_o, 0);
    tc::tc_fence_after();
    const float inv = 1.0f / sum;
#pragma unroll 1
    for (int ch = 0; ch < 2; ++ch) {
      uint32_t r[32];
      tc::tmem_ld32(tmem_O + lane_off + ch * 32, r);
      tc::tmem_ld_wait();
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        uint32_t out[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
          out[e] = pack_bf16(__uint_as_float(r[q4 * 8 + e * 2]) * inv,
                             __uint_as_float(r[q4 * 8 + e * 2 + 1]) * inv);
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(
                         base + OFF_O + tc::sw128_offset(row, ch * 4 + q4)),
                     "r"(out[0]), "r"(out[1]), "r"(out[2]), "r"(out[3])
                     : "memory");
      }
    }
    tc::fence_proxy_async();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (threadIdx.x == 64) {
      tc::tma_store_2d(&tmO, base + OFF_O, h * HD, row0);
      tc::tma_store_commit();
      tc::tma_store_wait_read();   // shared memory must outlive the bulk stores' reads
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem, 256);
  }
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

bool make2d(CUtensorMap* tm, const void* base, uint64_t cols, uint64_t rows, uint64_t ld_elems,
            uint32_t box_rows) {
  const uint64_t dims[2] = {cols, rows};
  const uint64_t str[1] = {ld_elems * 2};
  const uint32_t box[2] = {64, box_rows};
  return mm_make_tmap_bf16(tm, base, 2, dims, str, box);
}

}  // namespace

bool resattn_tc_supported(int64_t Lq, int64_t Lk, int64_t hd, int64_t ldq, int64_t ldk,
                          int64_t ldv, int64_t ldo) {
  return Lq == L && Lk == L && hd == HD && ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 &&
         ldo % 8 == 0;
}

int resattn_fwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                   int64_t ldv, const float* mask, int64_t mask_bs, const void* s_prev,
                   const float* c, void* s_out, void* o, int64_t ldo, float* lse, int64_t B,
                   int64_t H, int64_t Lq, int64_t Lk, int64_t hd, cudaStream_t st) {
  if (B <= 0 || H <= 0) return MMEMO_OK;
  MM_REQUIRE(q && k && v && o && lse);
  if (!aligned16(q) || !aligned16(k) || !aligned16(v) || !aligned16(o) ||
      (s_prev && !aligned16(s_prev)) || (s_out && !aligned16(s_out)))
    return MMEMO_ERR_ARG;
  // (idempotent: set on every call, so concurrent host threads cannot race on a guard variable)
  MM_CUDA_OK(cudaFuncSetAttribute(resattn_fwd_tc_kernel,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_FWD));
  CUtensorMap tmQ, tmK, tmV, tmSp, tmSo, tmO;
  const uint64_t rows = (uint64_t)B * L, srows = (uint64_t)B * H * L, d = (uint64_t)H * HD;
  bool ok = make2d(&tmQ, q, d, rows, ldq, L) && make2d(&tmK, k, d, rows, ldk, L) &&
            make2d(&tmV, v, d, rows, ldv, L) && make2d(&tmO, o, d, rows, ldo, L);
  // the score maps fall back to a valid dummy (the Q map) when the tensor is absent
  ok = ok && (s_prev ? make2d(&tmSp, s_prev, L, srows, L, L) : make2d(&tmSp, q, d, rows, ldq, L));
  ok = ok && (s_out ? make2d(&tmSo, s_out, L, srows, L, L) : make2d(&tmSo, q, d, rows, ldq, L));
  if (!ok) {
    mmemo_set_error("cuTensorMapEncodeTiled failed (resattn_fwd_tc)", __FILE__, __LINE__);
    return MMEMO_ERR_CUDA;
  }
  FwdArgs a = {};
  a.mask = mask; a.mask_bs = mask_bs; a.c = c; a.stat = lse; a.H = (int)H;
  a.has_prev = s_prev != nullptr; a.write_s = s_out != nullptr;
  a.sqrt_hd = (float)sqrt((double)hd);
  a.pf_dist = 2 * resattn_pf_distance();   // two CTAs per SM
  a.idesc_qk = tc::idesc_bf16(L, L, 0, 0);
  a.idesc_pv = tc::idesc_bf16(L, HD, 0, 1);
  dim3 grid((unsigned)H, (unsigned)B);
  MM_CUDA_OK(mm_launch(resattn_fwd_tc_kernel, grid, dim3(NTHREADS), SMEM_FWD, st, tmQ, tmK, tmV,
                       tmSp, tmSo, tmO, a));
  return MMEMO_OK;
}

// =================================================================================================
// Backward (kernel a'): one CTA per (batch, head), five contractions on tcgen05
//   [S_acc = Q K^T                       only when S was not stored (last layer of a chain)]
//   dP  = dO V^T                         (128 x 128 x 64)
//   regs: p = exp(s - max)/sum, D = rowsum(p*dP), dS = p*(dP - D) + dS_next,
//         dc += sum(dS*S_prev), dS_prev = c*dS        (two passes over TMEM; EIGHT warps: thread =
//         (query row, 64-key half) - two warps per TMEM lane quarter, the halves of D meet in
//         shared memory; with four warps this register phase was ~40 % of a CTA's time)
//   dV  = P^T dO, dK = dS^T Q            (A operands MN-major straight from the P / dS tiles)
//   dQ  = dS K                           (B operand MN-major = K as loaded)
// The S-sized tiles are updated in place in shared memory: S -> P, dS_next -> dS, S_prev -> dS_prev.
// 160 KB shared memory, all 512 TMEM columns: one CTA per SM.
// =================================================================================================
namespace {

constexpr uint32_t B_OFF_Q = 0, B_OFF_K = TILE_QKV, B_OFF_V = 2 * TILE_QKV, B_OFF_DO = 3 * TILE_QKV;
constexpr uint32_t B_OFF_A = 4 * TILE_QKV;             // S_in  -> P
constexpr uint32_t B_OFF_B = B_OFF_A + TILE_S;         // dS_next -> dS
constexpr uint32_t B_OFF_C = B_OFF_B + TILE_S;         // S_prev -> dS_prev
constexpr uint32_t B_OFF_MASK = B_OFF_C + TILE_S;
constexpr uint32_t B_OFF_BAR = B_OFF_MASK + 512;
constexpr uint32_t B_OFF_DSUM = B_OFF_BAR + 128;        // D halves: [2][128] floats
constexpr uint32_t SMEM_BWD = B_OFF_DSUM + 1024 + 1024;
constexpr int NTHREADS_BWD = 64 + 256;                 // TMA warp, MMA warp, eight compute warps

struct BwdArgs {
  const float* mask;
  int64_t mask_bs;
  const float* c;
  const float* stat;
  float* dc;
  int H;
  int has_s, has_prev, has_dsn, write_dsp;
  int pf_dist;
  float sqrt_hd;
  uint32_t idesc_nn128, idesc_tt64, idesc_nt64;
};

__device__ __forceinline__ void lds128(uint32_t addr, uint32_t* v) {
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(addr));
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint32_t* v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3])
               : "memory");
}

__global__ void __launch_bounds__(NTHREADS_BWD, 1)
resattn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                      const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmSprev,
                      const __grid_constant__ CUtensorMap tmDSn,
                      const __grid_constant__ CUtensorMap tmDSp,
                      const __grid_constant__ CUtensorMap tmDQ, const __grid_constant__ CUtensorMap tmDK,
                      const __grid_constant__ CUtensorMap tmDV, const BwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = tc::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const uint32_t bar_in = base + B_OFF_BAR, bar_s = bar_in + 8, bar_sp = bar_in + 16,
                 bar_dsn = bar_in + 24, bar_mm1 = bar_in + 32, bar_p2 = bar_in + 40,
                 bar_mm2 = bar_in + 48;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + B_OFF_BAR + 64);
  float* mask_s = reinterpret_cast<float*>(gbase + B_OFF_MASK);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, b = blockIdx.y;
  const int row0 = b * L;
  const int srow0 = (b * a.H + h) * L;

  if (threadIdx.x == 0) {
    tc::mbar_init(bar_in, 1);
    tc::mbar_init(bar_s, 1);
    tc::mbar_init(bar_sp, 1);
    tc::mbar_init(bar_dsn, 1);
    tc::mbar_init(bar_mm1, 1);
    tc::mbar_init(bar_p2, 256);
    tc::mbar_init(bar_mm2, 1);
    tc::fence_barrier_init();
  }
  pdl_trigger();
  if (warp == 1) {        // TMEM allocation overlaps the previous kernel's drain
    tc::tmem_alloc(base + B_OFF_BAR + 64, 512);
    tc::tmem_relinquish();
  }
  pdl_wait();             // before the first global access (mask row, TMA loads)
  if (threadIdx.x >= 64 && threadIdx.x < 64 + L) {
    const int j = threadIdx.x - 64;
    // additive mask term 1e8*(1-m), computed once per key (same fp32 ops as the reference)
    mask_s[j] = a.mask ? __fmul_rn(1.0e8f, __fsub_rn(1.0f, a.mask[(int64_t)b * a.mask_bs + j])) : 0.0f;
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tm_S = tmem, tm_dP = tmem + 128, tm_dV = tmem + 256, tm_dK = tmem + 320,
                 tm_dQ = tmem + 384;

  if (warp == 0) {
    if (lane == 0) {
      tc::mbar_expect_tx(bar_in, 4 * TILE_QKV);
      tc::tma_load_2d(base + B_OFF_DO, &tmDO, h * HD, row0, bar_in);
      tc::tma_load_2d(base + B_OFF_V, &tmV, h * HD, row0, bar_in);
      tc::tma_load_2d(base + B_OFF_Q, &tmQ, h * HD, row0, bar_in);
      tc::tma_load_2d(base + B_OFF_K, &tmK, h * HD, row0, bar_in);
      if (a.has_s) {
        tc::mbar_expect_tx(bar_s, TILE_S);
        tc::tma_load_2d(base + B_OFF_A, &tmS, 0, srow0, bar_s);
        tc::tma_load_2d(base + B_OFF_A + TILE_S / 2, &tmS, 64, srow0, bar_s);
      }
      if (a.has_prev) {
        tc::mbar_expect_tx(bar_sp, TILE_S);
        tc::tma_load_2d(base + B_OFF_C, &tmSprev, 0, srow0, bar_sp);
        tc::tma_load_2d(base + B_OFF_C + TILE_S / 2, &tmSprev, 64, srow0, bar_sp);
      }
      if (a.has_dsn) {
        tc::mbar_expect_tx(bar_dsn, TILE_S);
        tc::tma_load_2d(base + B_OFF_B, &tmDSn, 0, srow0, bar_dsn);
        tc::tma_load_2d(base + B_OFF_B + TILE_S / 2, &tmDSn, 64, srow0, bar_dsn);
      }
      const int nxt = blockIdx.y * gridDim.x + blockIdx.x + a.pf_dist;
      if (a.pf_dist > 0 && nxt < (int)(gridDim.x * gridDim.y)) {
        const int h2 = nxt % gridDim.x, b2 = nxt / gridDim.x, sr2 = (b2 * a.H + h2) * L;
        tc::tma_prefetch_2d(&tmDO, h2 * HD, b2 * L);
        tc::tma_prefetch_2d(&tmV, h2 * HD, b2 * L);
        tc::tma_prefetch_2d(&tmQ, h2 * HD, b2 * L);
        tc::tma_prefetch_2d(&tmK, h2 * HD, b2 * L);
        if (a.has_s) {
          tc::tma_prefetch_2d(&tmS, 0, sr2);
          tc::tma_prefetch_2d(&tmS, 64, sr2);
        }
        if (a.has_prev) {
          tc::tma_prefetch_2d(&tmSprev, 0, sr2);
          tc::tma_prefetch_2d(&tmSprev, 64, sr2);
        }
        if (a.has_dsn) {
          tc::tma_prefetch_2d(&tmDSn, 0, sr2);
          tc::tma_prefetch_2d(&tmDSn, 64, sr2);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      tc::mbar_wait(bar_in, 0);
      tc::tc_fence_after();
      if (!a.has_s) {
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          tc::umma_bf16(tm_S, tc::smem_desc_sw128(base + B_OFF_Q + k * 32, 16, 1024),
                        tc::smem_desc_sw128(base + B_OFF_K + k * 32, 16, 1024), a.idesc_nn128,
                        k > 0 ? 1u : 0u);
      }
#pragma unroll
      for (int k = 0; k < HD / 16; ++k)
        tc::umma_bf16(tm_dP, tc::smem_desc_sw128(base + B_OFF_DO + k * 32, 16, 1024),
                      tc::smem_desc_sw128(base + B_OFF_V + k * 32, 16, 1024), a.idesc_nn128,
                      k > 0 ? 1u : 0u);
      tc::umma_commit(bar_mm1);
      // second round: operands P (tile A) and dS (tile B) written by the softmax threads
      tc::mbar_wait(bar_p2, 0);
      tc::tc_fence_after();
#pragma unroll
      for (int k = 0; k < L / 16; ++k) {   // K = query rows, 16 per step = 2048 B in every tile
        const uint64_t pT = tc::smem_desc_sw128(base + B_OFF_A + k * 2048, TILE_S / 2, 1024);
        const uint64_t dsT = tc::smem_desc_sw128(base + B_OFF_B + k * 2048, TILE_S / 2, 1024);
        const uint64_t dOm = tc::smem_desc_sw128(base + B_OFF_DO + k * 2048, TILE_QKV, 1024);
        const uint64_t Qm = tc::smem_desc_sw128(base + B_OFF_Q + k * 2048, TILE_QKV, 1024);
        tc::umma_bf16(tm_dV, pT, dOm, a.idesc_tt64, k > 0 ? 1u : 0u);
        tc::umma_bf16(tm_dK, dsT, Qm, a.idesc_tt64, k > 0 ? 1u : 0u);
      }
#pragma unroll
      for (int k = 0; k < L / 16; ++k) {   // K = keys: dS K-major (two 64-key halves), K MN-major
        const uint64_t dsK = tc::smem_desc_sw128(
            base + B_OFF_B + (k >> 2) * (TILE_S / 2) + (k & 3) * 32, 16, 1024);
        const uint64_t Km = tc::smem_desc_sw128(base + B_OFF_K + k * 2048, TILE_QKV, 1024);
        tc::umma_bf16(tm_dQ, dsK, Km, a.idesc_nt64, k > 0 ? 1u : 0u);
      }
      tc::umma_commit(bar_mm2);
    }
  } else {
    const int quarter = warp & 3;
    const int khalf = (warp - 2) >> 2;              // which 64 keys (pass A / B), which 32 head
    const int row = quarter * 32 + lane;            // columns (epilogue) this thread handles
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    float* dsum = reinterpret_cast<float*>(gbase + B_OFF_DSUM);
    const float cval = (a.has_prev && a.c) ? a.c[0] : 0.f;
    const float* st2 = a.stat + 2 * ((int64_t)srow0 + row);
    const float mx = st2[0], inv = 1.0f / st2[1];
    const float kLog2e = 1.4426950408889634f;
    tc::mbar_wait(bar_mm1, 0);
    tc::tc_fence_after();
    if (a.has_s) tc::mbar_wait(bar_s, 0);
    if (a.has_prev) tc::mbar_wait(bar_sp, 0);
    if (a.has_dsn) tc::mbar_wait(bar_dsn, 0);
    // ---- pass A: P (bf16) into tile A, D = sum_j p*dP ------------------------------------------
    float D = 0.f;
#pragma unroll 1
    for (int ch = 2 * khalf; ch < 2 * khalf + 2; ++ch) {
      uint32_t dp[32], sa[32];
      tc::tmem_ld32(tm_dP + lane_off + ch * 32, dp);
      if (!a.has_s) tc::tmem_ld32(tm_S + lane_off + ch * 32, sa);
      tc::tmem_ld_wait();
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const uint32_t off = (ch >> 1) * (TILE_S / 2) + tc::sw128_offset(row, (ch & 1) * 4 + q4);
        uint32_t sv[4] = {0u, 0u, 0u, 0u}, pv[4] = {0u, 0u, 0u, 0u}, out[4];
        if (a.has_s) lds128(base + B_OFF_A + off, sv);
        else if (a.has_prev) lds128(base + B_OFF_C + off, pv);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i0 = q4 * 8 + e * 2, j = ch * 32 + i0;
          const float2 mk = *reinterpret_cast<const float2*>(mask_s + j);
          float s0, s1;
          if (a.has_s) {
            s0 = bf16_lo(sv[e]);
            s1 = bf16_hi(sv[e]);
          } else {
            s0 = __uint_as_float(sa[i0]) * 0.125f;   // == / sqrt(64), exact
            s1 = __uint_as_float(sa[i0 + 1]) * 0.125f;   // == / sqrt(64), exact
            if (a.has_prev) {
              s0 = __fadd_rn(s0, __fmul_rn(cval, bf16_lo(pv[e])));
              s1 = __fadd_rn(s1, __fmul_rn(cval, bf16_hi(pv[e])));
            }
            s0 = __fsub_rn(s0, mk.x);
            s1 = __fsub_rn(s1, mk.y);
            const uint32_t rr = pack_bf16(s0, s1);   // the forward softmax saw bf16-rounded scores
            s0 = bf16_lo(rr);
            s1 = bf16_hi(rr);
          }
          const float p0 = tc::ex2((s0 - mx) * kLog2e) * inv, p1 = tc::ex2((s1 - mx) * kLog2e) * inv;
          out[e] = pack_bf16(p0, p1);
          D = fmaf(bf16_lo(out[e]), __uint_as_float(dp[i0]), D);
          D = fmaf(bf16_hi(out[e]), __uint_as_float(dp[i0 + 1]), D);
        }
        sts128(base + B_OFF_A + off, out);
      }
    }
    // the two key halves of a row exchange their partial D
    dsum[khalf * 128 + row] = D;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    D += dsum[(khalf ^ 1) * 128 + row];
    // ---- pass B: dS into tile B, dc, dS_prev into tile C -----------------------------------------
    float dc_part = 0.f;
#pragma unroll 1
    for (int ch = 2 * khalf; ch < 2 * khalf + 2; ++ch) {
      uint32_t dp[32];
      tc::tmem_ld32(tm_dP + lane_off + ch * 32, dp);
      tc::tmem_ld_wait();
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const uint32_t off = (ch >> 1) * (TILE_S / 2) + tc::sw128_offset(row, (ch & 1) * 4 + q4);
        uint32_t pp[4], dn[4] = {0u, 0u, 0u, 0u}, sp[4] = {0u, 0u, 0u, 0u}, out[4], outp[4];
        lds128(base + B_OFF_A + off, pp);
        if (a.has_dsn) lds128(base + B_OFF_B + off, dn);
        if (a.has_prev) lds128(base + B_OFF_C + off, sp);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i0 = q4 * 8 + e * 2;
          float d0 = bf16_lo(pp[e]) * (__uint_as_float(dp[i0]) - D);
          float d1 = bf16_hi(pp[e]) * (__uint_as_float(dp[i0 + 1]) - D);
          if (a.has_dsn) {
            d0 += bf16_lo(dn[e]);
            d1 += bf16_hi(dn[e]);
          }
          out[e] = pack_bf16(d0, d1);
          if (a.has_prev) {
            dc_part = fmaf(d0, bf16_lo(sp[e]), dc_part);
            dc_part = fmaf(d1, bf16_hi(sp[e]), dc_part);
            outp[e] = pack_bf16(cval * d0, cval * d1);
          }
        }
        sts128(base + B_OFF_B + off, out);
        if (a.has_prev && a.write_dsp) sts128(base + B_OFF_C + off, outp);
      }
    }
    tc::fence_proxy_async();
    tc::mbar_arrive(bar_p2);
    if (a.has_prev && a.dc) {
      dc_part = warp_sum(dc_part);
      if (lane == 0) atomicAdd(a.dc, dc_part);
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (a.has_prev && a.write_dsp && threadIdx.x == 64) {
      tc::tma_store_2d(&tmDSp, base + B_OFF_C, 0, srow0);
      tc::tma_store_2d(&tmDSp, base + B_OFF_C + TILE_S / 2, 64, srow0);
      tc::tma_store_commit();
    }
    // ---- epilogue: dV, dK, dQ -> bf16 staging (V, K, Q tiles are dead) -> TMA stores ---------------
    tc::mbar_wait(bar_mm2, 0);
    tc::tc_fence_after();
    const float inv_sqrt = 1.0f * 0.125f;   // == / sqrt(64), exact
#pragma unroll 1
    for (int t = 0; t < 3; ++t) {
      const uint32_t src = t == 0 ? tm_dV : (t == 1 ? tm_dK : tm_dQ);
      const uint32_t dst = base + (t == 0 ? B_OFF_V : (t == 1 ? B_OFF_K : B_OFF_Q));
      const float sc = t == 0 ? 1.0f : inv_sqrt;
      {
        const int ch = khalf;                       // 32 of the 64 head columns per thread
        uint32_t r[32];
        tc::tmem_ld32(src + lane_off + ch * 32, r);
        tc::tmem_ld_wait();
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint32_t out[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            out[e] = pack_bf16(__uint_as_float(r[q4 * 8 + e * 2]) * sc,
                               __uint_as_float(r[q4 * 8 + e * 2 + 1]) * sc);
          sts128(dst + tc::sw128_offset(row, ch * 4 + q4), out);
        }
      }
    }
    tc::fence_proxy_async();
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (threadIdx.x == 64) {
      tc::tma_store_2d(&tmDV, base + B_OFF_V, h * HD, row0);
      tc::tma_store_2d(&tmDK, base + B_OFF_K, h * HD, row0);
      tc::tma_store_2d(&tmDQ, base + B_OFF_Q, h * HD, row0);
      tc::tma_store_commit();
      tc::tma_store_wait_read();
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem, 512);
  }
}

}  // namespace

int resattn_bwd_tc(const void* d_o, int64_t lddo, const void* q, int64_t ldq, const void* k,
                   int64_t ldk, const void* v, int64_t ldv, const float* mask, int64_t mask_bs,
                   const void* s, const void* s_prev, const float* c, const void* ds_next,
                   const float* lse, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv,
                   int64_t lddv, void* ds_prev, float* dc, int64_t B, int64_t H, cudaStream_t st) {
  if (B <= 0 || H <= 0) return MMEMO_OK;
  MM_REQUIRE(d_o && q && k && v && lse && dq && dk && dv);
  const void* ptrs[] = {d_o, q, k, v, dq, dk, dv, s, s_prev, ds_next, ds_prev};
  for (const void* p : ptrs)
    if (p && !aligned16(p)) return MMEMO_ERR_ARG;
  MM_CUDA_OK(cudaFuncSetAttribute(resattn_bwd_tc_kernel,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BWD));
  CUtensorMap tmQ, tmK, tmV, tmDO, tmS, tmSp, tmDSn, tmDSp, tmDQ, tmDK, tmDV;
  const uint64_t rows = (uint64_t)B * L, srows = (uint64_t)B * H * L, d = (uint64_t)H * HD;
  bool ok = make2d(&tmQ, q, d, rows, ldq, L) && make2d(&tmK, k, d, rows, ldk, L) &&
            make2d(&tmV, v, d, rows, ldv, L) && make2d(&tmDO, d_o, d, rows, lddo, L) &&
            make2d(&tmDQ, dq, d, rows, lddq, L) && make2d(&tmDK, dk, d, rows, lddk, L) &&
            make2d(&tmDV, dv, d, rows, lddv, L);
  auto score_map = [&](CUtensorMap* tm, const void* p) {
    return p ? make2d(tm, p, L, srows, L, L) : make2d(tm, q, d, rows, ldq, L);
  };
  ok = ok && score_map(&tmS, s) && score_map(&tmSp, s_prev) && score_map(&tmDSn, ds_next) &&
       score_map(&tmDSp, ds_prev);
  if (!ok) {
    mmemo_set_error("cuTensorMapEncodeTiled failed (resattn_bwd_tc)", __FILE__, __LINE__);
    return MMEMO_ERR_CUDA;
  }
  BwdArgs a = {};
  a.mask = mask; a.mask_bs = mask_bs; a.c = c; a.stat = lse; a.dc = dc; a.H = (int)H;
  a.has_s = s != nullptr; a.has_prev = s_prev != nullptr; a.has_dsn = ds_next != nullptr;
  a.write_dsp = ds_prev != nullptr;
  a.sqrt_hd = 8.0f;
  a.pf_dist = resattn_pf_distance();
  a.idesc_nn128 = tc::idesc_bf16(L, L, 0, 0);
  a.idesc_tt64 = tc::idesc_bf16(L, HD, 1, 1);
  a.idesc_nt64 = tc::idesc_bf16(L, HD, 0, 1);
  dim3 grid((unsigned)H, (unsigned)B);
  MM_CUDA_OK(mm_launch(resattn_bwd_tc_kernel, grid, dim3(NTHREADS_BWD), SMEM_BWD, st, tmQ, tmK, tmV,
                       tmDO, tmS, tmSp, tmDSn, tmDSp, tmDQ, tmDK, tmDV, a));
  return MMEMO_OK;
}
