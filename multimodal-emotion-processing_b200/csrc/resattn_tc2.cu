// Residual-attention core on tcgen05 / TMA for the long-sequence shapes (bf16, head_dim 64,
// Lk = 256, Lq = 128 or 256): the text-encoder configuration (rencecps, seq 256).  Same arithmetic
// as resattn_tc.cu (reference op order for the scores, bf16-rounded S / P, (max, sum) row
// statistics), tiled over 128-row query tiles and 128-key tiles.
//
//   forward  : one CTA per (batch, head, query tile).  S_acc = Q K^T is ONE 128 x 256 x 64 UMMA
//              chain into 256 TMEM columns; eight softmax warps (two per TMEM lane quarter, each
//              owning 128 keys of a row) finish the scores, exchange the row max / sum through
//              shared memory, write S (TMA store) and P; O = P V accumulates over all 256 keys.
//   backward : one CTA per (batch, head); key tiles outer, query tiles inner.  dV / dK of the
//              current key tile and dQ of BOTH query tiles stay in TMEM across the loop (all 512
//              columns); the row term D = rowsum(dO * O) is formed once from the saved forward
//              output, so each (query tile, key tile) step is a single pass over dP.
//
// 320 threads: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-9 = softmax /
// epilogue.  One CTA per SM (shared memory and TMEM), so the second softmax warpgroup is what
// keeps the SM's schedulers busy.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "resattn.h"
#include "tc_common.cuh"

namespace {

constexpr int TL = 128, HD = 64;
constexpr uint32_t T16 = TL * HD * 2;          // 16 KB: a 128 x 64 bf16 tile (Q/K/V/O or 64 keys of S)
constexpr int NTHREADS = 320, NSOFT = 256;
constexpr int NK = 2;                          // key tiles in the forward kernel (Lk = 256)

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
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
__device__ __forceinline__ void soft_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
constexpr uint32_t F_OFF_Q = 0, F_OFF_K = T16, F_OFF_V = F_OFF_K + NK * T16,
                   F_OFF_S = F_OFF_V + NK * T16,            // 2*NK tiles of 64 keys
                   F_OFF_P = F_OFF_S + 2 * NK * T16,
                   F_OFF_MASK = F_OFF_P + 2 * NK * T16,      // NK*128 floats
                   F_OFF_RED = F_OFF_MASK + NK * 512,        // [2][128] max, [2][128] sum
                   F_OFF_BAR = F_OFF_RED + 2048;
constexpr uint32_t F_OFF_O = F_OFF_Q;                        // O staging reuses the Q tile
constexpr uint32_t SMEM_FWD2 = F_OFF_BAR + 128 + 1024;

struct Fwd2Args {
  const float* mask;
  int64_t mask_bs;
  const float* c;
  float* stat;   // (B, H, Lq, 2)
  int H, nq;
  int has_prev, write_s;
  int pf_dist;   // prefetch the tile of CTA (linear id + pf_dist) into L2; 0 = off
  uint32_t idesc_qk, idesc_pv;
};

__global__ void __launch_bounds__(NTHREADS, 1)
resattn_fwd_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV,
                       const __grid_constant__ CUtensorMap tmSprev,
                       const __grid_constant__ CUtensorMap tmSout,
                       const __grid_constant__ CUtensorMap tmO, const Fwd2Args a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = tc::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const uint32_t bar_qk = base + F_OFF_BAR, bar_v = bar_qk + 8, bar_sp = bar_qk + 16,
                 bar_s = bar_qk + 24, bar_p = bar_qk + 32, bar_o = bar_qk + 40;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + F_OFF_BAR + 64);
  float* mask_s = reinterpret_cast<float*>(gbase + F_OFF_MASK);
  float* red = reinterpret_cast<float*>(gbase + F_OFF_RED);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, b = blockIdx.y / a.nq, qt = blockIdx.y - b * a.nq;
  const int Lq = a.nq * TL, Lk = NK * TL;
  const int qrow0 = b * Lq + qt * TL;                 // first row in the (B*Lq, ld) views
  const int krow0 = b * Lk;
  const int srow0 = (b * a.H + h) * Lq + qt * TL;     // first row in the (B*H*Lq, Lk) score views

  if (threadIdx.x == 0) {
    tc::mbar_init(bar_qk, 1);
    tc::mbar_init(bar_v, 1);
    tc::mbar_init(bar_sp, 1);
    tc::mbar_init(bar_s, 1);
    tc::mbar_init(bar_p, NSOFT);
    tc::mbar_init(bar_o, 1);
    tc::fence_barrier_init();
  }
  pdl_trigger();
  if (warp == 1) {        // TMEM allocation overlaps the previous kernel's drain
    tc::tmem_alloc(base + F_OFF_BAR + 64, 512);
    tc::tmem_relinquish();
  }
  pdl_wait();             // before the first global access (mask row, TMA loads)
  if (threadIdx.x >= 64) {
    const int j = threadIdx.x - 64;   // 256 softmax threads = 256 keys
    mask_s[j] = a.mask ? __fmul_rn(1.0e8f, __fsub_rn(1.0f, a.mask[(int64_t)b * a.mask_bs + j])) : 0.0f;
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_S = tmem, tmem_O = tmem + NK * TL;

  if (warp == 0) {
    if (lane == 0) {
      tc::tma_prefetch_desc(&tmQ);
      tc::tma_prefetch_desc(&tmK);
      tc::tma_prefetch_desc(&tmV);
      tc::mbar_expect_tx(bar_qk, (1 + NK) * T16);
      tc::tma_load_2d(base + F_OFF_Q, &tmQ, h * HD, qrow0, bar_qk);
      tc::tma_load_2d(base + F_OFF_K, &tmK, h * HD, krow0, bar_qk);       // 256-row box
      if (a.has_prev) {
        tc::mbar_expect_tx(bar_sp, 2 * NK * T16);
#pragma unroll
        for (int i = 0; i < 2 * NK; ++i)
          tc::tma_load_2d(base + F_OFF_S + i * T16, &tmSprev, i * 64, srow0, bar_sp);
      }
      tc::mbar_expect_tx(bar_v, NK * T16);
      tc::tma_load_2d(base + F_OFF_V, &tmV, h * HD, krow0, bar_v);
      // One CTA per SM: nothing hides this CTA's own load latency, so pull the tile of the CTA
      // that will follow on this SM (about one wave ahead) into L2 while this one computes.
      const int nxt = blockIdx.y * gridDim.x + blockIdx.x + a.pf_dist;
      if (a.pf_dist > 0 && nxt < (int)(gridDim.x * gridDim.y)) {
        const int h2 = nxt % gridDim.x, y2 = nxt / gridDim.x, b2 = y2 / a.nq, qt2 = y2 - b2 * a.nq;
        if (a.has_prev) {
          const int sr2 = (b2 * a.H + h2) * Lq + qt2 * TL;
#pragma unroll
          for (int i = 0; i < 2 * NK; ++i) tc::tma_prefetch_2d(&tmSprev, i * 64, sr2);
        }
        tc::tma_prefetch_2d(&tmQ, h2 * HD, b2 * Lq + qt2 * TL);
        tc::tma_prefetch_2d(&tmK, h2 * HD, b2 * Lk);
        tc::tma_prefetch_2d(&tmV, h2 * HD, b2 * Lk);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      tc::mbar_wait(bar_qk, 0);
      tc::tc_fence_after();
#pragma unroll
      for (int k = 0; k < HD / 16; ++k) {
        const uint64_t ad = tc::smem_desc_sw128(base + F_OFF_Q + k * 32, 16, 1024);
        const uint64_t bd = tc::smem_desc_sw128(base + F_OFF_K + k * 32, 16, 1024);
        tc::umma_bf16(tmem_S, ad, bd, a.idesc_qk, k > 0 ? 1u : 0u);
      }
      tc::umma_commit(bar_s);
      tc::mbar_wait(bar_p, 0);
      tc::mbar_wait(bar_v, 0);
      tc::tc_fence_after();
#pragma unroll
      for (int k = 0; k < NK * TL / 16; ++k) {
        const uint64_t ad = tc::smem_desc_sw128(base + F_OFF_P + (k >> 2) * T16 + (k & 3) * 32, 16, 1024);
        const uint64_t bd = tc::smem_desc_sw128(base + F_OFF_V + k * 2048, T16, 1024);
        tc::umma_bf16(tmem_O, ad, bd, a.idesc_pv, k > 0 ? 1u : 0u);
      }
      tc::umma_commit(bar_o);
    }
  } else {
    // ================= softmax / epilogue: thread = (query row, 128-key half) ==================
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may read
    const int cg = (warp - 2) >> 2;               // which 128 keys of the row
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const float cval = (a.has_prev && a.c) ? a.c[0] : 0.f;
    tc::mbar_wait(bar_s, 0);
    tc::tc_fence_after();
    if (a.has_prev) tc::mbar_wait(bar_sp, 0);
    float mx = -INFINITY;
#pragma unroll 1
    for (int ch = 0; ch < 4; ++ch) {
      uint32_t r[32];
      const int col = cg * TL + ch * 32;
      tc::tmem_ld32(tmem_S + lane_off + col, r);
      tc::tmem_ld_wait();
      const uint32_t tile = base + F_OFF_S + (uint32_t)(col >> 6) * T16;
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const uint32_t addr = tile + tc::sw128_offset(row, (ch & 1) * 4 + q4);
        uint32_t pv[4] = {0u, 0u, 0u, 0u};
        if (a.has_prev) lds128(addr, pv);
        uint32_t out[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = col + q4 * 8 + e * 2;
          const float2 mk = *reinterpret_cast<const float2*>(mask_s + j);
          float s0 = __uint_as_float(r[q4 * 8 + e * 2]) * 0.125f;       // == / sqrt(64), exact
          float s1 = __uint_as_float(r[q4 * 8 + e * 2 + 1]) * 0.125f;
          if (a.has_prev) {
            s0 = __fadd_rn(s0, __fmul_rn(cval, bf16_lo(pv[e])));
            s1 = __fadd_rn(s1, __fmul_rn(cval, bf16_hi(pv[e])));
          }
          s0 = __fsub_rn(s0, mk.x);
          s1 = __fsub_rn(s1, mk.y);
          out[e] = pack_bf16(s0, s1);
          mx = fmaxf(mx, fmaxf(bf16_lo(out[e]), bf16_hi(out[e])));
        }
        sts128(addr, out);
      }
    }
    red[cg * TL + row] = mx;
    soft_sync();
    mx = fmaxf(red[row], red[TL + row]);
    float sum = 0.f;
    const float kLog2e = 1.4426950408889634f;
#pragma unroll 1
    for (int c16 = 0; c16 < 16; ++c16) {
      const uint32_t off = (uint32_t)(cg * 2 + (c16 >> 3)) * T16 + tc::sw128_offset(row, c16 & 7);
      uint32_t sv[4], pe[4];
      lds128(base + F_OFF_S + off, sv);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float e0 = tc::ex2((bf16_lo(sv[e]) - mx) * kLog2e);
        const float e1 = tc::ex2((bf16_hi(sv[e]) - mx) * kLog2e);
        pe[e] = pack_bf16(e0, e1);
        sum += bf16_lo(pe[e]) + bf16_hi(pe[e]);
      }
      sts128(base + F_OFF_P + off, pe);
    }
    red[2 * TL + cg * TL + row] = sum;
    tc::fence_proxy_async();
    tc::mbar_arrive(bar_p);
    soft_sync();
    sum = red[2 * TL + row] + red[3 * TL + row];
    if (cg == 0) {
      float* st2 = a.stat + 2 * ((int64_t)srow0 + row);
      st2[0] = mx;
      st2[1] = sum;
    }
    if (a.write_s && threadIdx.x == 64) {
#pragma unroll
      for (int i = 0; i < 2 * NK; ++i)
        tc::tma_store_2d(&tmSout, base + F_OFF_S + i * T16, i * 64, srow0);
      tc::tma_store_commit();
    }
    tc::mbar_wait(bar_o, 0);
    tc::tc_fence_after();
    const float inv = 1.0f / sum;
    {
      uint32_t r[32];
      tc::tmem_ld32(tmem_O + lane_off + cg * 32, r);
      tc::tmem_ld_wait();
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        uint32_t out[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
          out[e] = pack_bf16(__uint_as_float(r[q4 * 8 + e * 2]) * inv,
                             __uint_as_float(r[q4 * 8 + e * 2 + 1]) * inv);
        sts128(base + F_OFF_O + tc::sw128_offset(row, cg * 4 + q4), out);
      }
    }
    tc::fence_proxy_async();
    soft_sync();
    if (threadIdx.x == 64) {
      tc::tma_store_2d(&tmO, base + F_OFF_O, h * HD, qrow0);
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

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
constexpr int MAX_NQ = 2, MAX_LK = 512;
constexpr uint32_t TILE_S = 2 * T16;                         // 128 x 128 bf16, two 64-key halves
constexpr uint32_t B_OFF_Q = 0, B_OFF_DO = MAX_NQ * T16, B_OFF_K = 2 * MAX_NQ * T16,
                   B_OFF_V = B_OFF_K + T16, B_OFF_A = B_OFF_V + T16,   // S_in -> P
                   B_OFF_B = B_OFF_A + TILE_S,                         // dS_next -> dS
                   B_OFF_C = B_OFF_B + TILE_S,                         // S_prev -> dS_prev
                   B_OFF_MASK = B_OFF_C + TILE_S,                      // Lk floats
                   B_OFF_RED = B_OFF_MASK + MAX_LK * 4,                // [2][MAX_NQ][128] floats
                   B_OFF_BAR = B_OFF_RED + 2 * MAX_NQ * TL * 4;
constexpr uint32_t SMEM_BWD2 = B_OFF_BAR + 128 + 1024;

struct Bwd2Args {
  const float* mask;
  int64_t mask_bs;
  const float* c;
  const float* stat;
  float* dc;
  const __nv_bfloat16 *d_o, *o;     // for D = rowsum(dO * O)
  int64_t lddo, ldo;
  int H, nq, nk;
  int has_s, has_prev, has_dsn, write_dsp;
  int pf_dist;   // L2 prefetch distance in CTAs for the next (batch, head); 0 = off
  uint32_t idesc_nn128, idesc_tt64, idesc_nt64;
};

__global__ void __launch_bounds__(NTHREADS, 1)
resattn_bwd_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                       const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmSprev,
                       const __grid_constant__ CUtensorMap tmDSn,
                       const __grid_constant__ CUtensorMap tmDSp,
                       const __grid_constant__ CUtensorMap tmDQ, const __grid_constant__ CUtensorMap tmDK,
                       const __grid_constant__ CUtensorMap tmDV, const Bwd2Args a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = tc::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const uint32_t bar_in = base + B_OFF_BAR, bar_kv = bar_in + 8, bar_s = bar_in + 16,
                 bar_sp = bar_in + 24, bar_dsn = bar_in + 32, bar_mm1 = bar_in + 40,
                 bar_p2 = bar_in + 48, bar_mm2 = bar_in + 56, bar_free = bar_in + 64;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + B_OFF_BAR + 96);
  float* mask_s = reinterpret_cast<float*>(gbase + B_OFF_MASK);
  float* red = reinterpret_cast<float*>(gbase + B_OFF_RED);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, b = blockIdx.y;
  const int nq = a.nq, nk = a.nk;
  const int Lq = nq * TL, Lk = nk * TL;
  const int qrow0 = b * Lq, krow0 = b * Lk;
  const int srow0 = (b * a.H + h) * Lq;

  if (threadIdx.x == 0) {
    tc::mbar_init(bar_in, 1);
    tc::mbar_init(bar_kv, 1);
    tc::mbar_init(bar_s, 1);
    tc::mbar_init(bar_sp, 1);
    tc::mbar_init(bar_dsn, 1);
    tc::mbar_init(bar_mm1, 1);
    tc::mbar_init(bar_p2, NSOFT);
    tc::mbar_init(bar_mm2, 1);
    tc::mbar_init(bar_free, 1);
    tc::fence_barrier_init();
  }
  pdl_trigger();
  if (warp == 1) {        // TMEM allocation overlaps the previous kernel's drain
    tc::tmem_alloc(base + B_OFF_BAR + 96, 512);
    tc::tmem_relinquish();
  }
  pdl_wait();             // before the first global access (mask row, TMA loads)
  if (threadIdx.x >= 64) {
    for (int j = threadIdx.x - 64; j < Lk; j += NSOFT)
      mask_s[j] = a.mask ? __fmul_rn(1.0e8f, __fsub_rn(1.0f, a.mask[(int64_t)b * a.mask_bs + j])) : 0.0f;
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tm_S = tmem, tm_dP = tmem + 128, tm_dV = tmem + 256, tm_dK = tmem + 320,
                 tm_dQ = tmem + 384;   // + 64 * qt

  if (warp == 0) {
    if (lane == 0) {
      tc::mbar_expect_tx(bar_in, 2 * nq * T16);
      for (int qt = 0; qt < nq; ++qt) {
        tc::tma_load_2d(base + B_OFF_DO + qt * T16, &tmDO, h * HD, qrow0 + qt * TL, bar_in);
        tc::tma_load_2d(base + B_OFF_Q + qt * T16, &tmQ, h * HD, qrow0 + qt * TL, bar_in);
      }
      int t = 0;
      for (int kt = 0; kt < nk; ++kt) {
        for (int qt = 0; qt < nq; ++qt, ++t) {
          // the previous step's MMAs have consumed K/V/A/B and its bulk stores have read C/K/V
          if (t > 0) tc::mbar_wait(bar_free, (t - 1) & 1);
          if (qt == 0) {
            tc::mbar_expect_tx(bar_kv, 2 * T16);
            tc::tma_load_2d(base + B_OFF_V, &tmV, h * HD, krow0 + kt * TL, bar_kv);
            tc::tma_load_2d(base + B_OFF_K, &tmK, h * HD, krow0 + kt * TL, bar_kv);
          }
          const int sr = srow0 + qt * TL, sc = kt * TL;
          if (a.has_s) {
            tc::mbar_expect_tx(bar_s, TILE_S);
            tc::tma_load_2d(base + B_OFF_A, &tmS, sc, sr, bar_s);
            tc::tma_load_2d(base + B_OFF_A + T16, &tmS, sc + 64, sr, bar_s);
          }
          if (a.has_dsn) {
            tc::mbar_expect_tx(bar_dsn, TILE_S);
            tc::tma_load_2d(base + B_OFF_B, &tmDSn, sc, sr, bar_dsn);
            tc::tma_load_2d(base + B_OFF_B + T16, &tmDSn, sc + 64, sr, bar_dsn);
          }
          if (a.has_prev) {
            tc::mbar_expect_tx(bar_sp, TILE_S);
            tc::tma_load_2d(base + B_OFF_C, &tmSprev, sc, sr, bar_sp);
            tc::tma_load_2d(base + B_OFF_C + T16, &tmSprev, sc + 64, sr, bar_sp);
          }
          // The shared-memory tiles are single-buffered, so the NEXT step's score tiles cannot be
          // loaded yet — but they can be pulled into L2 now, which turns the next step's loads
          // into L2 hits and keeps HBM busy while this step computes.  After the last step the
          // target is the (batch, head) that follows on this SM.
          if (a.pf_dist > 0) {
            int sr2 = -1, sc2 = 0;
            if (qt + 1 < nq) {
              sr2 = sr + TL; sc2 = sc;
            } else if (kt + 1 < nk) {
              sr2 = srow0; sc2 = sc + TL;
              tc::tma_prefetch_2d(&tmV, h * HD, krow0 + (kt + 1) * TL);
              tc::tma_prefetch_2d(&tmK, h * HD, krow0 + (kt + 1) * TL);
            } else {
              const int nxt = blockIdx.y * gridDim.x + blockIdx.x + a.pf_dist;
              if (nxt < (int)(gridDim.x * gridDim.y)) {
                const int h2 = nxt % gridDim.x, b2 = nxt / gridDim.x;
                sr2 = (b2 * a.H + h2) * Lq; sc2 = 0;
                for (int q2 = 0; q2 < nq; ++q2) {
                  tc::tma_prefetch_2d(&tmDO, h2 * HD, b2 * Lq + q2 * TL);
                  tc::tma_prefetch_2d(&tmQ, h2 * HD, b2 * Lq + q2 * TL);
                }
                tc::tma_prefetch_2d(&tmV, h2 * HD, b2 * Lk);
                tc::tma_prefetch_2d(&tmK, h2 * HD, b2 * Lk);
              }
            }
            if (sr2 >= 0) {
              if (a.has_s) {
                tc::tma_prefetch_2d(&tmS, sc2, sr2);
                tc::tma_prefetch_2d(&tmS, sc2 + 64, sr2);
              }
              if (a.has_dsn) {
                tc::tma_prefetch_2d(&tmDSn, sc2, sr2);
                tc::tma_prefetch_2d(&tmDSn, sc2 + 64, sr2);
              }
              if (a.has_prev) {
                tc::tma_prefetch_2d(&tmSprev, sc2, sr2);
                tc::tma_prefetch_2d(&tmSprev, sc2 + 64, sr2);
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      tc::mbar_wait(bar_in, 0);
      int t = 0;
      for (int kt = 0; kt < nk; ++kt) {
        tc::mbar_wait(bar_kv, kt & 1);
        for (int qt = 0; qt < nq; ++qt, ++t) {
          tc::tc_fence_after();
          const uint32_t q_t = base + B_OFF_Q + qt * T16, do_t = base + B_OFF_DO + qt * T16;
          if (!a.has_s) {
#pragma unroll
            for (int k = 0; k < HD / 16; ++k)
              tc::umma_bf16(tm_S, tc::smem_desc_sw128(q_t + k * 32, 16, 1024),
                            tc::smem_desc_sw128(base + B_OFF_K + k * 32, 16, 1024), a.idesc_nn128,
                            k > 0 ? 1u : 0u);
          }
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
            tc::umma_bf16(tm_dP, tc::smem_desc_sw128(do_t + k * 32, 16, 1024),
                          tc::smem_desc_sw128(base + B_OFF_V + k * 32, 16, 1024), a.idesc_nn128,
                          k > 0 ? 1u : 0u);
          tc::umma_commit(bar_mm1);
          tc::mbar_wait(bar_p2, t & 1);
          tc::tc_fence_after();
#pragma unroll
          for (int k = 0; k < TL / 16; ++k) {   // K = query rows of this tile
            const uint64_t pT = tc::smem_desc_sw128(base + B_OFF_A + k * 2048, T16, 1024);
            const uint64_t dsT = tc::smem_desc_sw128(base + B_OFF_B + k * 2048, T16, 1024);
            const uint64_t dOm = tc::smem_desc_sw128(do_t + k * 2048, T16, 1024);
            const uint64_t Qm = tc::smem_desc_sw128(q_t + k * 2048, T16, 1024);
            const uint32_t acc = (qt > 0 || k > 0) ? 1u : 0u;
            tc::umma_bf16(tm_dV, pT, dOm, a.idesc_tt64, acc);
            tc::umma_bf16(tm_dK, dsT, Qm, a.idesc_tt64, acc);
          }
#pragma unroll
          for (int k = 0; k < TL / 16; ++k) {   // K = keys of this tile
            const uint64_t dsK =
                tc::smem_desc_sw128(base + B_OFF_B + (k >> 2) * T16 + (k & 3) * 32, 16, 1024);
            const uint64_t Km = tc::smem_desc_sw128(base + B_OFF_K + k * 2048, T16, 1024);
            tc::umma_bf16(tm_dQ + 64 * qt, dsK, Km, a.idesc_nt64, (kt > 0 || k > 0) ? 1u : 0u);
          }
          tc::umma_commit(bar_mm2);
        }
      }
    }
  } else {
    // ============ thread = (query row of the tile, 64-key half of the key tile) =================
    const int quarter = warp & 3;
    const int cg = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const float cval = (a.has_prev && a.c) ? a.c[0] : 0.f;
    const float kLog2e = 1.4426950408889634f;
    float D[MAX_NQ], mxr[MAX_NQ], invr[MAX_NQ];
    // D = rowsum(dO * O) over this head's 64 columns; each thread of the pair sums 32 of them
#pragma unroll
    for (int qt = 0; qt < MAX_NQ; ++qt) {
      D[qt] = 0.f; mxr[qt] = 0.f; invr[qt] = 0.f;
      if (qt < nq) {
        const int64_t gr = (int64_t)qrow0 + qt * TL + row;
        const uint4* pd = reinterpret_cast<const uint4*>(a.d_o + gr * a.lddo + h * HD + cg * 32);
        const uint4* po = reinterpret_cast<const uint4*>(a.o + gr * a.ldo + h * HD + cg * 32);
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 x = pd[i], y = po[i];
          const uint32_t xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            acc = fmaf(bf16_lo(xs[e]), bf16_lo(ys[e]), acc);
            acc = fmaf(bf16_hi(xs[e]), bf16_hi(ys[e]), acc);
          }
        }
        red[(cg * MAX_NQ + qt) * TL + row] = acc;
        const float* st2 = a.stat + 2 * ((int64_t)srow0 + qt * TL + row);
        mxr[qt] = st2[0];
        invr[qt] = 1.0f / st2[1];
      }
    }
    soft_sync();
#pragma unroll
    for (int qt = 0; qt < MAX_NQ; ++qt)
      if (qt < nq) D[qt] = red[qt * TL + row] + red[(MAX_NQ + qt) * TL + row];

    float dc_part = 0.f;
    int t = 0;
    for (int kt = 0; kt < nk; ++kt) {
#pragma unroll 1
      for (int qt = 0; qt < nq; ++qt, ++t) {
        const uint32_t ph = t & 1;
        const float mx = qt == 0 ? mxr[0] : mxr[1], inv = qt == 0 ? invr[0] : invr[1];
        const float Dq = qt == 0 ? D[0] : D[1];
        tc::mbar_wait(bar_mm1, ph);
        tc::tc_fence_after();
        if (a.has_s) tc::mbar_wait(bar_s, ph);
        if (a.has_prev) tc::mbar_wait(bar_sp, ph);
        if (a.has_dsn) tc::mbar_wait(bar_dsn, ph);
#pragma unroll 1
        for (int ch = 0; ch < 2; ++ch) {
          uint32_t dp[32], sa[32];
          const int col = cg * 64 + ch * 32;           // column inside the 128-key tile
          tc::tmem_ld32(tm_dP + lane_off + col, dp);
          if (!a.has_s) tc::tmem_ld32(tm_S + lane_off + col, sa);
          tc::tmem_ld_wait();
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const uint32_t off = (uint32_t)cg * T16 + tc::sw128_offset(row, ch * 4 + q4);
            uint32_t sv[4] = {0u, 0u, 0u, 0u}, sp[4] = {0u, 0u, 0u, 0u}, dn[4] = {0u, 0u, 0u, 0u};
            uint32_t outp[4], outd[4], outc[4];
            if (a.has_s) lds128(base + B_OFF_A + off, sv);
            if (a.has_prev) lds128(base + B_OFF_C + off, sp);
            if (a.has_dsn) lds128(base + B_OFF_B + off, dn);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int i0 = q4 * 8 + e * 2, j = kt * TL + col + i0;
              const float2 mk = *reinterpret_cast<const float2*>(mask_s + j);
              float s0, s1;
              if (a.has_s) {
                s0 = bf16_lo(sv[e]);
                s1 = bf16_hi(sv[e]);
              } else {
                s0 = __uint_as_float(sa[i0]) * 0.125f;       // == / sqrt(64), exact
                s1 = __uint_as_float(sa[i0 + 1]) * 0.125f;
                if (a.has_prev) {
                  s0 = __fadd_rn(s0, __fmul_rn(cval, bf16_lo(sp[e])));
                  s1 = __fadd_rn(s1, __fmul_rn(cval, bf16_hi(sp[e])));
                }
                s0 = __fsub_rn(s0, mk.x);
                s1 = __fsub_rn(s1, mk.y);
                const uint32_t rr = pack_bf16(s0, s1);   // the forward softmax saw bf16 scores
                s0 = bf16_lo(rr);
                s1 = bf16_hi(rr);
              }
              const float p0 = tc::ex2((s0 - mx) * kLog2e) * inv, p1 = tc::ex2((s1 - mx) * kLog2e) * inv;
              outp[e] = pack_bf16(p0, p1);
              float d0 = bf16_lo(outp[e]) * (__uint_as_float(dp[i0]) - Dq);
              float d1 = bf16_hi(outp[e]) * (__uint_as_float(dp[i0 + 1]) - Dq);
              if (a.has_dsn) {
                d0 += bf16_lo(dn[e]);
                d1 += bf16_hi(dn[e]);
              }
              outd[e] = pack_bf16(d0, d1);
              if (a.has_prev) {
                dc_part = fmaf(d0, bf16_lo(sp[e]), dc_part);
                dc_part = fmaf(d1, bf16_hi(sp[e]), dc_part);
                outc[e] = pack_bf16(cval * d0, cval * d1);
              }
            }
            sts128(base + B_OFF_A + off, outp);
            sts128(base + B_OFF_B + off, outd);
            if (a.has_prev && a.write_dsp) sts128(base + B_OFF_C + off, outc);
          }
        }
        tc::tc_fence_before();       // the dP / S columns are overwritten by the next step's MMAs
        tc::fence_proxy_async();
        tc::mbar_arrive(bar_p2);
        soft_sync();
        const bool last_q = qt == nq - 1;
        if (threadIdx.x == 64 && a.has_prev && a.write_dsp) {
          const int sr = srow0 + qt * TL, sc = kt * TL;
          tc::tma_store_2d(&tmDSp, base + B_OFF_C, sc, sr);
          tc::tma_store_2d(&tmDSp, base + B_OFF_C + T16, sc + 64, sr);
          tc::tma_store_commit();
        }
        if (last_q) {
          // dV / dK of this key tile are complete: TMEM -> bf16 staging in the (dead) V / K tiles
          tc::mbar_wait(bar_mm2, ph);
          tc::tc_fence_after();
#pragma unroll 1
          for (int w = 0; w < 2; ++w) {
            uint32_t r[32];
            tc::tmem_ld32((w == 0 ? tm_dV : tm_dK) + lane_off + cg * 32, r);
            tc::tmem_ld_wait();
            const float sc = w == 0 ? 1.0f : 0.125f;
            const uint32_t dst = base + (w == 0 ? B_OFF_V : B_OFF_K);
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              uint32_t out[4];
#pragma unroll
              for (int e = 0; e < 4; ++e)
                out[e] = pack_bf16(__uint_as_float(r[q4 * 8 + e * 2]) * sc,
                                   __uint_as_float(r[q4 * 8 + e * 2 + 1]) * sc);
              sts128(dst + tc::sw128_offset(row, cg * 4 + q4), out);
            }
          }
          tc::tc_fence_before();
          tc::fence_proxy_async();
          soft_sync();
          if (threadIdx.x == 64) {
            tc::tma_store_2d(&tmDV, base + B_OFF_V, h * HD, krow0 + kt * TL);
            tc::tma_store_2d(&tmDK, base + B_OFF_K, h * HD, krow0 + kt * TL);
            tc::tma_store_commit();
          }
        }
        if (threadIdx.x == 64) {
          if (!last_q) tc::mbar_wait(bar_mm2, ph);   // A / B / K / V consumed by the MMAs
          tc::tma_store_wait_read();                 // C / K / V staging read by the bulk stores
          tc::mbar_arrive(bar_free);
        }
      }
    }
    if (a.has_prev && a.dc) {
      dc_part = warp_sum(dc_part);
      if (lane == 0) atomicAdd(a.dc, dc_part);
    }
    // ---- dQ of both query tiles: TMEM -> staging in the (dead) Q tiles -> TMA stores ------------
    tc::mbar_wait(bar_mm2, (t - 1) & 1);
    tc::tc_fence_after();
    for (int qt = 0; qt < nq; ++qt) {
      uint32_t r[32];
      tc::tmem_ld32(tm_dQ + 64 * qt + lane_off + cg * 32, r);
      tc::tmem_ld_wait();
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        uint32_t out[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
          out[e] = pack_bf16(__uint_as_float(r[q4 * 8 + e * 2]) * 0.125f,
                             __uint_as_float(r[q4 * 8 + e * 2 + 1]) * 0.125f);
        sts128(base + B_OFF_Q + qt * T16 + tc::sw128_offset(row, cg * 4 + q4), out);
      }
    }
    tc::fence_proxy_async();
    soft_sync();
    if (threadIdx.x == 64) {
      for (int qt = 0; qt < nq; ++qt)
        tc::tma_store_2d(&tmDQ, base + B_OFF_Q + qt * T16, h * HD, qrow0 + qt * TL);
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

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

bool make2d(CUtensorMap* tm, const void* base, uint64_t cols, uint64_t rows, uint64_t ld_elems,
            uint32_t box_rows) {
  const uint64_t dims[2] = {cols, rows};
  const uint64_t str[1] = {ld_elems * 2};
  const uint32_t box[2] = {64, box_rows};
  return mm_make_tmap_bf16(tm, base, 2, dims, str, box);
}

}  // namespace

// One CTA per SM is resident, so the CTA that follows on an SM is about one wave (= SM count)
// ahead in launch order.  MMEMO_ATTN_PF=0 turns the L2 prefetch off, =N overrides the distance.
int resattn_pf_distance() {
  const char* env = getenv("MMEMO_ATTN_PF");
  if (env) return atoi(env);
  static int cached = 0;   // same value from every thread: a benign race
  if (cached == 0) {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess)
      cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cached = n;
  }
  return cached;
}

bool resattn_tc2_supported(const mmemo_attn_problem& p, bool bwd) {
  // MMEMO_ATTN_TC2=0 routes these shapes to the mma.sync kernels (A/B timing, tools/prof_attn.py)
  const char* env = getenv("MMEMO_ATTN_TC2");
  if (env && env[0] == '0') return false;
  // the backward kernel is generic in the key-tile count; MMEMO_ATTN_TC2=2 also sends the single
  // tile (L = 128) backward through it instead of resattn_tc.cu's two-pass kernel
  const bool one_tile_bwd = bwd && env && env[0] == '2' && p.Lk == TL && p.Lq == TL;
  if (p.hd != HD || (p.Lq != TL && p.Lq != 2 * TL)) return false;
  if (p.Lk != NK * TL && !one_tile_bwd) return false;
  if (p.ldq % 8 || p.ldk % 8 || p.ldv % 8 || p.lds % 8 || p.lds < p.Lk || !p.lse) return false;
  if (!bwd) return p.ldo % 8 == 0;
  return p.o && p.ldo % 8 == 0 && p.lddo % 8 == 0 && p.lddq % 8 == 0 && p.lddk % 8 == 0 &&
         p.lddv % 8 == 0;
}

int resattn_fwd_tc2(const mmemo_attn_problem& p, cudaStream_t st) {
  if (p.B <= 0 || p.H <= 0) return MMEMO_OK;
  MM_REQUIRE(p.q && p.k && p.v && p.o && p.lse);
  const void* ptrs[] = {p.q, p.k, p.v, p.o, p.s_prev, p.s_out};
  for (const void* x : ptrs)
    if (x && !aligned16(x)) return MMEMO_ERR_ARG;
  MM_CUDA_OK(cudaFuncSetAttribute(resattn_fwd_tc2_kernel,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_FWD2));
  CUtensorMap tmQ, tmK, tmV, tmSp, tmSo, tmO;
  const uint64_t qrows = (uint64_t)p.B * p.Lq, krows = (uint64_t)p.B * p.Lk,
                 srows = (uint64_t)p.B * p.H * p.Lq, d = (uint64_t)p.H * HD;
  bool ok = make2d(&tmQ, p.q, d, qrows, p.ldq, TL) && make2d(&tmK, p.k, d, krows, p.ldk, NK * TL) &&
            make2d(&tmV, p.v, d, krows, p.ldv, NK * TL) && make2d(&tmO, p.o, d, qrows, p.ldo, TL);
  auto score_map = [&](CUtensorMap* tm, const void* x) {
    return x ? make2d(tm, x, p.Lk, srows, p.lds, TL) : make2d(tm, p.q, d, qrows, p.ldq, TL);
  };
  ok = ok && score_map(&tmSp, p.s_prev) && score_map(&tmSo, p.s_out);
  if (!ok) {
    mmemo_set_error("cuTensorMapEncodeTiled failed (resattn_fwd_tc2)", __FILE__, __LINE__);
    return MMEMO_ERR_CUDA;
  }
  Fwd2Args a = {};
  a.mask = p.mask; a.mask_bs = p.mask_bs; a.c = p.c; a.stat = p.lse; a.H = (int)p.H;
  a.nq = (int)(p.Lq / TL);
  a.has_prev = p.s_prev != nullptr; a.write_s = p.s_out != nullptr;
  a.pf_dist = resattn_pf_distance();
  a.idesc_qk = tc::idesc_bf16(TL, NK * TL, 0, 0);
  a.idesc_pv = tc::idesc_bf16(TL, HD, 0, 1);
  dim3 grid((unsigned)p.H, (unsigned)(p.B * a.nq));
  MM_CUDA_OK(mm_launch(resattn_fwd_tc2_kernel, grid, dim3(NTHREADS), SMEM_FWD2, st, tmQ, tmK, tmV,
                       tmSp, tmSo, tmO, a));
  return MMEMO_OK;
}

int resattn_bwd_tc2(const mmemo_attn_problem& p, cudaStream_t st) {
  if (p.B <= 0 || p.H <= 0) return MMEMO_OK;
  MM_REQUIRE(p.d_o && p.q && p.k && p.v && p.o && p.lse && p.dq && p.dk && p.dv);
  const void* ptrs[] = {p.d_o, p.q, p.k, p.v, p.o, p.dq, p.dk, p.dv, p.s, p.s_prev, p.ds_next,
                        p.ds_prev};
  for (const void* x : ptrs)
    if (x && !aligned16(x)) return MMEMO_ERR_ARG;
  MM_CUDA_OK(cudaFuncSetAttribute(resattn_bwd_tc2_kernel,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BWD2));
  CUtensorMap tmQ, tmK, tmV, tmDO, tmS, tmSp, tmDSn, tmDSp, tmDQ, tmDK, tmDV;
  const uint64_t qrows = (uint64_t)p.B * p.Lq, krows = (uint64_t)p.B * p.Lk,
                 srows = (uint64_t)p.B * p.H * p.Lq, d = (uint64_t)p.H * HD;
  bool ok = make2d(&tmQ, p.q, d, qrows, p.ldq, TL) && make2d(&tmK, p.k, d, krows, p.ldk, TL) &&
            make2d(&tmV, p.v, d, krows, p.ldv, TL) && make2d(&tmDO, p.d_o, d, qrows, p.lddo, TL) &&
            make2d(&tmDQ, p.dq, d, qrows, p.lddq, TL) && make2d(&tmDK, p.dk, d, krows, p.lddk, TL) &&
            make2d(&tmDV, p.dv, d, krows, p.lddv, TL);
  auto score_map = [&](CUtensorMap* tm, const void* x) {
    return x ? make2d(tm, x, p.Lk, srows, p.lds, TL) : make2d(tm, p.q, d, qrows, p.ldq, TL);
  };
  ok = ok && score_map(&tmS, p.s) && score_map(&tmSp, p.s_prev) && score_map(&tmDSn, p.ds_next) &&
       score_map(&tmDSp, p.ds_prev);
  if (!ok) {
    mmemo_set_error("cuTensorMapEncodeTiled failed (resattn_bwd_tc2)", __FILE__, __LINE__);
    return MMEMO_ERR_CUDA;
  }
  Bwd2Args a = {};
  a.mask = p.mask; a.mask_bs = p.mask_bs; a.c = p.c; a.stat = p.lse; a.dc = p.dc;
  a.d_o = static_cast<const __nv_bfloat16*>(p.d_o); a.o = static_cast<const __nv_bfloat16*>(p.o);
  a.lddo = p.lddo; a.ldo = p.ldo;
  a.H = (int)p.H; a.nq = (int)(p.Lq / TL); a.nk = (int)(p.Lk / TL);
  a.has_s = p.s != nullptr; a.has_prev = p.s_prev != nullptr; a.has_dsn = p.ds_next != nullptr;
  a.write_dsp = p.ds_prev != nullptr;
  a.pf_dist = resattn_pf_distance();
  a.idesc_nn128 = tc::idesc_bf16(TL, TL, 0, 0);
  a.idesc_tt64 = tc::idesc_bf16(TL, HD, 1, 1);
  a.idesc_nt64 = tc::idesc_bf16(TL, HD, 0, 1);
  dim3 grid((unsigned)p.H, (unsigned)p.B);
  MM_CUDA_OK(mm_launch(resattn_bwd_tc2_kernel, grid, dim3(NTHREADS), SMEM_BWD2, st, tmQ, tmK, tmV,
                       tmDO, tmS, tmSp, tmDSn, tmDSp, tmDQ, tmDK, tmDV, a));
  return MMEMO_OK;
}
