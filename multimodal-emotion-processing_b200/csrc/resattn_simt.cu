// Fused residual-attention core, SIMT version (fp32 math on CUDA cores).
//
// Serves the float32 parity mode for every shape and, in the bf16 build, the head sizes / lengths
// the tcgen05 kernel (resattn_tc.cu) does not take (hd = 16/32, ragged L).  One launch replaces the
// reference's  bmm -> div -> mul/add (c*S_prev) -> rsub/mul/sub_ (mask) -> softmax -> bmm ->
// transpose/contiguous  sequence (others/realformer.py:189-203 and its three twins).
//
// Forward : CTA = (query tile of 32 rows, head, batch).  K_h and V_h live in shared memory as fp32;
//           each warp owns 4 query rows at a time, lanes own keys (QK^T) then output dims (PV);
//           the score row stays in registers between QK^T, residual add, mask, softmax.
// Backward: CTA = (head, batch); flash-style loop over key tiles (<=128 keys) and 32-row query
//           chunks; P is rebuilt from the stored S and the saved row (max, sum); D = rowsum(dO*O).
#include <math.h>

#include "common.cuh"

#include "resattn.h"

namespace {

constexpr int R = 4;         // query rows per warp iteration
constexpr int FWD_WARPS = 4;
constexpr int QTILE = 32;    // query rows per forward CTA
constexpr int BWD_WARPS = 8;
constexpr int CHUNK = BWD_WARPS * R;  // 32 query rows per backward chunk
constexpr int MAXD = 4;      // output dims per lane (hd <= 128)

struct AttnArgs {
  const void *q, *k, *v;
  int64_t ldq, ldk, ldv;
  const float* mask;
  int64_t mask_bs, mask_rs;
  const void* s_prev;
  const float* c;
  void* s_out;
  void* o;
  int64_t ldo;
  float* lse;
  int B, H, Lq, Lk, hd;
  float sqrt_hd;
  // backward only
  const void *d_o, *s, *ds_next;
  int64_t lddo;
  void *dq, *dk, *dv, *ds_prev;
  int64_t lddq, lddk, lddv;
  float* dc;
  float* dq_ws;
  int KT;
};

__device__ __forceinline__ int lanes_per_group(int hd) {
  int l = 1;
  while (l < hd && l < 32) l <<= 1;
  return l;
}

// score of one (row, key) before softmax, in the reference's fp32 op order, rounded to T
template <typename T>
__device__ __forceinline__ float finish_score(float dot, float sqrt_hd, bool has_prev, float c,
                                              float prev, bool has_mask, float m) {
  float s = dot / sqrt_hd;
  if (has_prev) s = __fadd_rn(s, __fmul_rn(c, prev));
  if (has_mask) s = __fsub_rn(s, __fmul_rn(1.0e8f, __fsub_rn(1.0f, m)));
  return round_to<T>(s);
}

template <typename T, int MAXC>
__global__ void __launch_bounds__(FWD_WARPS * 32) resattn_fwd_kernel(AttnArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int hd = a.hd, Lk = a.Lk, Lq = a.Lq;
  const int kst = hd + 4;                       // padded row stride of Ks / Vs (float4 aligned)
  const int LkP = (Lk + 3) & ~3;
  float* Ks = smem;
  float* Vs = Ks + (size_t)Lk * kst;
  float* qs = Vs + (size_t)Lk * kst;            // [FWD_WARPS][R][hd]
  float* ps = qs + FWD_WARPS * R * hd;          // [FWD_WARPS][R][LkP]
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * QTILE;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const T* __restrict__ qg = static_cast<const T*>(a.q);
  const T* __restrict__ kg = static_cast<const T*>(a.k);
  const T* __restrict__ vg = static_cast<const T*>(a.v);

  for (int idx = tid; idx < Lk * hd; idx += FWD_WARPS * 32) {
    const int j = idx / hd, kk = idx - j * hd;
    Ks[j * kst + kk] = to_f(kg[((int64_t)b * Lk + j) * a.ldk + h * hd + kk]);
    Vs[j * kst + kk] = to_f(vg[((int64_t)b * Lk + j) * a.ldv + h * hd + kk]);
  }
  __syncthreads();

  const bool has_prev = a.s_prev != nullptr, has_mask = a.mask != nullptr;
  const float cval = (has_prev && a.c) ? a.c[0] : 0.f;
  const T* __restrict__ sprev = static_cast<const T*>(a.s_prev);
  T* __restrict__ sout = static_cast<T*>(a.s_out);
  float* myq = qs + warp * R * hd;
  float* myp = ps + warp * R * LkP;
  const int LPG = lanes_per_group(hd), G = 32 / LPG, grp = lane / LPG, dl = lane % LPG;
  const int nchunk = (Lk + 31) >> 5;
  const int ndl = (hd + 31) >> 5;  // output dims per lane actually used

  for (int rbase = q0 + warp * R; rbase < min(q0 + QTILE, Lq); rbase += FWD_WARPS * R) {
    // ---- stage the R query rows -----------------------------------------------------------
    for (int idx = lane; idx < R * hd; idx += 32) {
      const int r = idx / hd, kk = idx - r * hd;
      const int row = min(rbase + r, Lq - 1);
      myq[idx] = to_f(qg[((int64_t)b * Lq + row) * a.ldq + h * hd + kk]);
    }
    __syncwarp();
    // ---- QK^T: lane owns keys lane, lane+32, ... ------------------------------------------
    float sc[R][MAXC];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < MAXC; ++c) sc[r][c] = 0.f;
    for (int k4 = 0; k4 < hd; k4 += 4) {
      float4 qv[R];
#pragma unroll
      for (int r = 0; r < R; ++r) qv[r] = *reinterpret_cast<const float4*>(myq + r * hd + k4);
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        if (c < nchunk) {
          const int j = min(c * 32 + lane, Lk - 1);
          const float4 kv = *reinterpret_cast<const float4*>(Ks + j * kst + k4);
#pragma unroll
          for (int r = 0; r < R; ++r) {
            sc[r][c] = fmaf(qv[r].x, kv.x, sc[r][c]);
            sc[r][c] = fmaf(qv[r].y, kv.y, sc[r][c]);
            sc[r][c] = fmaf(qv[r].z, kv.z, sc[r][c]);
            sc[r][c] = fmaf(qv[r].w, kv.w, sc[r][c]);
          }
        }
      }
    }
    // ---- residual, mask, store S, softmax -------------------------------------------------
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = rbase + r;
      const bool row_ok = row < Lq;
      const int64_t srow = (((int64_t)b * a.H + h) * Lq + (row_ok ? row : 0)) * Lk;
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        const int j = c * 32 + lane;
        float s = -INFINITY;
        if (c < nchunk && j < Lk) {
          const float prev = has_prev ? to_f(sprev[srow + j]) : 0.f;
          const float m = has_mask
                              ? a.mask[(int64_t)b * a.mask_bs + (row_ok ? row : 0) * a.mask_rs + j]
                              : 1.f;
          s = finish_score<T>(sc[r][c], a.sqrt_hd, has_prev, cval, prev, has_mask, m);
          if (sout && row_ok) sout[srow + j] = from_f<T>(s);
        }
        sc[r][c] = s;
        mx = fmaxf(mx, s);
      }
      mx = warp_max(mx);
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        const float e = (sc[r][c] == -INFINITY) ? 0.f : expf(sc[r][c] - mx);
        sc[r][c] = e;
        sum += e;
      }
      sum = warp_sum(sum);
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        const int j = c * 32 + lane;
        if (c < nchunk && j < Lk) myp[r * LkP + j] = sc[r][c] / sum;
      }
      if (lane == 0 && row_ok && a.lse) {
        float* st2 = a.lse + 2 * (((int64_t)b * a.H + h) * Lq + row);
        st2[0] = mx;   // kept separately (not as max+log(sum)): a fully masked row has max = -1e8,
        st2[1] = sum;  // where adding log(sum) would be absorbed by fp32 rounding
      }
    }
    __syncwarp();
    // ---- PV: lane owns (key residue class, output dims) -----------------------------------
    float acc[R][MAXD];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int t = 0; t < MAXD; ++t) acc[r][t] = 0.f;
    for (int j = grp; j < Lk; j += G) {
      float vv[MAXD];
#pragma unroll
      for (int t = 0; t < MAXD; ++t) {
        const int dim = dl + 32 * t;
        vv[t] = (dim < hd) ? Vs[j * kst + dim] : 0.f;
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float p = myp[r * LkP + j];
#pragma unroll
        for (int t = 0; t < MAXD; ++t)
          if (t < ndl) acc[r][t] = fmaf(p, vv[t], acc[r][t]);
      }
    }
    for (int off = LPG; off < 32; off <<= 1)
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int t = 0; t < MAXD; ++t) acc[r][t] += __shfl_xor_sync(0xffffffffu, acc[r][t], off);
    if (grp == 0) {
      T* __restrict__ og = static_cast<T*>(a.o);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int row = rbase + r;
        if (row >= Lq) continue;
#pragma unroll
        for (int t = 0; t < MAXD; ++t) {
          const int dim = dl + 32 * t;
          if (dim < hd) og[((int64_t)b * Lq + row) * a.ldo + h * hd + dim] = from_f<T>(acc[r][t]);
        }
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(BWD_WARPS * 32) resattn_bwd_kernel(AttnArgs a) {
  constexpr int MAXC = 4;  // key tile <= 128
  extern __shared__ __align__(16) float smem[];
  const int hd = a.hd, Lk = a.Lk, Lq = a.Lq, KT = a.KT;
  const int kst = hd + 4;
  float* Ks = smem;
  float* Vs = Ks + (size_t)KT * kst;
  float* dKs = Vs + (size_t)KT * kst;
  float* dVs = dKs + (size_t)KT * kst;
  float* qs = dVs + (size_t)KT * kst;   // [CHUNK][kst]
  float* dOs = qs + CHUNK * kst;        // [CHUNK][kst]
  float* Ps = dOs + CHUNK * kst;        // [CHUNK][KT]
  float* dSs = Ps + CHUNK * KT;         // [CHUNK][KT]
  __shared__ float red[BWD_WARPS];
  const int b = blockIdx.y, h = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const T* __restrict__ qg = static_cast<const T*>(a.q);
  const T* __restrict__ kg = static_cast<const T*>(a.k);
  const T* __restrict__ vg = static_cast<const T*>(a.v);
  const T* __restrict__ dog = static_cast<const T*>(a.d_o);
  const T* __restrict__ og = static_cast<const T*>(a.o);
  const T* __restrict__ sg = static_cast<const T*>(a.s);
  const T* __restrict__ sprev = static_cast<const T*>(a.s_prev);
  const T* __restrict__ dsn = static_cast<const T*>(a.ds_next);
  T* __restrict__ dsp = static_cast<T*>(a.ds_prev);
  const bool has_prev = sprev != nullptr, has_mask = a.mask != nullptr;
  const float cval = (has_prev && a.c) ? a.c[0] : 0.f;
  const float inv_sqrt = 1.0f / a.sqrt_hd;
  const int LPG = lanes_per_group(hd), G = 32 / LPG, grp = lane / LPG, dl = lane % LPG;
  const bool multi_tile = Lk > KT;
  const int ndl = (hd + 31) >> 5;
  float* dqacc = a.dq_ws ? a.dq_ws : reinterpret_cast<float*>(a.dq);  // fp32 accumulator
  float dc_part = 0.f;

  for (int kt0 = 0; kt0 < Lk; kt0 += KT) {
    const int klen = min(KT, Lk - kt0);
    const int nchunk = (klen + 31) >> 5;
    for (int idx = tid; idx < KT * hd; idx += BWD_WARPS * 32) {
      const int j = idx / hd, kk = idx - j * hd;
      float kv = 0.f, vv = 0.f;
      if (j < klen) {
        kv = to_f(kg[((int64_t)b * Lk + kt0 + j) * a.ldk + h * hd + kk]);
        vv = to_f(vg[((int64_t)b * Lk + kt0 + j) * a.ldv + h * hd + kk]);
      }
      Ks[j * kst + kk] = kv;
      Vs[j * kst + kk] = vv;
      dKs[j * kst + kk] = 0.f;
      dVs[j * kst + kk] = 0.f;
    }
    __syncthreads();

    for (int r0 = 0; r0 < Lq; r0 += CHUNK) {
      // ================= phase A: row-parallel (each warp: R rows) ===========================
      const int lr0 = warp * R;
      float Drow[R];
      for (int idx = lane; idx < R * hd; idx += 32) {
        const int r = idx / hd, kk = idx - r * hd;
        const int row = r0 + lr0 + r;
        float qv = 0.f, dv = 0.f;
        if (row < Lq) {
          qv = to_f(qg[((int64_t)b * Lq + row) * a.ldq + h * hd + kk]);
          dv = to_f(dog[((int64_t)b * Lq + row) * a.lddo + h * hd + kk]);
        }
        qs[(lr0 + r) * kst + kk] = qv;
        dOs[(lr0 + r) * kst + kk] = dv;
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int row = r0 + lr0 + r;
        float d = 0.f;
        if (row < Lq)
          for (int kk = lane; kk < hd; kk += 32)
            d += to_f(dog[((int64_t)b * Lq + row) * a.lddo + h * hd + kk]) *
                 to_f(og[((int64_t)b * Lq + row) * a.ldo + h * hd + kk]);
        Drow[r] = warp_sum(d);
      }
      __syncwarp();
      // dP = dO V^T (and, when S was not stored, the recomputed QK^T)
      float dp[R][MAXC], sc[R][MAXC];
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < MAXC; ++c) { dp[r][c] = 0.f; sc[r][c] = 0.f; }
      const bool recompute = (sg == nullptr);
      for (int k4 = 0; k4 < hd; k4 += 4) {
        float4 dov[R], qv[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          dov[r] = *reinterpret_cast<const float4*>(dOs + (lr0 + r) * kst + k4);
          qv[r] = *reinterpret_cast<const float4*>(qs + (lr0 + r) * kst + k4);
        }
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
          if (c < nchunk) {
            const int j = c * 32 + lane;  // j < KT always (rows beyond klen are zero-filled)
            const float4 vv = *reinterpret_cast<const float4*>(Vs + j * kst + k4);
#pragma unroll
            for (int r = 0; r < R; ++r) {
              dp[r][c] = fmaf(dov[r].x, vv.x, dp[r][c]);
              dp[r][c] = fmaf(dov[r].y, vv.y, dp[r][c]);
              dp[r][c] = fmaf(dov[r].z, vv.z, dp[r][c]);
              dp[r][c] = fmaf(dov[r].w, vv.w, dp[r][c]);
            }
            if (recompute) {
              const float4 kv = *reinterpret_cast<const float4*>(Ks + j * kst + k4);
#pragma unroll
              for (int r = 0; r < R; ++r) {
                sc[r][c] = fmaf(qv[r].x, kv.x, sc[r][c]);
                sc[r][c] = fmaf(qv[r].y, kv.y, sc[r][c]);
                sc[r][c] = fmaf(qv[r].z, kv.z, sc[r][c]);
                sc[r][c] = fmaf(qv[r].w, kv.w, sc[r][c]);
              }
            }
          }
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int row = r0 + lr0 + r;
        const bool row_ok = row < Lq;
        const int64_t srow = (((int64_t)b * a.H + h) * Lq + (row_ok ? row : 0)) * Lk + kt0;
        const float* st2 = a.lse + 2 * (((int64_t)b * a.H + h) * Lq + (row_ok ? row : 0));
        const float rmax = st2[0], rinv = 1.0f / st2[1];
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
          const int j = c * 32 + lane;
          if (c >= nchunk) continue;
          float p = 0.f, ds = 0.f;
          if (row_ok && j < klen) {
            float s;
            if (recompute) {
              const float m = has_mask
                                  ? a.mask[(int64_t)b * a.mask_bs + row * a.mask_rs + kt0 + j]
                                  : 1.f;
              const float prev = has_prev ? to_f(sprev[srow + j]) : 0.f;
              s = finish_score<T>(sc[r][c], a.sqrt_hd, has_prev, cval, prev, has_mask, m);
            } else {
              s = to_f(sg[srow + j]);
            }
            p = expf(s - rmax) * rinv;
            ds = p * (dp[r][c] - Drow[r]);
            if (dsn) ds += to_f(dsn[srow + j]);
            if (has_prev) {
              dc_part = fmaf(ds, to_f(sprev[srow + j]), dc_part);
              if (dsp) dsp[srow + j] = from_f<T>(cval * ds);
            }
          }
          Ps[(lr0 + r) * KT + j] = p;
          dSs[(lr0 + r) * KT + j] = ds;
        }
      }
      __syncwarp();
      // dQ rows = dS K / sqrt(hd)
      {
        float acc[R][MAXD];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int t = 0; t < MAXD; ++t) acc[r][t] = 0.f;
        for (int j = grp; j < klen; j += G) {
          float kv[MAXD];
#pragma unroll
          for (int t = 0; t < MAXD; ++t) {
            const int dim = dl + 32 * t;
            kv[t] = (dim < hd) ? Ks[j * kst + dim] : 0.f;
          }
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float ds = dSs[(lr0 + r) * KT + j];
#pragma unroll
            for (int t = 0; t < MAXD; ++t)
              if (t < ndl) acc[r][t] = fmaf(ds, kv[t], acc[r][t]);
          }
        }
        for (int off = LPG; off < 32; off <<= 1)
#pragma unroll
          for (int r = 0; r < R; ++r)
#pragma unroll
            for (int t = 0; t < MAXD; ++t)
              acc[r][t] += __shfl_xor_sync(0xffffffffu, acc[r][t], off);
        if (grp == 0) {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const int row = r0 + lr0 + r;
            if (row >= Lq) continue;
#pragma unroll
            for (int t = 0; t < MAXD; ++t) {
              const int dim = dl + 32 * t;
              if (dim >= hd) continue;
              const float val = acc[r][t] * inv_sqrt;
              if (!multi_tile) {
                static_cast<T*>(a.dq)[((int64_t)b * Lq + row) * a.lddq + h * hd + dim] =
                    from_f<T>(val);
              } else {
                const int64_t ldacc = a.dq_ws ? (int64_t)a.H * hd : a.lddq;
                float* dst = dqacc + ((int64_t)b * Lq + row) * ldacc + h * hd + dim;
                *dst = (kt0 == 0) ? val : (*dst + val);
              }
            }
          }
        }
      }
      __syncthreads();
      // ================= phase B: (key, dim)-parallel accumulation of dK, dV =================
      for (int e = tid; e < klen * hd; e += BWD_WARPS * 32) {
        const int j = e / hd, dim = e - j * hd;
        float dkv = 0.f, dvv = 0.f;
#pragma unroll 8
        for (int r = 0; r < CHUNK; ++r) {
          dkv = fmaf(dSs[r * KT + j], qs[r * kst + dim], dkv);
          dvv = fmaf(Ps[r * KT + j], dOs[r * kst + dim], dvv);
        }
        dKs[j * kst + dim] += dkv;
        dVs[j * kst + dim] += dvv;
      }
      __syncthreads();
    }
    // ---- write this key tile's dK, dV ---------------------------------------------------------
    for (int e = tid; e < klen * hd; e += BWD_WARPS * 32) {
      const int j = e / hd, dim = e - j * hd;
      static_cast<T*>(a.dk)[((int64_t)b * Lk + kt0 + j) * a.lddk + h * hd + dim] =
          from_f<T>(dKs[j * kst + dim] * inv_sqrt);
      static_cast<T*>(a.dv)[((int64_t)b * Lk + kt0 + j) * a.lddv + h * hd + dim] =
          from_f<T>(dVs[j * kst + dim]);
    }
    __syncthreads();
  }
  // ---- multi-tile with a separate fp32 accumulator: convert this (b,h) slice to T -------------
  if (multi_tile && a.dq_ws) {
    __threadfence_block();
    __syncthreads();
    for (int e = tid; e < Lq * hd; e += BWD_WARPS * 32) {
      const int row = e / hd, dim = e - row * hd;
      static_cast<T*>(a.dq)[((int64_t)b * Lq + row) * a.lddq + h * hd + dim] =
          from_f<T>(a.dq_ws[((int64_t)b * Lq + row) * ((int64_t)a.H * hd) + h * hd + dim]);
    }
  }
  // ---- dc ---------------------------------------------------------------------------------------
  if (a.dc && has_prev) {
    dc_part = warp_sum(dc_part);
    if (lane == 0) red[warp] = dc_part;
    __syncthreads();
    if (tid == 0) {
      float t = 0.f;
      for (int w = 0; w < BWD_WARPS; ++w) t += red[w];
      atomicAdd(a.dc, t);
    }
  }
}

constexpr size_t SMEM_LIMIT = 227 * 1024;

template <typename T>
int fwd_dispatch(const AttnArgs& a, cudaStream_t st) {
  const int kst = a.hd + 4, LkP = (a.Lk + 3) & ~3;
  const size_t smem =
      sizeof(float) * ((size_t)2 * a.Lk * kst + FWD_WARPS * R * a.hd + FWD_WARPS * R * LkP);
  if (smem > SMEM_LIMIT || a.hd > 32 * MAXD || (a.hd & 3) || a.Lk > 512) return MMEMO_ERR_SHAPE;
  dim3 grid((unsigned)cdiv(a.Lq, QTILE), (unsigned)a.H, (unsigned)a.B);
#define MM_FWD(MC)                                                                              \
  {                                                                                             \
    MM_CUDA_OK(cudaFuncSetAttribute(resattn_fwd_kernel<T, MC>,                                  \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    resattn_fwd_kernel<T, MC><<<grid, FWD_WARPS * 32, smem, st>>>(a);                           \
  }
  if (a.Lk <= 64) MM_FWD(2)
  else if (a.Lk <= 128) MM_FWD(4)
  else if (a.Lk <= 288) MM_FWD(9)
  else MM_FWD(16)
#undef MM_FWD
  MM_LAUNCH_OK();
  return MMEMO_OK;
}

template <typename T>
int bwd_dispatch(AttnArgs a, cudaStream_t st) {
  if (a.hd > 32 * MAXD || (a.hd & 3)) return MMEMO_ERR_SHAPE;
  const int kst = a.hd + 4;
  int KT = 128;
  size_t smem = 0;
  for (; KT >= 32; KT >>= 1) {
    smem = sizeof(float) * ((size_t)4 * KT * kst + 2 * CHUNK * kst + 2 * CHUNK * KT);
    if (smem <= SMEM_LIMIT) break;
  }
  if (KT < 32) return MMEMO_ERR_SHAPE;
  if (a.Lk <= 32) KT = 32; else if (a.Lk <= 64 && KT > 64) KT = 64;
  smem = sizeof(float) * ((size_t)4 * KT * kst + 2 * CHUNK * kst + 2 * CHUNK * KT);
  a.KT = KT;
  if (a.Lk > KT && sizeof(T) != sizeof(float) && a.dq_ws == nullptr) return MMEMO_ERR_ARG;
  if (sizeof(T) == sizeof(float)) a.dq_ws = nullptr;  // fp32: accumulate in dq itself
  MM_CUDA_OK(cudaFuncSetAttribute(resattn_bwd_kernel<T>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)a.H, (unsigned)a.B);
  resattn_bwd_kernel<T><<<grid, BWD_WARPS * 32, smem, st>>>(a);
  MM_LAUNCH_OK();
  return MMEMO_OK;
}

}  // namespace

// entry points used by resattn.cu (shape routing between this file and resattn_tc.cu)
int resattn_fwd_simt(int bf16_mode, const void* q, int64_t ldq, const void* k, int64_t ldk,
                     const void* v, int64_t ldv, const float* mask, int64_t mask_bs,
                     int64_t mask_rs, const void* s_prev, const float* c, void* s_out, void* o,
                     int64_t ldo, float* lse, int64_t B, int64_t H, int64_t Lq, int64_t Lk,
                     int64_t hd, cudaStream_t st) {
  if (B <= 0 || H <= 0 || Lq <= 0) return MMEMO_OK;
  MM_REQUIRE(q && k && v && o && Lk > 0 && hd > 0);
  AttnArgs a = {};
  a.q = q; a.k = k; a.v = v; a.ldq = ldq; a.ldk = ldk; a.ldv = ldv;
  a.mask = mask; a.mask_bs = mask_bs; a.mask_rs = mask_rs;
  a.s_prev = s_prev; a.c = c; a.s_out = s_out; a.o = o; a.ldo = ldo; a.lse = lse;
  a.B = (int)B; a.H = (int)H; a.Lq = (int)Lq; a.Lk = (int)Lk; a.hd = (int)hd;
  a.sqrt_hd = (float)sqrt((double)hd);
  return bf16_mode ? fwd_dispatch<bf16>(a, st) : fwd_dispatch<float>(a, st);
}

int resattn_bwd_simt(int bf16_mode, const void* d_o, int64_t lddo, const void* q, int64_t ldq,
                     const void* k, int64_t ldk, const void* v, int64_t ldv, const float* mask,
                     int64_t mask_bs, int64_t mask_rs, const void* s, const void* s_prev,
                     const float* c, const void* ds_next, const void* o, int64_t ldo,
                     const float* lse, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv,
                     int64_t lddv, void* ds_prev, float* dc, float* dq_ws, int64_t B, int64_t H,
                     int64_t Lq, int64_t Lk, int64_t hd, cudaStream_t st) {
  if (B <= 0 || H <= 0 || Lq <= 0) return MMEMO_OK;
  MM_REQUIRE(d_o && q && k && v && o && lse && dq && dk && dv && Lk > 0 && hd > 0);
  AttnArgs a = {};
  a.q = q; a.k = k; a.v = v; a.ldq = ldq; a.ldk = ldk; a.ldv = ldv;
  a.mask = mask; a.mask_bs = mask_bs; a.mask_rs = mask_rs;
  a.s_prev = s_prev; a.c = c; a.o = const_cast<void*>(o); a.ldo = ldo;
  a.lse = const_cast<float*>(lse);
  a.B = (int)B; a.H = (int)H; a.Lq = (int)Lq; a.Lk = (int)Lk; a.hd = (int)hd;
  a.sqrt_hd = (float)sqrt((double)hd);
  a.d_o = d_o; a.lddo = lddo; a.s = s; a.ds_next = ds_next;
  a.dq = dq; a.dk = dk; a.dv = dv; a.ds_prev = ds_prev;
  a.lddq = lddq; a.lddk = lddk; a.lddv = lddv; a.dc = dc; a.dq_ws = dq_ws;
  return bf16_mode ? bwd_dispatch<bf16>(a, st) : bwd_dispatch<float>(a, st);
}
