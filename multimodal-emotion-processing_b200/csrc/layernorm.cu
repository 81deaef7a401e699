// Gated residual + LayerNorm (+ReLU), forward and backward.  One warp per row; the row lives in
// registers (lane owns columns lane, lane+32, ...), statistics by warp shuffles, two-pass variance.
// HBM-bound: fwd reads res,x and writes y once; bwd reads dy,res,x(,y) once and writes dres,dx.
//   y = act( LN(res + gate*x) * gamma + beta )
// Replaces others/realformer.py:207-208,263  cmu-mosei/run.py:261  Ren-MME/run.py:166,213.
#include "common.cuh"

namespace {

// NPL = columns per lane (template): d <= 32*NPL.  fwd: NPL<=32 (d<=1024); bwd keeps 7 row
// arrays in registers, so NPL<=16 (d<=512).
constexpr int LN_WARPS = 8;

template <typename T, int NPL>
__global__ void __launch_bounds__(LN_WARPS * 32)
add_ln_fwd_kernel(const T* __restrict__ res, int64_t ldres, const T* __restrict__ x, int64_t ldx,
                  const float* __restrict__ gate, const float* __restrict__ gamma,
                  const float* __restrict__ beta, T* __restrict__ y, int64_t ldy,
                  float* __restrict__ mean_out, float* __restrict__ rstd_out, int64_t M, int d,
                  float eps, int relu) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  if (row >= M) return;
  const float g = gate ? gate[0] : 1.f;
  const int npl = (d + 31) >> 5;
  float z[NPL];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    z[i] = 0.f;
    const int col = lane + 32 * i;
    if (i < npl && col < d) {
      float v = g * to_f(x[row * ldx + col]);
      if (res) v += to_f(res[row * ldres + col]);
      z[i] = v;
      sum += v;
    }
  }
  const float mean = warp_sum(sum) / (float)d;
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int col = lane + 32 * i;
    if (i < npl && col < d) {
      const float t = z[i] - mean;
      var = fmaf(t, t, var);
    }
  }
  const float rstd = rsqrtf(warp_sum(var) / (float)d + eps);
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int col = lane + 32 * i;
    if (i < npl && col < d) {
      float v = (z[i] - mean) * rstd * gamma[col] + beta[col];
      if (relu) v = fmaxf(v, 0.f);
      y[row * ldy + col] = from_f<T>(v);
    }
  }
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
}

template <typename T, int NPL>
__global__ void __launch_bounds__(LN_WARPS * 32)
add_ln_bwd_kernel(const T* __restrict__ dy, int64_t lddy, const T* __restrict__ res, int64_t ldres,
                  const T* __restrict__ x, int64_t ldx, const float* __restrict__ gate,
                  const float* __restrict__ gamma, const T* __restrict__ y, int64_t ldy,
                  const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                  T* __restrict__ dres, int64_t lddres, T* __restrict__ dx, int64_t lddx,
                  float* __restrict__ dgate, float* __restrict__ dgamma, float* __restrict__ dbeta,
                  int64_t M, int d, int relu) {
  extern __shared__ float red[];  // [LN_WARPS][2*d]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float g = gate ? gate[0] : 1.f;
  const int npl = (d + 31) >> 5;
  float gam[NPL], dgam[NPL], dbet[NPL];
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int col = lane + 32 * i;
    gam[i] = (i < npl && col < d) ? gamma[col] : 0.f;
    dgam[i] = 0.f;
    dbet[i] = 0.f;
  }
  float dg = 0.f;
  for (int64_t row = (int64_t)blockIdx.x * LN_WARPS + warp; row < M;
       row += (int64_t)gridDim.x * LN_WARPS) {
    const float mean = mean_in[row], rstd = rstd_in[row];
    float xh[NPL], w[NPL], xv[NPL];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      xh[i] = 0.f; w[i] = 0.f; xv[i] = 0.f;
      const int col = lane + 32 * i;
      if (i < npl && col < d) {
        xv[i] = to_f(x[row * ldx + col]);
        float z = g * xv[i];
        if (res) z += to_f(res[row * ldres + col]);
        xh[i] = (z - mean) * rstd;
        float dyv = to_f(dy[row * lddy + col]);
        if (relu && !(to_f(y[row * ldy + col]) > 0.f)) dyv = 0.f;
        dgam[i] = fmaf(dyv, xh[i], dgam[i]);
        dbet[i] += dyv;
        w[i] = dyv * gam[i];
        s1 += w[i];
        s2 = fmaf(w[i], xh[i], s2);
      }
    }
    s1 = warp_sum(s1) / (float)d;
    s2 = warp_sum(s2) / (float)d;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int col = lane + 32 * i;
      if (i < npl && col < d) {
        const float dz = rstd * (w[i] - s1 - xh[i] * s2);
        if (dres) dres[row * lddres + col] = from_f<T>(dz);
        dx[row * lddx + col] = from_f<T>(g * dz);
        dg = fmaf(dz, xv[i], dg);
      }
    }
  }
  // block reduction of the parameter gradients, then one atomic per column per CTA
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int col = lane + 32 * i;
    if (i < npl && col < d) {
      red[warp * 2 * d + col] = dgam[i];
      red[warp * 2 * d + d + col] = dbet[i];
    }
  }
  dg = warp_sum(dg);
  __shared__ float dgs[LN_WARPS];
  if (lane == 0) dgs[warp] = dg;
  __syncthreads();
  for (int col = threadIdx.x; col < 2 * d; col += LN_WARPS * 32) {
    float t = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < LN_WARPS; ++w2) t += red[w2 * 2 * d + col];
    if (col < d) { if (dgamma) atomicAdd(dgamma + col, t); }
    else { if (dbeta) atomicAdd(dbeta + col - d, t); }
  }
  if (threadIdx.x == 0 && dgate && gate) {
    float t = 0.f;
    for (int w2 = 0; w2 < LN_WARPS; ++w2) t += dgs[w2];
    atomicAdd(dgate, t);
  }
}

template <typename T>
int fwd(const void* res, int64_t ldres, const void* x, int64_t ldx, const float* gate,
        const float* gamma, const float* beta, void* y, int64_t ldy, float* mean, float* rstd,
        int64_t M, int64_t d, float eps, int relu, cudaStream_t st) {
  if (M <= 0) return MMEMO_OK;
  MM_REQUIRE(x && gamma && beta && y && d > 0);
  if (d > 1024) return MMEMO_ERR_SHAPE;
#define MM_LN_FWD(N_)                                                                         \
  add_ln_fwd_kernel<T, N_><<<(unsigned)cdiv(M, LN_WARPS), LN_WARPS * 32, 0, st>>>(            \
      static_cast<const T*>(res), ldres, static_cast<const T*>(x), ldx, gate, gamma, beta,    \
      static_cast<T*>(y), ldy, mean, rstd, M, (int)d, eps, relu)
  if (d <= 128) MM_LN_FWD(4);
  else if (d <= 256) MM_LN_FWD(8);
  else if (d <= 512) MM_LN_FWD(16);
  else MM_LN_FWD(32);
#undef MM_LN_FWD
  MM_LAUNCH_OK();
  return MMEMO_OK;
}

template <typename T>
int bwd(const void* dy, int64_t lddy, const void* res, int64_t ldres, const void* x, int64_t ldx,
        const float* gate, const float* gamma, const void* y, int64_t ldy, const float* mean,
        const float* rstd, void* dres, int64_t lddres, void* dx, int64_t lddx, float* dgate,
        float* dgamma, float* dbeta, int64_t M, int64_t d, int relu, cudaStream_t st) {
  if (M <= 0) return MMEMO_OK;
  MM_REQUIRE(dy && x && gamma && mean && rstd && dx && d > 0 && (!relu || y));
  if (d > 512) return MMEMO_ERR_SHAPE;
  int64_t blocks = cdiv(M, LN_WARPS);
  if (blocks > 148 * 2) blocks = 148 * 2;   // few CTAs: one atomic per column per CTA at the end
  const size_t smem = sizeof(float) * LN_WARPS * 2 * d;
#define MM_LN_BWD(N_)                                                                          \
  add_ln_bwd_kernel<T, N_><<<(unsigned)blocks, LN_WARPS * 32, smem, st>>>(                     \
      static_cast<const T*>(dy), lddy, static_cast<const T*>(res), ldres,                      \
      static_cast<const T*>(x), ldx, gate, gamma, static_cast<const T*>(y), ldy, mean, rstd,   \
      static_cast<T*>(dres), lddres, static_cast<T*>(dx), lddx, dgate, dgamma, dbeta, M,       \
      (int)d, relu)
  if (d <= 128) MM_LN_BWD(4);
  else if (d <= 256) MM_LN_BWD(8);
  else MM_LN_BWD(16);
#undef MM_LN_BWD
  MM_LAUNCH_OK();
  return MMEMO_OK;
}

}  // namespace

extern "C" {
int mmemo_add_ln_fwd_f32(const void* res, int64_t ldres, const void* x, int64_t ldx,
                         const float* gate, const float* gamma, const float* beta, void* y,
                         int64_t ldy, float* mean, float* rstd, int64_t M, int64_t d, float eps,
                         int relu, mmemo_stream_t s) {
  return fwd<float>(res, ldres, x, ldx, gate, gamma, beta, y, ldy, mean, rstd, M, d, eps, relu,
                    mm_stream(s));
}
int mmemo_add_ln_fwd_bf16(const void* res, int64_t ldres, const void* x, int64_t ldx,
                          const float* gate, const float* gamma, const float* beta, void* y,
                          int64_t ldy, float* mean, float* rstd, int64_t M, int64_t d, float eps,
                          int relu, mmemo_stream_t s) {
  return fwd<bf16>(res, ldres, x, ldx, gate, gamma, beta, y, ldy, mean, rstd, M, d, eps, relu,
                   mm_stream(s));
}
int mmemo_add_ln_bwd_f32(const void* dy, int64_t lddy, const void* res, int64_t ldres,
                         const void* x, int64_t ldx, const float* gate, const float* gamma,
                         const void* y, int64_t ldy, const float* mean, const float* rstd,
                         void* dres, int64_t lddres, void* dx, int64_t lddx, float* dgate,
                         float* dgamma, float* dbeta, int64_t M, int64_t d, int relu,
                         mmemo_stream_t s) {
  return bwd<float>(dy, lddy, res, ldres, x, ldx, gate, gamma, y, ldy, mean, rstd, dres, lddres,
                    dx, lddx, dgate, dgamma, dbeta, M, d, relu, mm_stream(s));
}
int mmemo_add_ln_bwd_bf16(const void* dy, int64_t lddy, const void* res, int64_t ldres,
                          const void* x, int64_t ldx, const float* gate, const float* gamma,
                          const void* y, int64_t ldy, const float* mean, const float* rstd,
                          void* dres, int64_t lddres, void* dx, int64_t lddx, float* dgate,
                          float* dgamma, float* dbeta, int64_t M, int64_t d, int relu,
                          mmemo_stream_t s) {
  return bwd<bf16>(dy, lddy, res, ldres, x, ldx, gate, gamma, y, ldy, mean, rstd, dres, lddres,
                   dx, lddx, dgate, dgamma, dbeta, M, d, relu, mm_stream(s));
}
}
