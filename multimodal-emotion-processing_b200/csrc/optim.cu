// Fused gradient clipping + Adam / AdamW step over a list of tensors.
// Replaces  nn.utils.clip_grad_norm_(model.parameters(), CLIP); optimizer.step()
//   others/realformer.py:314-315,342 (Adam)   cmu-mosei/run.py:368-369,398   Ren-MME/run.py:336-337,379
//   rencecps/run.py:175-176,202   robot_demo.py:471-472,502 (AdamW)
// which torch runs as ~5 multi-tensor launches for the norm/clip and ~10 for the update over
// 100-280 small tensors.  Here: one launch for the global squared norm, one for the update (the
// clip coefficient is computed on the device from the norm, so there is no host sync), per chunk
// of 96 tensors.  HBM-bound: reads p, g, m, v and writes p, m, v once (28 B per parameter).
#include "common.cuh"

namespace {

constexpr int OPT_MAXT = 96;

struct NormTable {
  const float* g[OPT_MAXT];
  long long n[OPT_MAXT];
  int count;
};

struct AdamTable {
  float* p[OPT_MAXT];
  float* g[OPT_MAXT];
  float* m[OPT_MAXT];
  float* v[OPT_MAXT];
  long long n[OPT_MAXT];
  int count;
};

__device__ __forceinline__ float clip_coef(const float* sqnorm, float max_norm) {
  // torch.nn.utils.clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to 1
  if (!sqnorm) return 1.f;
  const float c = max_norm / (sqrtf(sqnorm[0]) + 1e-6f);
  return c < 1.f ? c : 1.f;
}

__global__ void __launch_bounds__(256) sqnorm_multi_kernel(NormTable t, float* __restrict__ out) {
  __shared__ float red[8];
  const int which = blockIdx.y;
  const float* __restrict__ g = t.g[which];
  const long long n = t.n[which];
  float acc = 0.f;
  const bool vec = (reinterpret_cast<uintptr_t>(g) & 15) == 0;
  if (vec) {
    const long long n4 = n / 4;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4;
         i += (long long)gridDim.x * 256) {
      const float4 x = reinterpret_cast<const float4*>(g)[i];
      acc = fmaf(x.x, x.x, acc); acc = fmaf(x.y, x.y, acc);
      acc = fmaf(x.z, x.z, acc); acc = fmaf(x.w, x.w, acc);
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n - n4 * 4)) {
      const float x = g[n4 * 4 + threadIdx.x];
      acc = fmaf(x, x, acc);
    }
  } else {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n;
         i += (long long)gridDim.x * 256)
      acc = fmaf(g[i], g[i], acc);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float s = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0 && s != 0.f) atomicAdd(out, s);
  }
}

__global__ void __launch_bounds__(256)
scale_multi_kernel(AdamTable t, const float* __restrict__ sqnorm, float max_norm) {
  const float coef = clip_coef(sqnorm, max_norm);
  if (coef == 1.f) return;
  const int which = blockIdx.y;
  float* __restrict__ g = t.g[which];
  const long long n = t.n[which];
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n;
       i += (long long)gridDim.x * 256)
    g[i] *= coef;
}

struct AdamHyper {
  float lr, beta1, beta2, eps, weight_decay;
  float step_size;      // lr / (1 - beta1^t)
  float inv_bc2_sqrt;   // 1 / sqrt(1 - beta2^t)
  float max_norm;
  int decoupled;        // 1: AdamW (p *= 1 - lr*wd), 0: Adam (g += wd*p)
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamHyper& h,
                                         float coef) {
  g *= coef;
  if (h.decoupled) p *= 1.f - h.lr * h.weight_decay;
  else if (h.weight_decay != 0.f) g = fmaf(h.weight_decay, p, g);
  m = fmaf(1.f - h.beta1, g - m, m);                    // lerp(m, g, 1-beta1)
  v = fmaf((1.f - h.beta2) * g, g, v * h.beta2);        // v*beta2 + (1-beta2)*g*g
  const float denom = sqrtf(v) * h.inv_bc2_sqrt + h.eps;
  p -= h.step_size * (m / denom);
}

__global__ void __launch_bounds__(256)
adam_multi_kernel(AdamTable t, AdamHyper h, const float* __restrict__ sqnorm) {
  const int which = blockIdx.y;
  float* __restrict__ p = t.p[which];
  const float* __restrict__ g = t.g[which];
  float* __restrict__ m = t.m[which];
  float* __restrict__ v = t.v[which];
  const long long n = t.n[which];
  const float coef = clip_coef(sqnorm, h.max_norm);
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                     reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  if (vec) {
    const long long n4 = n / 4;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4;
         i += (long long)gridDim.x * 256) {
      float4 pp = reinterpret_cast<float4*>(p)[i];
      const float4 gg = reinterpret_cast<const float4*>(g)[i];
      float4 mm = reinterpret_cast<float4*>(m)[i];
      float4 vv = reinterpret_cast<float4*>(v)[i];
      adam_one(pp.x, gg.x, mm.x, vv.x, h, coef);
      adam_one(pp.y, gg.y, mm.y, vv.y, h, coef);
      adam_one(pp.z, gg.z, mm.z, vv.z, h, coef);
      adam_one(pp.w, gg.w, mm.w, vv.w, h, coef);
      reinterpret_cast<float4*>(p)[i] = pp;
      reinterpret_cast<float4*>(m)[i] = mm;
      reinterpret_cast<float4*>(v)[i] = vv;
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n - n4 * 4)) {
      const long long i = n4 * 4 + threadIdx.x;
      adam_one(p[i], g[i], m[i], v[i], h, coef);
    }
  } else {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n;
         i += (long long)gridDim.x * 256)
      adam_one(p[i], g[i], m[i], v[i], h, coef);
  }
}

unsigned grid_x(long long nmax) {
  long long bx = cdiv(cdiv(nmax, 4), 256);
  if (bx > 148 * 2) bx = 148 * 2;
  return (unsigned)(bx < 1 ? 1 : bx);
}

}  // namespace

extern "C" {

int mmemo_grad_sqnorm_f32(int count, const float* const* grads, const int64_t* numel,
                          float* sqnorm_out, mmemo_stream_t s) {
  MM_REQUIRE(count >= 0 && sqnorm_out && (count == 0 || (grads && numel)));
  cudaStream_t st = mm_stream(s);
  MM_CUDA_OK(cudaMemsetAsync(sqnorm_out, 0, sizeof(float), st));
  for (int base = 0; base < count; base += OPT_MAXT) {
    NormTable t = {};
    long long nmax = 0;
    t.count = count - base < OPT_MAXT ? count - base : OPT_MAXT;
    for (int i = 0; i < t.count; ++i) {
      MM_REQUIRE(grads[base + i] && numel[base + i] >= 0);
      t.g[i] = grads[base + i];
      t.n[i] = numel[base + i];
      nmax = t.n[i] > nmax ? t.n[i] : nmax;
    }
    sqnorm_multi_kernel<<<dim3(grid_x(nmax), (unsigned)t.count), 256, 0, st>>>(t, sqnorm_out);
    MM_LAUNCH_OK();
  }
  return MMEMO_OK;
}

int mmemo_clip_grads_f32(int count, float* const* grads, const int64_t* numel, const float* sqnorm,
                         float max_norm, mmemo_stream_t s) {
  MM_REQUIRE(count >= 0 && sqnorm && max_norm > 0.f && (count == 0 || (grads && numel)));
  cudaStream_t st = mm_stream(s);
  for (int base = 0; base < count; base += OPT_MAXT) {
    AdamTable t = {};
    long long nmax = 0;
    t.count = count - base < OPT_MAXT ? count - base : OPT_MAXT;
    for (int i = 0; i < t.count; ++i) {
      MM_REQUIRE(grads[base + i]);
      t.g[i] = grads[base + i];
      t.n[i] = numel[base + i];
      nmax = t.n[i] > nmax ? t.n[i] : nmax;
    }
    scale_multi_kernel<<<dim3(grid_x(nmax * 4), (unsigned)t.count), 256, 0, st>>>(t, sqnorm, max_norm);
    MM_LAUNCH_OK();
  }
  return MMEMO_OK;
}

int mmemo_adam_step_f32(int count, float* const* params, const float* const* grads,
                        float* const* exp_avg, float* const* exp_avg_sq, const int64_t* numel,
                        float lr, float beta1, float beta2, float eps, float weight_decay,
                        int decoupled, int64_t step, const float* sqnorm, float max_norm,
                        mmemo_stream_t s) {
  MM_REQUIRE(count >= 0 && step >= 1 && (count == 0 || (params && grads && exp_avg && exp_avg_sq && numel)));
  MM_REQUIRE(!sqnorm || max_norm > 0.f);
  cudaStream_t st = mm_stream(s);
  AdamHyper h = {};
  h.lr = lr; h.beta1 = beta1; h.beta2 = beta2; h.eps = eps; h.weight_decay = weight_decay;
  // bias corrections in double like torch's Python scalars (1 - beta ** step)
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  h.step_size = (float)((double)lr / bc1);
  h.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  h.max_norm = max_norm;
  h.decoupled = decoupled;
  for (int base = 0; base < count; base += OPT_MAXT) {
    AdamTable t = {};
    long long nmax = 0;
    t.count = count - base < OPT_MAXT ? count - base : OPT_MAXT;
    for (int i = 0; i < t.count; ++i) {
      const int j = base + i;
      MM_REQUIRE(params[j] && grads[j] && exp_avg[j] && exp_avg_sq[j] && numel[j] >= 0);
      t.p[i] = params[j];
      t.g[i] = const_cast<float*>(grads[j]);
      t.m[i] = exp_avg[j];
      t.v[i] = exp_avg_sq[j];
      t.n[i] = numel[j];
      nmax = t.n[i] > nmax ? t.n[i] : nmax;
    }
    adam_multi_kernel<<<dim3(grid_x(nmax), (unsigned)t.count), 256, 0, st>>>(t, h, sqnorm);
    MM_LAUNCH_OK();
  }
  return MMEMO_OK;
}

}  // extern "C"
