// Device-side batch assembly: ragged per-utterance feature sequences -> the padded
// (N, m_len, D) float32 batch + (N, m_len) 0/1 mask the models take, with the reference's NaN/Inf
// scrub.  Replaces the per-sample numpy code of the reference data loaders:
//   tail   : others/realformer.py:72-82   masking(features[-m_len:], m_len)  (keep the LAST m_len rows)
//   head   : keep the first m_len rows (the first view of cmu-mosei/run.py:139, `m[:m_len-3]`)
//   stride : robot_demo.py:86-99,115-150  T >= m_len -> rows 0, gap, 2*gap, ... with gap = T // m_len
//            (video, audio and text features of the demo)
// Rows beyond the kept ones are zero, mask = 1 on kept rows; a sequence of length 0 (the
// 'no_name' context slots, others/realformer.py:108-113) gives an all-zero sample and mask.
// HBM-bound gather: one CTA per (sample, block of rows), 16-byte accesses when D % 4 == 0.
#include <math.h>

#include "common.cuh"

namespace {

__device__ __forceinline__ float scrub(float x, int do_scrub, float value) {
  return (do_scrub && (isnan(x) || isinf(x))) ? value : x;
}

__global__ void __launch_bounds__(256)
assemble_kernel(const float* __restrict__ flat, const int64_t* __restrict__ row_start,
                const int64_t* __restrict__ n_rows, float* __restrict__ out,
                float* __restrict__ mask, int m_len, int D, int mode, int do_scrub,
                float scrub_value, int rows_per_cta) {
  const int n = blockIdx.y;
  const int64_t T = n_rows[n];
  const int64_t start = row_start[n];
  const int64_t keep = T < m_len ? T : m_len;
  const int64_t gap = (mode == 2 && T >= m_len && m_len > 0) ? T / m_len : 1;
  const int64_t first = mode == 0 ? T - keep : 0;
  const int t0 = blockIdx.x * rows_per_cta;
  const int t1 = min(m_len, t0 + rows_per_cta);
  if (mask)
    for (int t = t0 + threadIdx.x; t < t1; t += 256)
      mask[(int64_t)n * m_len + t] = t < keep ? 1.f : 0.f;
  float* dst = out + ((int64_t)n * m_len + t0) * D;
  const int64_t total = (int64_t)(t1 - t0) * D;
  if ((D & 3) == 0 && (reinterpret_cast<uintptr_t>(flat) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    const int D4 = D >> 2;
    for (int64_t i = threadIdx.x; i < total / 4; i += 256) {
      const int t = t0 + (int)(i / D4), c = (int)(i % D4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t < keep) {
        v = reinterpret_cast<const float4*>(flat + (start + first + (int64_t)t * gap) * D)[c];
        v.x = scrub(v.x, do_scrub, scrub_value); v.y = scrub(v.y, do_scrub, scrub_value);
        v.z = scrub(v.z, do_scrub, scrub_value); v.w = scrub(v.w, do_scrub, scrub_value);
      }
      reinterpret_cast<float4*>(dst)[i] = v;
    }
  } else {
    for (int64_t i = threadIdx.x; i < total; i += 256) {
      const int t = t0 + (int)(i / D), c = (int)(i % D);
      float v = 0.f;
      if (t < keep) v = scrub(flat[(start + first + (int64_t)t * gap) * D + c], do_scrub, scrub_value);
      dst[i] = v;
    }
  }
}

// cmu-mosei/run.py:104-151 (masking, non-BERT branch): every sample is prefixed by three
// STATISTICS rows (column-wise max, min, mean over ALL its rows, after the audio NaN/Inf scrub) and
// followed by m_len-3 body rows.  A sequence with T >= m_len-3 rows yields two views: the head
// (rows 0 .. m_len-4) and the tail (the last m_len-3 rows); a shorter one yields one zero-padded
// view with mask = 1 on its T+3 rows.  Thread = feature column: the T rows of a column are walked
// once for the statistics (coalesced across the columns), then the body rows are copied.
__global__ void __launch_bounds__(128)
assemble_stats_kernel(const float* __restrict__ flat, const int64_t* __restrict__ row_start,
                      const int64_t* __restrict__ n_rows, float* __restrict__ out,
                      float* __restrict__ mask, int m_len, int D, int view, int do_scrub,
                      float scrub_value) {
  const int n = blockIdx.y;
  const int64_t T = n_rows[n];
  const int64_t start = row_start[n];
  const int body = m_len - 3;
  const bool two = T >= body;
  if (mask && blockIdx.x == 0)
    for (int t = threadIdx.x; t < m_len; t += 128)
      mask[(int64_t)n * m_len + t] = (T > 0 && (two || t < T + 3)) ? 1.f : 0.f;
  const int c = blockIdx.x * 128 + threadIdx.x;
  if (c >= D) return;
  float* dst = out + (int64_t)n * m_len * D + c;
  if (T <= 0) {                       // 'no_name' slot (cmu-mosei/run.py:162-168): all zeros
    for (int t = 0; t < m_len; ++t) dst[(int64_t)t * D] = 0.f;
    return;
  }
  const float* src = flat + start * D + c;
  float mx = -INFINITY, mn = INFINITY;
  double sum = 0.0;
  for (int64_t t = 0; t < T; ++t) {
    const float v = scrub(src[t * D], do_scrub, scrub_value);
    mx = fmaxf(mx, v);
    mn = fminf(mn, v);
    sum += (double)v;
  }
  if (m_len > 0) dst[0] = mx;
  if (m_len > 1) dst[(int64_t)D] = mn;
  if (m_len > 2) dst[2 * (int64_t)D] = (float)(sum / (double)T);
  const int64_t first = (view == 1 && two) ? T - body : 0;
  for (int r = 0; r < body; ++r) {
    const int64_t t = first + r;
    dst[(int64_t)(3 + r) * D] = t < T ? scrub(src[t * D], do_scrub, scrub_value) : 0.f;
  }
}

}  // namespace

extern "C" int mmemo_assemble_stats_batch_f32(const float* flat, const int64_t* row_start,
                                              const int64_t* n_rows, float* out, float* mask,
                                              int64_t N, int64_t m_len, int64_t D, int view,
                                              int do_scrub, float scrub_value, mmemo_stream_t s) {
  if (N <= 0 || m_len <= 0 || D <= 0) return N == 0 || m_len == 0 || D == 0 ? MMEMO_OK : MMEMO_ERR_ARG;
  MM_REQUIRE(flat && row_start && n_rows && out && (view == 0 || view == 1) && m_len >= 3);
  if (N > 65535 || m_len > (1 << 20) || D > (1 << 20)) return MMEMO_ERR_SHAPE;
  assemble_stats_kernel<<<dim3((unsigned)cdiv(D, 128), (unsigned)N), 128, 0, mm_stream(s)>>>(
      flat, row_start, n_rows, out, mask, (int)m_len, (int)D, view, do_scrub, scrub_value);
  MM_LAUNCH_OK();
  return MMEMO_OK;
}

extern "C" int mmemo_assemble_batch_f32(const float* flat, const int64_t* row_start,
                                        const int64_t* n_rows, float* out, float* mask, int64_t N,
                                        int64_t m_len, int64_t D, int mode, int do_scrub,
                                        float scrub_value, mmemo_stream_t s) {
  if (N <= 0 || m_len <= 0 || D <= 0) return N == 0 || m_len == 0 || D == 0 ? MMEMO_OK : MMEMO_ERR_ARG;
  MM_REQUIRE(flat && row_start && n_rows && out && mode >= 0 && mode <= 2);
  if (N > 65535 || m_len > (1 << 20) || D > (1 << 20)) return MMEMO_ERR_SHAPE;
  // ~16 KB of output per CTA
  int64_t rows_per_cta = cdiv(4096, D);
  if (rows_per_cta > m_len) rows_per_cta = m_len;
  dim3 grid((unsigned)cdiv(m_len, rows_per_cta), (unsigned)N);
  assemble_kernel<<<grid, 256, 0, mm_stream(s)>>>(flat, row_start, n_rows, out, mask, (int)m_len,
                                                  (int)D, mode, do_scrub, scrub_value,
                                                  (int)rows_per_cta);
  MM_LAUNCH_OK();
  return MMEMO_OK;
}
