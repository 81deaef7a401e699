// Residual-attention entry points: route between the tcgen05/TMA kernels (resattn_tc.cu; bf16,
// hd = 64, L = 128) and the SIMT kernels (resattn_simt.cu; everything else, all of float32).
#include "common.cuh"
#include "resattn.h"

extern "C" {
int mmemo_resattn_uses_tensor_cores(int64_t Lq, int64_t Lk, int64_t hd, int64_t ld) {
  return resattn_tc_supported(Lq, Lk, hd, ld, ld, ld, ld) ? 1 : 0;
}

int mmemo_resattn_fwd_f32(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                          int64_t ldv, const float* mask, int64_t mask_bs, int64_t mask_rs,
                          const void* s_prev, const float* c, void* s_out, void* o, int64_t ldo,
                          float* lse, int64_t B, int64_t H, int64_t Lq, int64_t Lk, int64_t hd,
                          mmemo_stream_t s) {
  return resattn_fwd_simt(0, q, ldq, k, ldk, v, ldv, mask, mask_bs, mask_rs, s_prev, c, s_out, o,
                          ldo, lse, B, H, Lq, Lk, hd, mm_stream(s));
}
int mmemo_resattn_fwd_bf16(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                           int64_t ldv, const float* mask, int64_t mask_bs, int64_t mask_rs,
                           const void* s_prev, const float* c, void* s_out, void* o, int64_t ldo,
                           float* lse, int64_t B, int64_t H, int64_t Lq, int64_t Lk, int64_t hd,
                           mmemo_stream_t s) {
  if (mask_rs == 0 && lse && resattn_tc_supported(Lq, Lk, hd, ldq, ldk, ldv, ldo))
    return resattn_fwd_tc(q, ldq, k, ldk, v, ldv, mask, mask_bs, s_prev, c, s_out, o, ldo, lse, B,
                          H, Lq, Lk, hd, mm_stream(s));
  return resattn_fwd_simt(1, q, ldq, k, ldk, v, ldv, mask, mask_bs, mask_rs, s_prev, c, s_out, o,
                          ldo, lse, B, H, Lq, Lk, hd, mm_stream(s));
}
int mmemo_resattn_bwd_f32(const void* d_o, int64_t lddo, const void* q, int64_t ldq, const void* k,
                          int64_t ldk, const void* v, int64_t ldv, const float* mask,
                          int64_t mask_bs, int64_t mask_rs, const void* sc, const void* s_prev,
                          const float* c, const void* ds_next, const void* o, int64_t ldo,
                          const float* lse, void* dq, int64_t lddq, void* dk, int64_t lddk,
                          void* dv, int64_t lddv, void* ds_prev, float* dc, float* dq_ws, int64_t B,
                          int64_t H, int64_t Lq, int64_t Lk, int64_t hd, mmemo_stream_t s) {
  return resattn_bwd_simt(0, d_o, lddo, q, ldq, k, ldk, v, ldv, mask, mask_bs, mask_rs, sc, s_prev,
                          c, ds_next, o, ldo, lse, dq, lddq, dk, lddk, dv, lddv, ds_prev, dc, dq_ws,
                          B, H, Lq, Lk, hd, mm_stream(s));
}
int mmemo_resattn_bwd_bf16(const void* d_o, int64_t lddo, const void* q, int64_t ldq,
                           const void* k, int64_t ldk, const void* v, int64_t ldv,
                           const float* mask, int64_t mask_bs, int64_t mask_rs, const void* sc,
                           const void* s_prev, const float* c, const void* ds_next, const void* o,
                           int64_t ldo, const float* lse, void* dq, int64_t lddq, void* dk,
                           int64_t lddk, void* dv, int64_t lddv, void* ds_prev, float* dc,
                           float* dq_ws, int64_t B, int64_t H, int64_t Lq, int64_t Lk, int64_t hd,
                           mmemo_stream_t s) {
  if (mask_rs == 0 && d_o && q && lse &&
      resattn_tc_supported(Lq, Lk, hd, ldq, ldk, ldv, lddo) && lddq % 8 == 0 && lddk % 8 == 0 &&
      lddv % 8 == 0)
    return resattn_bwd_tc(d_o, lddo, q, ldq, k, ldk, v, ldv, mask, mask_bs, sc, s_prev, c, ds_next,
                          lse, dq, lddq, dk, lddk, dv, lddv, ds_prev, dc, B, H, mm_stream(s));
  return resattn_bwd_simt(1, d_o, lddo, q, ldq, k, ldk, v, ldv, mask, mask_bs, mask_rs, sc, s_prev,
                          c, ds_next, o, ldo, lse, dq, lddq, dk, lddk, dv, lddv, ds_prev, dc, dq_ws,
                          B, H, Lq, Lk, hd, mm_stream(s));
}
}
