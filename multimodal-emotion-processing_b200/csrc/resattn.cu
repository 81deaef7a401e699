// Residual-attention entry points: route between the tcgen05/TMA kernels (resattn_tc.cu: bf16,
// hd = 64, L = 128; resattn_tc2.cu: hd = 64, Lk = 256, Lq = 128/256), the warp-level tensor-core kernels (resattn_mma.cu; bf16, hd = 16/32/64, any
// L, grouped) and the SIMT kernels (resattn_simt.cu; all of float32, 3-D masks, odd head sizes).
#include "common.cuh"
#include "resattn.h"

namespace {
mmemo_attn_problem make_problem(const void* q, int64_t ldq, const void* k, int64_t ldk,
                                const void* v, int64_t ldv, const float* mask, int64_t mask_bs,
                                const void* s_prev, const float* c, void* s_out, void* o,
                                int64_t ldo, float* lse, int64_t B, int64_t H, int64_t Lq,
                                int64_t Lk, int64_t hd) {
  mmemo_attn_problem a = {};
  a.q = q; a.k = k; a.v = v; a.ldq = ldq; a.ldk = ldk; a.ldv = ldv;
  a.mask = mask; a.mask_bs = mask_bs; a.s_prev = s_prev; a.c = c; a.s_out = s_out; a.lds = Lk;
  a.o = o; a.ldo = ldo; a.lse = lse; a.B = B; a.H = H; a.Lq = Lq; a.Lk = Lk; a.hd = hd;
  return a;
}
}  // namespace

extern "C" {
int mmemo_resattn_uses_mma(int64_t Lq, int64_t Lk, int64_t hd, int64_t ld, int same_kv, int bwd) {
  alignas(16) static char dummy[32];
  mmemo_attn_problem a = make_problem(dummy, ld, dummy, ld, same_kv ? dummy : dummy + 16, ld,
                                      nullptr, 0, nullptr, nullptr, nullptr, dummy, ld,
                                      reinterpret_cast<float*>(dummy), 1, 1, Lq, Lk, hd);
  a.d_o = dummy; a.lddo = ld; a.dq = dummy; a.dk = dummy; a.dv = dummy;
  a.lddq = a.lddk = a.lddv = ld;
  return resattn_mma_supported(a, bwd != 0) ? 1 : 0;
}
int mmemo_resattn_fwd_grouped_bf16(int n, const mmemo_attn_problem* ps, mmemo_stream_t s) {
  MM_REQUIRE(ps && n >= 1);
  return resattn_mma_fwd(ps, n, mm_stream(s));
}
int mmemo_resattn_bwd_grouped_bf16(int n, const mmemo_attn_problem* ps, mmemo_stream_t s) {
  MM_REQUIRE(ps && n >= 1);
  return resattn_mma_bwd(ps, n, mm_stream(s));
}

int mmemo_resattn_kernel_path(int64_t Lq, int64_t Lk, int64_t hd, int64_t ld, int bwd) {
  alignas(16) static char dummy[32];
  if (resattn_tc_supported(Lq, Lk, hd, ld, ld, ld, ld)) return 3;
  mmemo_attn_problem a = make_problem(dummy, ld, dummy, ld, dummy + 16, ld, nullptr, 0, nullptr,
                                      nullptr, nullptr, dummy, ld, reinterpret_cast<float*>(dummy),
                                      1, 1, Lq, Lk, hd);
  a.d_o = dummy; a.lddo = ld; a.dq = dummy; a.dk = dummy; a.dv = dummy;
  a.lddq = a.lddk = a.lddv = ld;
  if (resattn_tc2_supported(a, bwd != 0)) return 2;
  return resattn_mma_supported(a, bwd != 0) ? 1 : 0;
}

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
  if (mask_rs == 0 && lse && B > 0 && H > 0 && Lq > 0) {
    const mmemo_attn_problem a = make_problem(q, ldq, k, ldk, v, ldv, mask, mask_bs, s_prev, c,
                                              s_out, o, ldo, lse, B, H, Lq, Lk, hd);
    if (resattn_tc2_supported(a, false)) return resattn_fwd_tc2(a, mm_stream(s));
    if (resattn_mma_supported(a, false)) return resattn_mma_fwd(&a, 1, mm_stream(s));
  }
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
  if (mask_rs == 0 && lse && o && B > 0 && H > 0 && Lq == 128 && Lk == 128) {
    mmemo_attn_problem a = make_problem(q, ldq, k, ldk, v, ldv, mask, mask_bs, s_prev, c, nullptr,
                                        const_cast<void*>(o), ldo, const_cast<float*>(lse), B, H,
                                        Lq, Lk, hd);
    a.d_o = d_o; a.lddo = lddo; a.s = sc; a.ds_next = ds_next;
    a.dq = dq; a.dk = dk; a.dv = dv; a.lddq = lddq; a.lddk = lddk; a.lddv = lddv;
    a.ds_prev = ds_prev; a.dc = dc;
    if (resattn_tc2_supported(a, true)) return resattn_bwd_tc2(a, mm_stream(s));
  }
  if (mask_rs == 0 && d_o && q && lse &&
      resattn_tc_supported(Lq, Lk, hd, ldq, ldk, ldv, lddo) && lddq % 8 == 0 && lddk % 8 == 0 &&
      lddv % 8 == 0)
    return resattn_bwd_tc(d_o, lddo, q, ldq, k, ldk, v, ldv, mask, mask_bs, sc, s_prev, c, ds_next,
                          lse, dq, lddq, dk, lddk, dv, lddv, ds_prev, dc, B, H, mm_stream(s));
  if (mask_rs == 0 && lse && o && B > 0 && H > 0 && Lq > 0) {
    mmemo_attn_problem a = make_problem(q, ldq, k, ldk, v, ldv, mask, mask_bs, s_prev, c, nullptr,
                                        const_cast<void*>(o), ldo, const_cast<float*>(lse), B, H,
                                        Lq, Lk, hd);
    a.d_o = d_o; a.lddo = lddo; a.s = sc; a.ds_next = ds_next;
    a.dq = dq; a.dk = dk; a.dv = dv; a.lddq = lddq; a.lddk = lddk; a.lddv = lddv;
    a.ds_prev = ds_prev; a.dc = dc;
    if (resattn_tc2_supported(a, true)) return resattn_bwd_tc2(a, mm_stream(s));
    if (resattn_mma_supported(a, true)) return resattn_mma_bwd(&a, 1, mm_stream(s));
  }
  return resattn_bwd_simt(1, d_o, lddo, q, ldq, k, ldk, v, ldv, mask, mask_bs, mask_rs, sc, s_prev,
                          c, ds_next, o, ldo, lse, dq, lddq, dk, lddk, dv, lddv, ds_prev, dc, dq_ws,
                          B, H, Lq, Lk, hd, mm_stream(s));
}
}
