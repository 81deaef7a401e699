// Internal interface between resattn.cu (C ABI), resattn_simt.cu and resattn_tc.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mmemo.h"

int resattn_fwd_simt(int bf16_mode, const void* q, int64_t ldq, const void* k, int64_t ldk,
                     const void* v, int64_t ldv, const float* mask, int64_t mask_bs,
                     int64_t mask_rs, const void* s_prev, const float* c, void* s_out, void* o,
                     int64_t ldo, float* lse, int64_t B, int64_t H, int64_t Lq, int64_t Lk,
                     int64_t hd, cudaStream_t st);
int resattn_bwd_simt(int bf16_mode, const void* d_o, int64_t lddo, const void* q, int64_t ldq,
                     const void* k, int64_t ldk, const void* v, int64_t ldv, const float* mask,
                     int64_t mask_bs, int64_t mask_rs, const void* s, const void* s_prev,
                     const float* c, const void* ds_next, const void* o, int64_t ldo,
                     const float* lse, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv,
                     int64_t lddv, void* ds_prev, float* dc, float* dq_ws, int64_t B, int64_t H,
                     int64_t Lq, int64_t Lk, int64_t hd, cudaStream_t st);

// tcgen05 / TMA path (bf16, hd == 64, Lq == Lk == 128)
bool resattn_tc_supported(int64_t Lq, int64_t Lk, int64_t hd, int64_t ldq, int64_t ldk,
                          int64_t ldv, int64_t ldo);
int resattn_bwd_tc(const void* d_o, int64_t lddo, const void* q, int64_t ldq, const void* k,
                   int64_t ldk, const void* v, int64_t ldv, const float* mask, int64_t mask_bs,
                   const void* s, const void* s_prev, const float* c, const void* ds_next,
                   const float* lse, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv,
                   int64_t lddv, void* ds_prev, float* dc, int64_t B, int64_t H, cudaStream_t st);
int resattn_fwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                   int64_t ldv, const float* mask, int64_t mask_bs, const void* s_prev,
                   const float* c, void* s_out, void* o, int64_t ldo, float* lse, int64_t B,
                   int64_t H, int64_t Lq, int64_t Lk, int64_t hd, cudaStream_t st);

// tcgen05 / TMA path for the long-sequence shapes (bf16, hd == 64, Lk == 256, Lq in {128, 256})
bool resattn_tc2_supported(const mmemo_attn_problem& a, bool bwd);
int resattn_fwd_tc2(const mmemo_attn_problem& a, cudaStream_t st);
int resattn_bwd_tc2(const mmemo_attn_problem& a, cudaStream_t st);
// L2 prefetch distance (CTAs) for the one-CTA-per-SM attention kernels: the SM count, or MMEMO_ATTN_PF
int resattn_pf_distance();

// mma.sync (m16n8k16 bf16) path: hd in {16, 32, 64}, any L that fits shared memory; grouped
bool resattn_mma_supported(const mmemo_attn_problem& a, bool bwd);
int resattn_mma_fwd(const mmemo_attn_problem* ps, int n, cudaStream_t st);
int resattn_mma_bwd(const mmemo_attn_problem* ps, int n, cudaStream_t st);
