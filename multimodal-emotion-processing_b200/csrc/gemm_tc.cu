// Persistent tcgen05 / TMA GEMM for sm_100a:  C[M,N] (+)= epi( A * B^T ), bf16 operands, fp32
// accumulation in TMEM.  One CTA per SM loops over 128 x 128 output tiles (x split-K slices).
//
//   warp 0   TMA producer : cp.async.bulk.tensor (SWIZZLE_128B) into a 4-stage shared-memory ring
//   warp 1   MMA issuer   : one thread issues tcgen05.mma (UMMA 128x128x16, cta_group::1) into one
//                           of TWO 128-column TMEM accumulators; tcgen05.commit frees ring slots and
//                           hands the finished accumulator to the epilogue
//   warps 2-5 epilogue    : tcgen05.ld (thread = row) -> bias / position table / ReLU / ReLU-mask /
//                           accumulate -> bf16 or fp32 tile in swizzled shared memory -> TMA store.
//                           While they drain accumulator i, the MMA warp already fills i+1.
// All global traffic goes through TMA: operands, the optional "aux" tile (old C for accumulate, or
// the saved activation whose sign masks a ReLU gradient; prefetched one tile ahead) and the output.
// Operand "majors" are encoded in the UMMA descriptors, so y = x w^T (K-major A, B), dx = dy w
// (B MN-major) and dw = dy^T x (A and B MN-major) run without transposes in HBM.
// Split-K (weight gradients: tiny output, K = B*L rows): partial fp32 tiles go by TMA to a
// caller-provided workspace; a second kernel reduces them in a fixed order (deterministic).
#include "common.cuh"
#include "gemm.h"
#include "tc_common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 64, STAGES = 4;
constexpr uint32_t A_TILE = BM * BK * 2, B_TILE = BN * BK * 2;
constexpr uint32_t RING = STAGES * (A_TILE + B_TILE);   // 128 KB
constexpr uint32_t OUT_BYTES = 64 * 1024;                // fp32 staging, or bf16 staging + aux tile
constexpr int NTHREADS = 192;
constexpr uint32_t SMEM_BYTES = RING + OUT_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 512 /*bias*/;
constexpr int AUX_NONE = 0, AUX_ACC = 1, AUX_RELU = 2;

struct TcArgs {
  int M, N, K;
  int c_bf16;              // output element type of the TMA store (0: fp32)
  const float* bias;
  const float* pos;
  int pos_period;
  int relu, aux;
  int a_mn, b_mn;
  uint32_t idesc;
  int splits, kb_per, kb_total;
  int reduce_add;          // split-K slices are summed into the fp32 C by TMA reduce-add
  int tiles_m, tiles_n;
};

__device__ __forceinline__ void decode(const TcArgs& a, int w, int& tm, int& tn, int& sp) {
  sp = w % a.splits;
  const int t = w / a.splits;
  tm = t / a.tiles_n;
  tn = t - tm * a.tiles_n;
}

__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmAux,
               const TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = tc::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + STAGES * A_TILE;
  const uint32_t sOut = base + RING, sAux = sOut + 32 * 1024;
  const uint32_t bars = sOut + OUT_BYTES;
  const uint32_t bar_full = bars, bar_empty = bars + 8 * STAGES;
  const uint32_t bar_accf = bars + 16 * STAGES, bar_acce = bar_accf + 16, bar_aux = bar_acce + 16;
  const uint32_t tmem_slot = bar_aux + 8;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  float* sbias = reinterpret_cast<float*>(smem_raw + (bars + 256 - raw));   // bias of this tile

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = a.tiles_m * a.tiles_n * a.splits;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(bar_full + 8 * s, 1);
      tc::mbar_init(bar_empty + 8 * s, 1);
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(bar_accf + 8 * i, 1);
      tc::mbar_init(bar_acce + 8 * i, 128);
    }
    tc::mbar_init(bar_aux, 1);
    tc::fence_barrier_init();
    tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmB);
    tc::tma_prefetch_desc(&tmC);
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, 2 * BN);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t it = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x) {
        int tm, tn, sp;
        decode(a, w, tm, tn, sp);
        const int m0 = tm * BM, n0 = tn * BN;
        const int kb_beg = sp * a.kb_per, kb_end = min(a.kb_total, kb_beg + a.kb_per);
        for (int kb = kb_beg; kb < kb_end; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          tc::mbar_wait(bar_empty + 8 * s, ph ^ 1);
          tc::mbar_expect_tx(bar_full + 8 * s, A_TILE + B_TILE);
          const int k0 = kb * BK;
          const uint32_t da = sA + s * A_TILE, db = sB + s * B_TILE;
          if (!a.a_mn) {
            tc::tma_load_2d(da, &tmA, k0, m0, bar_full + 8 * s);            // box {64 k, 128 m}
          } else {
            tc::tma_load_2d(da, &tmA, m0, k0, bar_full + 8 * s);            // box {64 m, 64 k}
            tc::tma_load_2d(da + A_TILE / 2, &tmA, m0 + 64, k0, bar_full + 8 * s);
          }
          if (!a.b_mn) {
            tc::tma_load_2d(db, &tmB, k0, n0, bar_full + 8 * s);
          } else {
            tc::tma_load_2d(db, &tmB, n0, k0, bar_full + 8 * s);
            tc::tma_load_2d(db + B_TILE / 2, &tmB, n0 + 64, k0, bar_full + 8 * s);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      uint32_t it = 0, ti = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++ti) {
        int tm, tn, sp;
        decode(a, w, tm, tn, sp);
        const int kb_beg = sp * a.kb_per, kb_end = min(a.kb_total, kb_beg + a.kb_per);
        const uint32_t buf = ti & 1, aph = (ti >> 1) & 1;
        tc::mbar_wait(bar_acce + 8 * buf, aph ^ 1);     // epilogue has drained this accumulator
        tc::tc_fence_after();
        const uint32_t tacc = tmem_base + buf * BN;
        for (int kb = kb_beg; kb < kb_end; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          tc::mbar_wait(bar_full + 8 * s, ph);
          tc::tc_fence_after();
          const uint32_t da = sA + s * A_TILE, db = sB + s * B_TILE;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major: +32 B per 16 k inside the 128 B swizzle row; MN-major: +16 rows of 128 B
            const uint64_t ad = a.a_mn ? tc::smem_desc_sw128(da + k * 2048, A_TILE / 2, 1024)
                                       : tc::smem_desc_sw128(da + k * 32, 16, 1024);
            const uint64_t bd = a.b_mn ? tc::smem_desc_sw128(db + k * 2048, B_TILE / 2, 1024)
                                       : tc::smem_desc_sw128(db + k * 32, 16, 1024);
            tc::umma_bf16(tacc, ad, bd, a.idesc, (kb > kb_beg || k > 0) ? 1u : 0u);
          }
          tc::umma_commit(bar_empty + 8 * s);   // ring slot reusable once these MMAs have read it
        }
        tc::umma_commit(bar_accf + 8 * buf);    // accumulator complete
      }
    }
  } else {
    // ================= epilogue: warps 2..5, TMEM lane quarter = warp % 4 =================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int te = threadIdx.x - 64;
    const bool partial = a.splits > 1;
    const bool out_f32 = partial || !a.c_bf16;
    auto issue_aux = [&](int w) {   // old C (accumulate) or saved activation (ReLU mask), bf16
      int tm, tn, sp;
      decode(a, w, tm, tn, sp);
      tc::mbar_expect_tx(bar_aux, 32 * 1024);
      tc::tma_load_2d(sAux, &tmAux, tn * BN, tm * BM, bar_aux);
      tc::tma_load_2d(sAux + 16384, &tmAux, tn * BN + 64, tm * BM, bar_aux);
    };
    if (a.aux && te == 0 && (int)blockIdx.x < total) issue_aux(blockIdx.x);
    uint32_t ti = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x, ++ti) {
      int tm, tn, sp;
      decode(a, w, tm, tn, sp);
      const int m0 = tm * BM, n0 = tn * BN;
      const int64_t m = (int64_t)m0 + row;
      const int pos_row = a.pos ? (int)((uint32_t)(m0 + row) % (uint32_t)a.pos_period) : 0;
      const uint32_t buf = ti & 1, aph = (ti >> 1) & 1;
      tc::mbar_wait(bar_accf + 8 * buf, aph);
      tc::tc_fence_after();
      if (te == 0) tc::tma_store_wait_read();          // previous tile's store has left staging
      if (a.bias) sbias[te] = (n0 + te < a.N) ? __ldg(a.bias + n0 + te) : 0.f;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (a.aux) tc::mbar_wait(bar_aux, ti & 1);
      const uint32_t taddr = tmem_base + buf * BN + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
      for (int ch = 0; ch < BN / 32; ++ch) {
        uint32_t r[32];
        tc::tmem_ld32(taddr + ch * 32, r);
        tc::tmem_ld_wait();
        // rows >= M / columns >= N hold zeros or junk that the TMA store clips: no guards needed
        if (a.bias) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            r[i] = __float_as_uint(__uint_as_float(r[i]) + sbias[ch * 32 + i]);
        }
        if (a.pos && m < a.M) {   // position table row of this output row (period = seq length)
          const float* prow = a.pos + (int64_t)pos_row * a.N + n0 + ch * 32;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (n0 + ch * 32 + i < a.N) r[i] = __float_as_uint(__uint_as_float(r[i]) + __ldg(prow + i));
        }
        if (a.relu) {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(fmaxf(__uint_as_float(r[i]), 0.f));
        }
        if (a.aux) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t x[4];
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3])
                         : "r"(sAux + (ch >> 1) * 16384 + tc::sw128_offset(row, (ch & 1) * 4 + c)));
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float lo = __uint_as_float(x[e] << 16), hi = __uint_as_float(x[e] & 0xFFFF0000u);
              float v0 = __uint_as_float(r[8 * c + 2 * e]), v1 = __uint_as_float(r[8 * c + 2 * e + 1]);
              if (a.aux == AUX_ACC) { v0 += lo; v1 += hi; }
              else { if (!(lo > 0.f)) v0 = 0.f; if (!(hi > 0.f)) v1 = 0.f; }
              r[8 * c + 2 * e] = __float_as_uint(v0);
              r[8 * c + 2 * e + 1] = __float_as_uint(v1);
            }
          }
        }
        if (out_f32) {      // 4 panels of 32 fp32 columns: [128 rows][128 B]
#pragma unroll
          for (int c = 0; c < 8; ++c)
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(
                             sOut + ch * 16384 + tc::sw128_offset(row, c)),
                         "r"(r[4 * c]), "r"(r[4 * c + 1]), "r"(r[4 * c + 2]), "r"(r[4 * c + 3])
                         : "memory");
        } else {            // 2 panels of 64 bf16 columns
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              __nv_bfloat162 t = __floats2bfloat162_rn(__uint_as_float(r[8 * c + 2 * e]),
                                                       __uint_as_float(r[8 * c + 2 * e + 1]));
              pk[e] = *reinterpret_cast<uint32_t*>(&t);
            }
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(
                             sOut + (ch >> 1) * 16384 + tc::sw128_offset(row, (ch & 1) * 4 + c)),
                         "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3])
                         : "memory");
          }
        }
      }
      // accumulator drained -> MMA warp may reuse it; staging complete -> TMA store
      tc::tc_fence_before();
      tc::mbar_arrive(bar_acce + 8 * buf);
      tc::fence_proxy_async();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (te == 0) {
        if (partial && a.reduce_add) {   // C += slice, summed by the L2
#pragma unroll
          for (int p = 0; p < 4; ++p)
            if (n0 + p * 32 < a.N) tc::tma_reduce_add_2d(&tmC, sOut + p * 16384, n0 + p * 32, m0);
        } else if (partial) {   // workspace viewed as [(tile*splits + split)*128 rows][128 fp32 cols]
          const int prow = ((tm * a.tiles_n + tn) * a.splits + sp) * BM;
#pragma unroll
          for (int p = 0; p < 4; ++p) tc::tma_store_2d(&tmC, sOut + p * 16384, p * 32, prow);
        } else if (out_f32) {
#pragma unroll
          for (int p = 0; p < 4; ++p)
            if (n0 + p * 32 < a.N) tc::tma_store_2d(&tmC, sOut + p * 16384, n0 + p * 32, m0);
        } else {
          tc::tma_store_2d(&tmC, sOut, n0, m0);
          if (n0 + 64 < a.N) tc::tma_store_2d(&tmC, sOut + 16384, n0 + 64, m0);
        }
        tc::tma_store_commit();
        if (a.aux && w + (int)gridDim.x < total) issue_aux(w + gridDim.x);   // prefetch next aux
      }
    }
    if (te == 0) tc::tma_store_wait_all();
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 2 * BN);
  }
}

struct ReduceArgs {
  const float* partial;
  void* C;
  int64_t ldc;
  int M, N, splits, tiles_n, c_bf16, accumulate;
  const float* bias;
};

// sum the split-K partial tiles in a fixed order, then bias / accumulate / convert / store
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const ReduceArgs a) {
  const int tile = blockIdx.x;
  const int m0 = (tile / a.tiles_n) * BM, n0 = (tile % a.tiles_n) * BN;
  const float* src = a.partial + (size_t)tile * a.splits * BM * BN;
  for (int e = threadIdx.x + blockIdx.y * 256; e < BM * BN / 4; e += 256 * gridDim.y) {
    const int row = e / (BN / 4), c4 = (e % (BN / 4)) * 4;
    const int64_t m = (int64_t)m0 + row;
    const int n = n0 + c4;
    if (m >= a.M || n >= a.N) continue;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < a.splits; ++s) {
      const float4 v = *reinterpret_cast<const float4*>(src + ((size_t)s * BM + row) * BN + c4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    float x[4] = {acc.x, acc.y, acc.z, acc.w};
    if (!a.c_bf16 && n + 4 <= a.N && !a.bias) {   // common case (weight gradients): one 16-byte store
      float4* c = reinterpret_cast<float4*>(static_cast<float*>(a.C) + m * a.ldc + n);
      float4 v = make_float4(x[0], x[1], x[2], x[3]);
      if (a.accumulate) {
        const float4 o = *c;
        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
      }
      *c = v;
      continue;
    }
    for (int j = 0; j < 4 && n + j < a.N; ++j) {
      if (a.bias) x[j] += __ldg(a.bias + n + j);
      if (a.c_bf16) {
        bf16* c = static_cast<bf16*>(a.C) + m * a.ldc + n + j;
        *c = from_f<bf16>(a.accumulate ? x[j] + to_f(*c) : x[j]);
      } else {
        float* c = static_cast<float*>(a.C) + m * a.ldc + n + j;
        *c = a.accumulate ? x[j] + *c : x[j];
      }
    }
  }
}

float* g_ws = nullptr;
size_t g_ws_bytes = 0;

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

bool make_tmap(CUtensorMap* out, CUtensorMapDataType dt, const void* base, int rank,
               const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
  PFN_encodeTiled enc = mm_get_encode_tiled();
  if (!enc) return false;
  cuuint64_t gd[5], gs[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i + 1 < rank) gs[i] = strides_bytes[i];
  }
  const CUresult r = enc(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace

extern "C" int mmemo_set_workspace(void* ptr, int64_t bytes) {
  g_ws = static_cast<float*>(ptr);
  g_ws_bytes = ptr ? (size_t)bytes : 0;
  return MMEMO_OK;
}

PFN_encodeTiled mm_get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

bool mm_make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                       const uint64_t* strides_bytes, const uint32_t* box) {
  return make_tmap(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box);
}
bool mm_make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box) {
  return make_tmap(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box);
}

bool gemm_tc_supported(const GemmArgs& g, int c_bf16) {
  if (g.M < 64 || g.N < 64 || g.K < 64) return false;
  if ((double)g.M * (double)g.N * (double)g.K < (double)(1 << 22)) return false;
  if (g.M > (1ll << 31) - 256 || g.N > (1ll << 31) - 256 || g.K > (1ll << 31) - 256) return false;
  const bool a_k = (g.sAk == 1 && g.sAm % 8 == 0), a_mn = (g.sAm == 1 && g.sAk % 8 == 0);
  const bool b_k = (g.sBk == 1 && g.sBn % 8 == 0), b_mn = (g.sBn == 1 && g.sBk % 8 == 0);
  if (!(a_k || a_mn) || !(b_k || b_mn)) return false;
  if (!aligned16(g.A) || !aligned16(g.B) || !aligned16(g.C)) return false;
  if (g.ldc % (c_bf16 ? 8 : 4) != 0) return false;
  if (g.relu_src && (!g.relu_src_bf16 || !aligned16(g.relu_src) || g.ldrelu % 8 != 0)) return false;
  if (g.relu_src && (g.accumulate || !c_bf16)) return false;   // one aux tile at a time
  if (g.accumulate && !c_bf16) return false;                   // fp32 C += ... : SIMT path
  return true;
}

int gemm_tc(const GemmArgs& g, int c_bf16, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    MM_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)SMEM_BYTES));
    attr_done = true;
  }
  const bool a_mn = !(g.sAk == 1 && g.sAm % 8 == 0);
  const bool b_mn = !(g.sBk == 1 && g.sBn % 8 == 0);
  const int tiles_n = (int)cdiv(g.N, BN), tiles_m = (int)cdiv(g.M, BM);
  const int tiles = tiles_m * tiles_n;
  const int kb_total = (int)cdiv(g.K, BK);
  const int sms = num_sms();
  // split-K when the output has too few tiles to occupy the machine and K is long (needs a linear
  // epilogue: bias and accumulate are applied by the reduce kernel)
  int splits = 1;
  const bool reduce_add = !c_bf16 && !g.bias;   // fp32 C: slices are summed by TMA reduce-add
  if (tiles * 2 <= sms && kb_total >= 16 && !g.relu && !g.relu_src && !g.pos) {
    splits = sms / tiles;                               // one balanced round of work items
    if (splits > kb_total / 4) splits = kb_total / 4;
    if (splits > 32) splits = 32;
    if (!reduce_add) {
      const size_t per_split = (size_t)tiles * BM * BN * sizeof(float);
      if (per_split * (size_t)splits > g_ws_bytes) splits = (int)(g_ws_bytes / per_split);
    }
    if (splits < 2) splits = 1;
  }
  const int kb_per = (int)cdiv(kb_total, splits);
  splits = (int)cdiv(kb_total, kb_per);               // every slice owns >= 1 k-block
  const bool partial = splits > 1;
  const int aux = partial ? AUX_NONE : (g.relu_src ? AUX_RELU : (g.accumulate ? AUX_ACC : AUX_NONE));

  CUtensorMap tmA, tmB, tmC, tmAux;
  {
    uint64_t dims[2], str[1];
    uint32_t box[2];
    if (!a_mn) { dims[0] = g.K; dims[1] = g.M; str[0] = g.sAm * 2; box[0] = 64; box[1] = BM; }
    else       { dims[0] = g.M; dims[1] = g.K; str[0] = g.sAk * 2; box[0] = 64; box[1] = BK; }
    bool ok = mm_make_tmap_bf16(&tmA, g.A, 2, dims, str, box);
    if (!b_mn) { dims[0] = g.K; dims[1] = g.N; str[0] = g.sBn * 2; box[0] = 64; box[1] = BN; }
    else       { dims[0] = g.N; dims[1] = g.K; str[0] = g.sBk * 2; box[0] = 64; box[1] = BK; }
    ok = ok && mm_make_tmap_bf16(&tmB, g.B, 2, dims, str, box);
    if (partial && !reduce_add) {
      dims[0] = BN; dims[1] = (uint64_t)tiles * splits * BM; str[0] = BN * 4; box[0] = 32; box[1] = BM;
      ok = ok && mm_make_tmap_f32(&tmC, g_ws, 2, dims, str, box);
    } else if (!c_bf16) {
      dims[0] = g.N; dims[1] = g.M; str[0] = g.ldc * 4; box[0] = 32; box[1] = BM;
      ok = ok && mm_make_tmap_f32(&tmC, g.C, 2, dims, str, box);
    } else {
      dims[0] = g.N; dims[1] = g.M; str[0] = g.ldc * 2; box[0] = 64; box[1] = BM;
      ok = ok && mm_make_tmap_bf16(&tmC, g.C, 2, dims, str, box);
    }
    if (aux == AUX_RELU) {
      dims[0] = g.N; dims[1] = g.M; str[0] = g.ldrelu * 2; box[0] = 64; box[1] = BM;
      ok = ok && mm_make_tmap_bf16(&tmAux, g.relu_src, 2, dims, str, box);
    } else if (aux == AUX_ACC) {
      dims[0] = g.N; dims[1] = g.M; str[0] = g.ldc * 2; box[0] = 64; box[1] = BM;
      ok = ok && mm_make_tmap_bf16(&tmAux, g.C, 2, dims, str, box);
    } else {
      tmAux = tmA;
    }
    if (!ok) {
      mmemo_set_error("cuTensorMapEncodeTiled failed (gemm_tc)", __FILE__, __LINE__);
      return MMEMO_ERR_CUDA;
    }
  }
  TcArgs a = {};
  a.M = (int)g.M; a.N = (int)g.N; a.K = (int)g.K; a.c_bf16 = c_bf16;
  a.bias = partial ? nullptr : g.bias;
  a.pos = g.pos; a.pos_period = (int)g.pos_period;
  a.relu = g.relu; a.aux = aux;
  a.a_mn = a_mn; a.b_mn = b_mn;
  a.idesc = tc::idesc_bf16(BM, BN, a_mn, b_mn);
  a.splits = splits; a.kb_per = kb_per; a.kb_total = kb_total;
  a.reduce_add = partial && reduce_add;
  a.tiles_m = tiles_m; a.tiles_n = tiles_n;
  const int total = tiles * splits;
  const int grid = total < sms ? total : sms;
  if (a.reduce_add && !g.accumulate)   // the slices accumulate into C: start from zero
    MM_CUDA_OK(cudaMemset2DAsync(g.C, g.ldc * sizeof(float), 0, g.N * sizeof(float), g.M, st));
  gemm_tc_kernel<<<grid, NTHREADS, SMEM_BYTES, st>>>(tmA, tmB, tmC, tmAux, a);
  MM_LAUNCH_OK();
  if (partial && !a.reduce_add) {
    ReduceArgs r = {};
    r.partial = g_ws; r.C = g.C; r.ldc = g.ldc; r.M = (int)g.M; r.N = (int)g.N;
    r.splits = splits; r.tiles_n = tiles_n; r.c_bf16 = c_bf16; r.accumulate = g.accumulate;
    r.bias = g.bias;
    int ysplit = (int)cdiv(2 * sms, tiles);
    if (ysplit > 16) ysplit = 16;
    splitk_reduce_kernel<<<dim3((unsigned)tiles, (unsigned)ysplit), 256, 0, st>>>(r);
    MM_LAUNCH_OK();
  }
  return MMEMO_OK;
}
