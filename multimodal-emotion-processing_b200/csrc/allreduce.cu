// Gradient exchange for the data-parallel path: two-shot sum all-reduce of a
// float32 range that lives in symmetric memory (the same virtual layout on every rank).
//
//   rank r owns the r-th 1/world slice of the range
//   NVLS path   : v = multimem.ld_reduce.add [mc + i]   (the NVSwitch sums the replica of every rank)
//                 multimem.st [mc + i], v               (the switch writes v into every replica)
//   peer path   : v = sum_p ld [buf_p + i];  st [buf_p + i], v  for every p (plain NVLink P2P)
//
// Blocks of the same index on all ranks meet at a flag barrier in the ranks' signal pads before
// (every rank's gradients are written) and after (every replica is complete).  The kernel needs
// no shared memory and its CTAs are 256 threads x <= 96 registers (24 576 registers), which is what
// a persistent GEMM CTA (320 threads x 128) or a tcgen05 attention CTA leaves free on an SM: the
// all-reduce co-resides with them instead of queueing behind one-CTA-per-SM grids whose work is
// partitioned statically (a 512-thread version could not, and the GEMMs it delayed exposed nearly
// the whole transfer).  The LayerNorm backward fills the register file, so it and the GEMMs shrink
// their grids by the stream's SM budget while a bucket is in flight (dp.GradReducer).
#include "common.cuh"

namespace {

constexpr unsigned long long kSpinLimitNs = 20000000000ull;  // 20 s, then trap instead of hanging

__device__ __forceinline__ unsigned long long now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// flag := 1 once the peer has consumed the previous one (release: our earlier writes are visible)
__device__ __forceinline__ void signal_put(uint32_t* flag) {
    const unsigned long long t0 = now_ns();
    uint32_t old;
    do {
        asm volatile("atom.global.release.sys.cas.b32 %0, [%1], 0, 1;"
                     : "=r"(old) : "l"(flag) : "memory");
        if (old != 0 && now_ns() - t0 > kSpinLimitNs) __trap();
    } while (old != 0);
}

// wait for flag == 1 and reset it to 0 (acquire)
__device__ __forceinline__ void signal_wait(uint32_t* flag) {
    const unsigned long long t0 = now_ns();
    uint32_t old;
    do {
        asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], 1, 0;"
                     : "=r"(old) : "l"(flag) : "memory");
        if (old != 1 && now_ns() - t0 > kSpinLimitNs) __trap();
    } while (old != 1);
}

__device__ __forceinline__ void rank_barrier(uint32_t* const* pads, int slot_base, int rank,
                                             int world) {
    __syncthreads();
    if ((int)threadIdx.x < world) {
        const int peer = threadIdx.x;
        const int slot = slot_base + blockIdx.x * world;
        signal_put(pads[peer] + slot + rank);
        signal_wait(pads[rank] + slot + peer);
    }
    __syncthreads();
}

__device__ __forceinline__ float4 mc_ld_reduce(const float* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st(float* mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

constexpr int kMaxWorld = 16;
constexpr int kUnroll = 12;
constexpr int kThreads = 256;

template <bool NVLS>
__global__ void __maxnreg__(96)
allreduce_kernel(float* mc, float* const* bufs, uint32_t* const* pads, int slot_base,
                 int64_t offset, int64_t n, int rank, int world) {
    rank_barrier(pads, slot_base, rank, world);

    const int64_t per = n / 4 / world;                       // float4 per rank slice
    const int64_t first = offset / 4 + (int64_t)rank * per;
    // kUnroll independent 16-byte requests per thread in flight: one NVLink round trip is ~2 us, so
    // bandwidth comes from the number of outstanding loads, not from the number of CTAs
    const int64_t chunk = (int64_t)blockDim.x * kUnroll;
    for (int64_t c = (int64_t)blockIdx.x * chunk; c < per; c += (int64_t)gridDim.x * chunk) {
        float4 v[kUnroll];
        if (NVLS) {
            float4* base = reinterpret_cast<float4*>(mc) + first;
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int64_t i = c + u * blockDim.x + threadIdx.x;
                if (i < per) v[u] = mc_ld_reduce(reinterpret_cast<const float*>(base + i));
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int64_t i = c + u * blockDim.x + threadIdx.x;
                if (i < per) mc_st(reinterpret_cast<float*>(base + i), v[u]);
            }
        } else {
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            // fixed rank order so that every replica receives bit-identical sums
            for (int p = 0; p < world; ++p) {
                const float4* src = reinterpret_cast<const float4*>(bufs[p]) + first;
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    const int64_t i = c + u * blockDim.x + threadIdx.x;
                    if (i < per) {
                        const float4 t = __ldcv(src + i);
                        v[u].x += t.x; v[u].y += t.y; v[u].z += t.z; v[u].w += t.w;
                    }
                }
            }
            for (int p = 0; p < world; ++p) {
                float4* dst = reinterpret_cast<float4*>(bufs[p]) + first;
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    const int64_t i = c + u * blockDim.x + threadIdx.x;
                    if (i < per) dst[i] = v[u];
                }
            }
        }
    }

    rank_barrier(pads, slot_base, rank, world);
}

}  // namespace

extern "C" int mmemo_allreduce_sum_f32(void* multicast_ptr, void* const* buffer_ptrs_dev,
                                       void* const* signal_pad_ptrs_dev, int64_t signal_slot_base,
                                       int64_t offset_elems, int64_t n_elems, int rank, int world,
                                       int blocks, mmemo_stream_t stream) {
    if (!buffer_ptrs_dev || !signal_pad_ptrs_dev || world < 1 || world > kMaxWorld || rank < 0 ||
        rank >= world || blocks < 1 || n_elems < 0 || offset_elems < 0)
        return MMEMO_ERR_ARG;
    if (n_elems % (4 * world) != 0 || offset_elems % 4 != 0) return MMEMO_ERR_SHAPE;
    if (n_elems == 0) return MMEMO_OK;
    cudaStream_t s = mm_stream(stream);
    auto bufs = reinterpret_cast<float* const*>(buffer_ptrs_dev);
    auto pads = reinterpret_cast<uint32_t* const*>(signal_pad_ptrs_dev);
    if (multicast_ptr)
        allreduce_kernel<true><<<blocks, kThreads, 0, s>>>(static_cast<float*>(multicast_ptr), bufs,
                                                      pads, (int)signal_slot_base, offset_elems,
                                                      n_elems, rank, world);
    else
        allreduce_kernel<false><<<blocks, kThreads, 0, s>>>(nullptr, bufs, pads, (int)signal_slot_base,
                                                       offset_elems, n_elems, rank, world);
    MM_LAUNCH_OK();
    return MMEMO_OK;
}
