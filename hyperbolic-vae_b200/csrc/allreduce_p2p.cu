// Gradient all-reduce over NVLink peer memory: the flat gradient bucket lives in symmetric memory (every
// rank can address every rank's copy), so the sum is one kernel per rank instead of a library collective:
//   barrier  ->  rank r sums slice r of all W copies (peer loads, fixed rank order: bit-identical on every rank)
//            ->  writes the result into all W copies (peer stores)  ->  barrier.
// In place and race-free: rank r reads only slice r and writes only slice r, everywhere.  The payload is 2-4 MB, so the
// cost is two NVLink round trips plus the barriers (~15 us) where NCCL's launch + protocol took 40-70 us in the step.
// The reference has no distributed code (devices=1, training/trainer_mnist.py:19); this is the data-parallel exchange
// the build adds.  Barriers are per block (block b of every rank meets block b of every other rank) on binary
// semaphores in the symmetric signal pad: no epoch counter, so a captured CUDA graph can replay the kernel.
#include "hvae_common.cuh"

namespace hvae {

constexpr int kArBlocks = 128;  // upper bound (pad slots are sized for it); the caller passes the block count, agreed across ranks
constexpr long long kArSpinCycles = 8000000000LL;  // ~4 s at 2 GHz: a peer that never arrives is an error, not a hang
constexpr int kArErrSlot = 0;   // pad word 0 of this rank's own pad: set to 1 when a barrier timed out
constexpr int kArUnroll = 4;    // positions per thread in flight: the kernel is bound by NVLink round-trip latency
constexpr int kArThreads = 512;
constexpr int kArMaxWorld = 16;

__device__ __forceinline__ uint32_t cas_sys(uint32_t* addr, uint32_t cmp, uint32_t val) {
    uint32_t old;
    asm volatile("atom.cas.acq_rel.sys.global.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
    return old;
}

// ---- epoch barrier (one-way signals) -------------------------------------------------------------------------------
// The CAS barrier below costs a remote ATOMIC round trip per (block, peer) and measured 13-17 us per barrier (a 16 KB
// exchange took 26-37 us where NCCL takes 19).  Here ONE block talks to the peers with plain release STORES of a launch
// epoch (posted writes: one NVLink one-way trip), the other blocks are released through a flag in local memory:
//   pad words (per site):  [0, W) arrival flags of barrier 1 (peer p writes word p), [W, 2W) of barrier 2,
//                          2W: epoch of the last completed launch, 2W+1: blocks done with the data phase, 2W+2: local go flag
// The epoch lives in device memory and advances by one per launch, so a captured CUDA graph replays the kernel.
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// block 0, threads p < world: tell every peer "rank has reached barrier `which` of launch `epoch`" and wait for theirs
__device__ __forceinline__ bool peers_meet(uint32_t* const* pads, int rank, int world, int base, int which, uint32_t epoch) {
    bool ok = true;
    if ((int)threadIdx.x < world) {
        const int p = threadIdx.x;
        st_release_sys(pads[p] + base + which * world + rank, epoch);
        const uint32_t* mine = pads[rank] + base + which * world + p;
        const long long t0 = clock64();
        while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
            if (clock64() - t0 > kArSpinCycles) { ok = false; pads[rank][kArErrSlot] = 1u; break; }
        }
    }
    return ok;
}
// barrier 1 (all blocks): every rank's inputs are complete.  Returns the launch epoch.
__device__ __forceinline__ uint32_t grid_enter(uint32_t* const* pads, int rank, int world, int base) {
    uint32_t* my = pads[rank] + base;
    const uint32_t epoch = ld_acquire_gpu(my + 2 * world) + 1u;   // (block 0 publishes it only after every block has arrived)
    if (blockIdx.x == 0) {
        peers_meet(pads, rank, world, base, 0, epoch);
        __syncthreads();
        if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(my + 2 * world + 2), "r"(epoch) : "memory");
    } else {
        if (threadIdx.x == 0) {
            const long long t0 = clock64();
            while ((int32_t)(ld_acquire_gpu(my + 2 * world + 2) - epoch) < 0) {
                if (clock64() - t0 > kArSpinCycles) { pads[rank][kArErrSlot] = 1u; break; }
            }
        }
        __syncthreads();
    }
    return epoch;
}
// barrier 2: this rank's stores are visible everywhere AND every peer's stores into this copy have landed
__device__ __forceinline__ void grid_leave(uint32_t* const* pads, int rank, int world, int base, uint32_t epoch) {
    uint32_t* my = pads[rank] + base;
    __threadfence_system();
    __syncthreads();
    if (blockIdx.x != 0) {
        if (threadIdx.x == 0) atomicAdd(my + 2 * world + 1, 1u);
        return;
    }
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        while (ld_acquire_gpu(my + 2 * world + 1) < gridDim.x - 1) {
            if (clock64() - t0 > kArSpinCycles) { pads[rank][kArErrSlot] = 1u; break; }
        }
        my[2 * world + 1] = 0u;
        __threadfence_system();
    }
    __syncthreads();
    peers_meet(pads, rank, world, base, 1, epoch);
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(my + 2 * world), "r"(epoch) : "memory");
}

__device__ __forceinline__ float4 ld_peer(const float* p) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

__global__ void __launch_bounds__(kArThreads)
k_allreduce_p2p(float* const* __restrict__ bufs, uint32_t* const* __restrict__ pads, int rank, int world, int64_t off,
                int64_t n, int slot_base, float scale) {
    const uint32_t epoch = grid_enter(pads, rank, world, slot_base);  // every rank's gradients are complete (its earlier kernels have retired)
    const int64_t n4 = n >> 2;
    const int64_t per = (n4 + world - 1) / world;
    const int64_t lo = (int64_t)rank * per, hi = (lo + per < n4) ? lo + per : n4;
    float* base[kArMaxWorld];
#pragma unroll
    for (int p = 0; p < kArMaxWorld; ++p) base[p] = (p < world) ? bufs[p] + off : nullptr;
    const int64_t stride = (int64_t)gridDim.x * kArThreads;
    for (int64_t i0 = lo + (int64_t)blockIdx.x * kArThreads + threadIdx.x; i0 < hi; i0 += stride * kArUnroll) {
        float4 acc[kArUnroll];
#pragma unroll
        for (int u = 0; u < kArUnroll; ++u) {
            const int64_t i = i0 + u * stride;
            acc[u] = (i < hi) ? ld_peer(base[0] + 4 * i) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
#pragma unroll
        for (int p = 1; p < kArMaxWorld; ++p) {
            if (p < world) {
                float4 v[kArUnroll];
#pragma unroll
                for (int u = 0; u < kArUnroll; ++u) {
                    const int64_t i = i0 + u * stride;
                    v[u] = (i < hi) ? ld_peer(base[p] + 4 * i) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                }
#pragma unroll
                for (int u = 0; u < kArUnroll; ++u) {
                    acc[u].x += v[u].x; acc[u].y += v[u].y; acc[u].z += v[u].z; acc[u].w += v[u].w;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kArUnroll; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < hi) {
                float4 a = acc[u];
                a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;
#pragma unroll
                for (int p = 0; p < kArMaxWorld; ++p)
                    if (p < world) *reinterpret_cast<float4*>(base[p] + 4 * i) = a;
            }
        }
    }
    grid_leave(pads, rank, world, slot_base, epoch);  // every rank's stores into this copy have landed
}

// ---- NVLS variant: the NVSwitch does the sum -----------------------------------------------------------------------
// With the bucket also mapped at a MULTICAST address (torch symmetric memory's multicast_ptr), one
// multimem.ld_reduce returns the sum of all W copies of a 16-byte word (reduced inside the switch) and one multimem.st
// writes a word into all W copies: rank r handles slice r with 1/W of the loads of the peer-loop kernel above and ONE
// NVLink round trip instead of W dependent ones.  Same barriers, same in-place/race-free slicing.
__device__ __forceinline__ float4 mm_ld_reduce(const float* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(mc)
                 : "memory");
    return v;
}
__device__ __forceinline__ void mm_st(float* mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

__global__ void __launch_bounds__(kArThreads)
k_allreduce_nvls(float* __restrict__ mc, uint32_t* const* __restrict__ pads, int rank, int world, int64_t off, int64_t n,
                 int slot_base, float scale) {
    const uint32_t epoch = grid_enter(pads, rank, world, slot_base);  // every rank's gradients are complete
    const int64_t n4 = n >> 2;
    const int64_t per = (n4 + world - 1) / world;
    const int64_t lo = (int64_t)rank * per, hi = (lo + per < n4) ? lo + per : n4;
    float* base = mc + off;
    // latency-bound (a few MB): every thread puts kNvlsUnroll in-switch reductions in flight before it touches a result, and
    // the grid is sized so that the slice is ONE such pass where it can be (128 blocks x 512 threads x 8 words = 8 MB)
    constexpr int kNvlsUnroll = 8;
    const int64_t stride = (int64_t)gridDim.x * kArThreads;
    for (int64_t i0 = lo + (int64_t)blockIdx.x * kArThreads + threadIdx.x; i0 < hi; i0 += stride * kNvlsUnroll) {
        float4 acc[kNvlsUnroll];
#pragma unroll
        for (int u = 0; u < kNvlsUnroll; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < hi) acc[u] = mm_ld_reduce(base + 4 * i);
        }
#pragma unroll
        for (int u = 0; u < kNvlsUnroll; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < hi) {
                float4 a = acc[u];
                a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;
                mm_st(base + 4 * i, a);
            }
        }
    }
    grid_leave(pads, rank, world, slot_base, epoch);  // every rank's multicast stores have landed in this copy
}

}  // namespace hvae

using namespace hvae;

extern "C" int hvae_allreduce_p2p_slots(int world) { return world > 0 ? kArBlocks * world : 0; }

// buf_ptrs_dev / pad_ptrs_dev: DEVICE arrays of `world` pointers (this rank's view of every rank's bucket and signal
// pad, e.g. torch symmetric memory's buffer_ptrs_dev / signal_pad_ptrs_dev).  Reduces elements [offset, offset + n) of
// the buckets in place (n and offset multiples of 4), result scaled by `scale`.  pad_slot_base: first uint32 slot of the
// pad this call may use (hvae_allreduce_p2p_slots(world) slots, zero-initialised; concurrent calls need disjoint ranges).
// blocks: grid size, 1..128, MUST be the same on every rank (block b meets block b); 0 picks the default (64).
extern "C" int hvae_allreduce_p2p_f32(const void* buf_ptrs_dev, const void* pad_ptrs_dev, int rank, int world, int64_t offset,
                                      int64_t n, int pad_slot_base, float scale, int blocks, void* stream) {
    if (world < 1 || world > kArMaxWorld || rank < 0 || rank >= world || n <= 0 || (n & 3) || (offset & 3) || offset < 0)
        return HVAE_ESHAPE;
    if (!buf_ptrs_dev || !pad_ptrs_dev || pad_slot_base < 0) return HVAE_EARG;
    if (blocks == 0) blocks = 64;
    if (blocks < 1 || blocks > kArBlocks) return HVAE_EARG;
    if (pad_slot_base < 1) return HVAE_EARG;  // word 0 is the error word
    k_allreduce_p2p<<<blocks, kArThreads, 0, (cudaStream_t)stream>>>((float* const*)buf_ptrs_dev, (uint32_t* const*)pad_ptrs_dev,
                                                                        rank, world, offset, n, pad_slot_base, scale);
    return check_launch();
}

// NVLS flavour: mc_ptr = the bucket's MULTICAST address on this rank (one pointer, not a table); the switch reduces.
// Everything else as hvae_allreduce_p2p_f32.  Requires NVSwitch multicast support (the caller checks that the
// symmetric-memory handle has a multicast pointer).
extern "C" int hvae_allreduce_nvls_f32(void* mc_ptr, const void* pad_ptrs_dev, int rank, int world, int64_t offset, int64_t n,
                                       int pad_slot_base, float scale, int blocks, void* stream) {
    if (world < 1 || world > kArMaxWorld || rank < 0 || rank >= world || n <= 0 || (n & 3) || (offset & 3) || offset < 0)
        return HVAE_ESHAPE;
    if (!mc_ptr || !pad_ptrs_dev) return HVAE_EARG;
    if (blocks == 0) {   // one pass of 8 words per thread over this rank's slice where 128 blocks allow it
        const int64_t per4 = ((n >> 2) + world - 1) / world;
        const int64_t want = (per4 + (int64_t)kArThreads * 8 - 1) / ((int64_t)kArThreads * 8);
        blocks = (int)(want < 4 ? 4 : (want > kArBlocks ? kArBlocks : want));
    }
    if (blocks < 1 || blocks > kArBlocks || pad_slot_base < 1) return HVAE_EARG;
    k_allreduce_nvls<<<blocks, kArThreads, 0, (cudaStream_t)stream>>>((float*)mc_ptr, (uint32_t* const*)pad_ptrs_dev, rank, world,
                                                                       offset, n, pad_slot_base, scale);
    return check_launch();
}
