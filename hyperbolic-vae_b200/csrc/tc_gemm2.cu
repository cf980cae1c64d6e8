// K1-TC / K2-TC on CTA pairs: the dense contractions of the Mobius and gyroplane layers as cta_group::2 tcgen05 GEMMs
// (bf16 operands, fp32 TMEM accumulators) with the hyperbolic math fused into the epilogue.
//
// reference: hyperbolic_vae/layers.py:145-147 (MobiusLayer -> geoopt mobius_matvec's tensordot),
//            layers.py:193-210 (Distance2PoincareHyperplanes, a == p) and layers.py:96-121 -> manifolds.py:41-65
//            (GeodesicLayer / normdist2plane, a != p: TWO inner products per (row, plane), <x,p> and <x,a>).
//
// Why a second kernel: with cta_group::1 a 128x128x16 MMA reads 8 KB of operands from shared memory every 64 cycles,
// the SM's whole 128 B/clk, while TMA writes the next stage into the same memory (tc_gemm.cu measured 0.35 of the bf16
// peak on config 5, nothing saturated).  Here two CTAs of a cluster (the two SMs of a TPC) own one 256 x 256 tile:
// each CTA stages its own 128 rows of A and HALF of the B tile (128 of the 256 columns), one thread of the leader CTA
// issues tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16: 128 cycles), the hardware reads each half from its own SM.
// Operand bytes per flop from shared memory and from L2 are both half of the 128x128 kernel's.
//
// Structure (cluster of 2 CTAs, one CTA per SM, persistent over tiles, 576 threads per CTA):
//   warp 0      TMA producer (both CTAs): cp.async.bulk.tensor.2d.cta_group::2 -> own shared memory, 6-stage ring;
//               the transaction bytes of BOTH CTAs complete on the LEADER's full barrier
//   warp 1      MMA issuer (leader CTA only); tcgen05.commit.multicast frees the stage / publishes the accumulator in
//               both CTAs
//   warps 2-17  epilogue (both CTAs): each CTA drains its own 128 TMEM lanes x 256 columns (2 accumulator stages =
//               all 512 TMEM columns: the epilogue of tile i overlaps the MMAs of tile i+1); 16x256b TMEM loads,
//               fused math, sector-complete streaming stores; every epilogue warp of both CTAs releases the stage on
//               the leader's barrier
// GEO epilogue (a != p): the B tile is [128 rows of p | 128 rows of a] - the leader CTA stages the planes' p rows, the
// peer their a rows - so accumulator columns j and 128 + j hold <x,p_j> and <x,a_j> of the same plane: one GEMM, N-concatenated.
#include <cuda.h>
#include <cuda_bf16.h>

#include "tc_common.cuh"
#include "tc_gemm2.cuh"

namespace hvae {
namespace tc2 {

using namespace hvae::tc;

constexpr int BMC = 128;                     // rows per CTA
constexpr int BNH = kTileN / 2;              // B rows staged per CTA
constexpr int BK = kTcBK, UMMA_K = 16;
constexpr int STAGES = 4, ACC_STAGES = 2;
constexpr int EPI_WARPS = 4 * kCG;
constexpr int THREADS = 64 + 32 * EPI_WARPS;
constexpr uint32_t TILE_A_BYTES = BMC * BK * 2, TILE_B_BYTES = BNH * BK * 2;
constexpr uint32_t STAGE_BYTES = TILE_A_BYTES + TILE_B_BYTES;
constexpr uint32_t COLC_BYTES = EPI_WARPS * (kTileN / kCG) * 16;  // per-warp slices of per-column float4 constants
constexpr uint32_t STG_BYTES = 32 * 128;                   // per-warp staging box of the TMA store: 32 rows x 32 fp32
constexpr uint32_t BAR_BYTES = 512;
constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + EPI_WARPS * STG_BYTES + 1024 /*align slack*/ + BAR_BYTES + COLC_BYTES;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory");
// A-resident schedule (K <= 512): a CTA keeps its 128 x K panel of A for ALL n-tiles of the m-block and streams only its
// half of the B tiles.  The streaming schedule moves 512 KB from L2 per 256x256x512 tile - 12 TB/s at the measured mainloop
// rate, the L2's whole read bandwidth, so the output stores (which also pass through L2) ADD to the kernel time
// (measured: 0.72 ms without stores, 1.24 ms with; 8.6 + 4.3 GB at 10.4 TB/s).  Keeping A halves the operand traffic.
// Shared memory: 8 x 16 KB panel + 3 x 16 KB B ring + 16 x 2 KB staging boxes (32 rows x 16 fp32, 64B swizzle).
constexpr int ARES_KB = 8;                                 // k-blocks of the resident panel: K <= 512
constexpr int ARES_NSB = 3;                                // B ring
constexpr uint32_t STG_BYTES_ARES = 32 * 64;
constexpr uint32_t SMEM_BYTES_ARES = ARES_KB * TILE_A_BYTES + ARES_NSB * TILE_B_BYTES + EPI_WARPS * STG_BYTES_ARES + 1024 + BAR_BYTES + COLC_BYTES;
static_assert(SMEM_BYTES_ARES <= 232448, "exceeds the 227 KB per-CTA shared memory");
constexpr uint32_t TMEM_COLS = ACC_STAGES * kTileN;  // 512: the whole tensor memory
// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, M = 256 (the pair), N = 256
constexpr uint32_t kIdesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTileN >> 3) << 17) | ((uint32_t)(kPairM >> 4) << 24);

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// the shared::cluster address of `addr` (an address in this CTA's shared memory) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
    // default semantics (release at CTA scope): orders this thread's tcgen05.ld (after tcgen05.fence::before_thread_sync)
    // without a cluster-scope fence - .release.cluster made every epilogue warp wait for its global stores to drain
    // (MEMBAR, 18 % of the stall samples) before the MMA warp got the accumulator stage back
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// 2-CTA TMA load: data lands in THIS CTA's shared memory, the bytes are counted on the barrier `cluster_bar` (the leader's)
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t cluster_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(cluster_bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma2(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives (once the MMAs issued so far have retired) on the barrier at this shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit2(uint32_t bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask)
                 : "memory");
}
// asinh for the bf16-operand epilogues: sign(y) log2(|y| + sqrt(y^2 + 1)) * ln2 with two MUFU ops and no small-|y| branch.
// Absolute error ~1e-7 (relative 1e-7 / |y| for tiny y): far inside the 1e-2 budget of the bf16 mode, whose inputs
// <x,p> already carry 4e-3; the fp32 kernels keep asinh_fast (polynomial below 0.25).  Returns asinh(y) / ln 2.
__device__ __forceinline__ float asinh_lg2(float y) {
    float s, l;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(fmaf(y, y, 1.0f)));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(fabsf(y) + s));            // argument >= 1: never denormal, l >= 0
    return __uint_as_float(__float_as_uint(l) | (__float_as_uint(y) & 0x80000000u));   // copysign(l, y)
}

// 32 lanes x 16 consecutive columns (thread = TMEM lane)
__device__ __forceinline__ void tmem_ld16x(uint32_t addr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// accumulator columns of n-tile nt of an N-column problem: 256, or the ragged remainder rounded up to the MMA's N step (16)
__device__ __forceinline__ int tile_cols(int64_t N, int64_t nt) {
    const int64_t rem = N - nt * kTileN;
    return rem >= kTileN ? kTileN : (int)((rem + 15) & ~(int64_t)15);
}

template <int EPI, bool ARES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
k_tc_gemm2(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
           const __grid_constant__ CUtensorMap map_b2, const __grid_constant__ CUtensorMap map_d,
           const __grid_constant__ CUtensorMap map_g, Params2 prm) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;  // 128B swizzle wants 1024-byte aligned tiles (same offset in both CTAs)
    constexpr int NSTG = ARES ? ARES_NSB : STAGES;                        // ring depth (ARES: the ring holds B tiles only)
    constexpr uint32_t RING_BYTES = ARES ? ARES_KB * TILE_A_BYTES + ARES_NSB * TILE_B_BYTES : STAGES * STAGE_BYTES;
    constexpr uint32_t SBYTES = ARES ? STG_BYTES_ARES : STG_BYTES;        // staging box per epilogue warp
    constexpr int SB_COLS = ARES ? 16 : 32;                               // its width in fp32 columns
    const uint32_t stg_base = base + RING_BYTES;                          // 1024-byte aligned (swizzle atom)
    const uint32_t bars = stg_base + EPI_WARPS * SBYTES;
    const uint32_t colc_base = bars + BAR_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (NSTG + s); };
    auto tfull_bar = [&](int s) { return bars + 8u * (2 * NSTG + s); };
    auto tempty_bar = [&](int s) { return bars + 8u * (2 * NSTG + ACC_STAGES + s); };
    auto afull_bar = [&](int kb) { return bars + 8u * (2 * NSTG + 2 * ACC_STAGES + kb); };             // ARES: panel slice kb landed
    auto aempty_bar = [&](int kb) { return bars + 8u * (2 * NSTG + 2 * ACC_STAGES + ARES_KB + kb); };  // ARES: ... may be overwritten
    auto gload_bar = [&](int w) { return bars + 8u * (2 * NSTG + 2 * ACC_STAGES + 2 * ARES_KB + w); };     // GYRO_BWD: per-warp g box landed
    const uint32_t tmem_slot = bars + 8u * (2 * NSTG + 2 * ACC_STAGES + 2 * ARES_KB + EPI_WARPS);
    static_assert(8 * (2 * NSTG + 2 * ACC_STAGES + 2 * ARES_KB + EPI_WARPS) + 4 <= BAR_BYTES, "barrier block");
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));
    // streaming: [stage][A | B];  A-resident: [A panel: kb][B ring: stage]
    auto a_tile = [&](int stage_or_kb) { return ARES ? base + (uint32_t)stage_or_kb * TILE_A_BYTES : base + (uint32_t)stage_or_kb * STAGE_BYTES; };
    auto b_tile = [&](int stage) {
        return ARES ? base + ARES_KB * TILE_A_BYTES + (uint32_t)stage * TILE_B_BYTES : base + (uint32_t)stage * STAGE_BYTES + TILE_A_BYTES;
    };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();       // 0 = leader (issues the MMAs), 1 = peer
    const int64_t pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    constexpr int TN = (EPI == EPI_GEO) ? kTileN / 2 : kTileN;   // output columns per tile
    const int64_t m_tiles = (prm.M + kPairM - 1) / kPairM, n_tiles = (prm.N + TN - 1) / TN;
    const int k_blocks = (int)((prm.K + BK - 1) / BK);
    const int S = (!ARES && prm.splits > 1) ? prm.splits : 1;
    // work unit: one output tile (x split) when streaming, one m-block with all its n-tiles when A-resident
    const int64_t units = ARES ? m_tiles : m_tiles * n_tiles * S;
    // MOBIUS_F (A-resident only): column tiles of the Gram operand ahead of every m-block's n-tiles
    const int g_tiles = (EPI == EPI_MOBIUS_F) ? (int)((prm.K + kTileN - 1) / kTileN) : 0;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        if (EPI == EPI_GEO) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b2) : "memory");
        if (EPI != EPI_ROWDOT) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_d) : "memory");
        if (EPI == EPI_GYRO_BWD) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_g) : "memory");
        for (int s = 0; s < NSTG; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < ACC_STAGES; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 2 * EPI_WARPS); }
        for (int kb = 0; kb < ARES_KB; ++kb) { mbar_init(afull_bar(kb), 1); mbar_init(aempty_bar(kb), 1); }
        for (int w = 0; w < EPI_WARPS; ++w) mbar_init(gload_bar(w), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // the same warp of both CTAs allocates (and later frees) the pair's tensor memory
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    cluster_sync_all();   // barriers of both CTAs are initialised before any remote arrive / multicast commit
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ===== TMA producer (both CTAs): own 128 rows of A, own half of the B tile =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, pphase = 0;
            const uint32_t lead_full = mapa(full_bar(0), 0), lead_afull = mapa(afull_bar(0), 0);
            const CUtensorMap* mb = (EPI == EPI_GEO && rank) ? &map_b2 : &map_b;
            for (int64_t u = pair; u < units; u += npairs) {
                const int64_t tile = ARES ? u : u / S;
                const int sp = ARES ? 0 : (int)(u % S);
                const int kb0 = (int)((int64_t)sp * k_blocks / S), kb1 = (int)((int64_t)(sp + 1) * k_blocks / S);
                const int64_t mt = ARES ? u : tile / n_tiles;
                const int64_t nt0 = ARES ? 0 : tile % n_tiles, nt1 = ARES ? n_tiles : nt0 + 1;
                const int m0 = (int)(mt * kPairM + rank * BMC);
                for (int64_t nt = nt0 - g_tiles; nt < nt1; ++nt) {
                    // a ragged last n-tile is issued as a narrower MMA (N rounded up to 16): each CTA stages N/2 of its columns
                    // (MOBIUS_F: tiles nt < 0 are the column tiles of the (K, K) Gram operand)
                    const bool gram = nt < 0;
                    const int ncols = (EPI == EPI_GEO) ? kTileN : (gram ? tile_cols(prm.K, nt + g_tiles) : tile_cols(prm.N, nt));
                    const int n0 = (EPI == EPI_GEO) ? (int)(nt * TN) : (int)((gram ? nt + g_tiles : nt) * kTileN + rank * (ncols >> 1));
                    if (EPI == EPI_MOBIUS_F) mb = gram ? &map_b2 : &map_b;
                    for (int kb = kb0; kb < kb1; ++kb) {
                        if (ARES && nt == nt0 - g_tiles) {
                            // slice kb of this m-block's panel, as soon as the previous m-block's last n-tile is done with it
                            mbar_wait(aempty_bar(kb), pphase ^ 1u);
                            if (rank == 0) mbar_expect_tx(afull_bar(kb), 2u * TILE_A_BYTES);
                            tma_load_2d_2sm(a_tile(kb), &map_a, lead_afull + 8u * kb, kb * BK, m0);
                        }
                        mbar_wait(empty_bar(stage), phase ^ 1u);   // this CTA's copy of the stage has been consumed
                        if (rank == 0) mbar_expect_tx(full_bar(stage), ARES ? 2u * TILE_B_BYTES : 2u * STAGE_BYTES);
                        // K-major: one {64 k, 128 rows} box; MN-major: two {64 MN, 64 k-rows} boxes (8 KB each)
                        if (!ARES) {
                            if (prm.a_mn) {
                                tma_load_2d_2sm(a_tile(stage), &map_a, lead_full + 8u * stage, m0, kb * BK);
                                tma_load_2d_2sm(a_tile(stage) + 8192u, &map_a, lead_full + 8u * stage, m0 + 64, kb * BK);
                            } else {
                                tma_load_2d_2sm(a_tile(stage), &map_a, lead_full + 8u * stage, kb * BK, m0);
                            }
                        }
                        if (!ARES && prm.b_mn) {
                            tma_load_2d_2sm(b_tile(stage), mb, lead_full + 8u * stage, n0, kb * BK);
                            tma_load_2d_2sm(b_tile(stage) + 8192u, mb, lead_full + 8u * stage, n0 + 64, kb * BK);
                        } else {
                            tma_load_2d_2sm(b_tile(stage), mb, lead_full + 8u * stage, kb * BK, n0);
                        }
                        if (++stage == NSTG) { stage = 0; phase ^= 1u; }
                    }
                }
                pphase ^= 1u;
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only) =====
        if (rank == 0 && lane == 0) {
            int stage = 0, as = 0;
            uint32_t phase = 0, aphase = 0, pphase = 0;
            for (int64_t u = pair; u < units; u += npairs) {
                const int64_t tile = ARES ? u : u / S;
                const int sp = ARES ? 0 : (int)(u % S);
                const int kb0 = (int)((int64_t)sp * k_blocks / S), kb1 = (int)((int64_t)(sp + 1) * k_blocks / S);
                const int64_t nt0 = ARES ? 0 : tile % n_tiles, nt1 = ARES ? n_tiles : nt0 + 1;
                for (int64_t nt = nt0 - g_tiles; nt < nt1; ++nt) {
                    mbar_wait(tempty_bar(as), aphase ^ 1u);   // the epilogues of BOTH CTAs have drained this accumulator
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + (uint32_t)(as * kTileN);
                    const int ncols = (EPI == EPI_GEO) ? kTileN : (nt < 0 ? tile_cols(prm.K, nt + g_tiles) : tile_cols(prm.N, nt));
                    const uint32_t idesc_n = (kIdesc2 & ~(0x3Fu << 17)) | ((uint32_t)(ncols >> 3) << 17);
                    for (int kb = kb0; kb < kb1; ++kb) {
                        if (ARES && nt == nt0 - g_tiles) {
                            mbar_wait(afull_bar(kb), pphase);     // both CTAs' slices kb of the panel have landed
                            tc_fence_after();
                        }
                        mbar_wait(full_bar(stage), phase);    // both CTAs' tiles of this stage have landed
                        tc_fence_after();
                        const bool amn = !ARES && prm.a_mn, bmn = !ARES && prm.b_mn;
                        const uint64_t da = amn ? make_desc_mn(a_tile(stage)) : make_desc(a_tile(ARES ? kb : stage));
                        const uint64_t db = bmn ? make_desc_mn(b_tile(stage)) : make_desc(b_tile(stage));
                        // k-step in the (addr >> 4) field: K-major 16 bf16 = 32 B inside the swizzle atom (+2);
                        // MN-major 16 contraction rows = 2048 B (+128)
                        const uint64_t sa = amn ? 128u : 2u, sb = bmn ? 128u : 2u;
                        const uint32_t idesc = idesc_n | (amn ? (1u << 15) : 0u) | (bmn ? (1u << 16) : 0u);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k)
                            umma2(tmem_d, da + sa * (uint64_t)k, db + sb * (uint64_t)k, idesc, (uint32_t)(((kb - kb0) | k) != 0));
                        umma_commit2(empty_bar(stage));       // frees the stage in both CTAs when these MMAs retire
                        if (ARES && nt == nt1 - 1) umma_commit2(aempty_bar(kb));   // last use of the panel slice
                        if (++stage == NSTG) { stage = 0; phase ^= 1u; }
                    }
                    umma_commit2(tfull_bar(as));              // accumulator complete: wakes both epilogues
                    if (++as == ACC_STAGES) { as = 0; aphase ^= 1u; }
                }
                pphase ^= 1u;
            }
        }
        __syncwarp();
    } else {
        // ===== epilogue: warp w owns TMEM lane quarter (w % 4) and column group (w - 2) / 4 of this CTA's 128 rows =====
        // 32x32b loads: lane l holds 32 consecutive columns of ONE row (TMEM lane 32q + l), so the row scalars are per
        // thread.  Results go through a per-warp 4 KB shared-memory box (32 rows x 128 B, 128B-swizzled: conflict-free
        // 16-byte writes) and leave as ONE TMA tensor store per box: full 128-byte lines, ragged edges clipped by the
        // tensor map.  (Direct st.global from the 16x256b fragment layout - 8 rows x 32 B per instruction - ran at one
        // L1 wavefront / L2 tag lookup per 32 bytes and added 0.54 ms to a 0.74 ms mainloop on config 5.)
        const int q = warp & 3;
        const int cg = (warp - 2) >> 2;
        const uint32_t stage_buf = stg_base + (uint32_t)(warp - 2) * SBYTES;      // this warp's staging box
        const uint32_t lead_tempty = mapa(tempty_bar(0), 0);
        int as = 0, mparity = 0;
        uint32_t aphase = 0, gphase = 0;
        constexpr int COLS = TN / kCG;   // 64 (32 for GEO)
        const uint32_t colc_w = colc_base + (uint32_t)(warp - 2) * (uint32_t)(kTileN / kCG) * 16u;   // this warp's column constants
        for (int64_t u = pair; u < units; u += npairs) {
          const int64_t tile = ARES ? u : u / S;
          const int sp = ARES ? 0 : (int)(u % S);
          const int64_t mt = ARES ? u : tile / n_tiles;
          const int64_t nt0 = ARES ? 0 : tile % n_tiles, nt1 = ARES ? n_tiles : nt0 + 1;
          const int64_t row_w = mt * kPairM + rank * BMC + q * 32;   // first row of this warp's 32
          const int64_t grow = row_w + lane;
          const bool rok = grow < prm.M;
          float rs = 1.0f, x2r = 0.0f, cf = 0.0f;
          if ((EPI == EPI_MOBIUS || (EPI == EPI_PLAIN && prm.rowscale)) && rok) rs = __ldg(prm.rowscale + grow);
          if ((EPI == EPI_GYRO || EPI == EPI_GYRO_LEAN || EPI == EPI_GEO || EPI == EPI_GYRO_BWD || EPI == EPI_MOBIUS_F) && rok)
              x2r = __ldg(prm.x2 + grow);
          if (EPI == EPI_MOBIUS_F) {
              // |M x_b|^2 = x_b^T G x_b: accumulator rows of x G dotted with the row of x the MMAs read (the resident bf16
              // panel: row r of k-block kb at a_tile(kb) + 128 r, 16-byte chunk ch at position ch ^ (r % 8)).  The 4 column
              // groups x g_tiles partial sums of a row meet in shared memory (the column-constant slices, unused here; two
              // buffers alternate between m-blocks) in a fixed order: deterministic.
              float* part = reinterpret_cast<float*>(smem_raw + (colc_base - raw)) + (mparity ? 2 * kCG * BMC : 0);
              const int rloc = q * 32 + lane;
              for (int t = 0; t < g_tiles; ++t) {
                  mbar_wait(tfull_bar(as), aphase);
                  tc_fence_after();
                  float accr = 0.0f;
#pragma unroll 1
                  for (int cb = cg * COLS; cb < (cg + 1) * COLS; cb += 32) {
                      float v[32];
                      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * kTileN + cb), v);
                      if (cb + 32 >= (cg + 1) * COLS) {
                          tc_fence_before();
                          __syncwarp();
                          if (lane == 0) mbar_arrive_cluster(lead_tempty + 8u * as);
                      }
                      const int j0 = t * kTileN + cb;           // contraction index of the chunk's first column
                      const uint32_t rowa = a_tile(j0 >> 6) + (uint32_t)rloc * 128u;
                      const int ch0 = (j0 & 63) >> 3;
#pragma unroll
                      for (int cc = 0; cc < 4; ++cc) {
                          if (j0 + 8 * cc < prm.K) {             // (K % 8 == 0; columns past K hold stale accumulators)
                              uint32_t w0, w1, w2, w3;
                              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                                           : "r"(rowa + (uint32_t)(((ch0 + cc) ^ (lane & 7)) << 4)) : "memory");
                              const uint32_t ww[4] = {w0, w1, w2, w3};
#pragma unroll
                              for (int e = 0; e < 4; ++e) {
                                  accr = fmaf(v[8 * cc + 2 * e], __uint_as_float(ww[e] << 16), accr);
                                  accr = fmaf(v[8 * cc + 2 * e + 1], __uint_as_float(ww[e] & 0xffff0000u), accr);
                              }
                          }
                      }
                  }
                  part[(t * kCG + cg) * BMC + rloc] = accr;
                  if (++as == ACC_STAGES) { as = 0; aphase ^= 1u; }
              }
              asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");   // the epilogue warps of this CTA only
              float mx2 = 0.0f;
              for (int i = 0; i < g_tiles * kCG; ++i) mx2 += part[i * BMC + rloc];
              mx2 = fmaxf(mx2, 0.0f);
              if (prm.mxsq_out && cg == 0 && rok) prm.mxsq_out[grow] = mx2;
              // y_b = rs (M x_b): geoopt mobius_matvec's rescale, then project (k_mobius_rowscale's arithmetic)
              const float xn = fmaxf(sqrtf(x2r), kMinNorm);
              const float mxn_raw = sqrtf(mx2), mxn = fmaxf(mxn_raw, kMinNorm);
              const float tt = tanh_c(mxn / xn * artanh_c(prm.gp.sc * xn));
              rs = prm.gp.rsc * tt / mxn;
              const float yn = fmaxf(prm.gp.rsc * tt * (mxn_raw / mxn), kMinNorm);
              if (yn > prm.gp.maxnorm) rs = rs / yn * prm.gp.maxnorm;
              if (mx2 == 0.0f) rs = 0.0f;
              mparity ^= 1;
          }
          const bool axpy = (EPI == EPI_PLAIN) && prm.axpy_x != nullptr;
          if (axpy && rok) cf = prm.axpy_coef ? __ldg(prm.axpy_coef + grow) : 1.0f;
          for (int64_t nt = nt0; nt < nt1; ++nt) {
            float accr = 0.0f;
            if (EPI == EPI_GYRO || EPI == EPI_GYRO_LEAN || EPI == EPI_GEO || EPI == EPI_GYRO_BWD) {
                // per-column constants of this warp's COLS columns of the tile, in the warp's OWN shared-memory slice: no
                // block-wide barrier (a bar.sync of all 16 epilogue warps per tile was 2 of ~11 stall cycles per issue)
                __syncwarp();   // (the warp's previous tile no longer reads the slice)
                for (int j = lane; j < COLS; j += 32) {
                    const int64_t n = nt * TN + cg * COLS + j;
                    const bool in = n < prm.N;
                    const float p2 = in ? __ldg(prm.p2 + n) : 0.0f;
                    const float bb = (prm.bias && in) ? __ldg(prm.bias + n) : 0.0f;
                    float4 cv;
                    if (EPI == EPI_GYRO || EPI == EPI_GYRO_LEAN || EPI == EPI_GYRO_BWD) {
                        // lean path constants (see below): u = r (1 + c|p|^2), v = r |p|^2, r = 2 sqrt(c) / ((1 - c|p|^2)|p|)
                        const float c = prm.gp.c, pn = sqrtf(p2);
                        const float rcol = 2.0f * prm.gp.sc / ((1.0f - c * p2) * pn + kMinNorm);
                        // z: the bias (forward) / v/u = |p|^2 / (1 + c|p|^2) (backward: weight of the row sums)
                        cv = make_float4(rcol * (1.0f + c * p2), rcol * p2, EPI == EPI_GYRO_BWD ? p2 / (1.0f + c * p2) : bb, p2);
                    } else {
                        cv = make_float4(p2, in ? __ldg(prm.pa + n) : 0.0f, in ? __ldg(prm.an + n) : 0.0f, bb);
                    }
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(colc_w + 16u * (uint32_t)j), "f"(cv.x), "f"(cv.y),
                                 "f"(cv.z), "f"(cv.w)
                                 : "memory");
                }
                __syncwarp();
            }
            mbar_wait(tfull_bar(as), aphase);
            tc_fence_after();
#pragma unroll 1
            for (int cb = cg * COLS; cb < (cg + 1) * COLS; cb += 32) {
                const bool last = cb + 32 >= (cg + 1) * COLS;
                float v[32];
#ifdef HVAE_EXPERIMENT
                if (prm.dbg & 2) {
                    if (last) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive_cluster(lead_tempty + 8u * as); }
                    continue;
                }
#endif
                const int64_t n0 = nt * TN + cb;
                const bool inb = row_w < prm.M && n0 < prm.N;   // (warp-uniform) the box touches the matrix
                if (EPI == EPI_GYRO_BWD && inb) {
                    // the upstream-gradient box of this chunk: TMA -> this warp's staging buffer (its previous store has
                    // been read out), overlapping the accumulator load below
                    if (lane == 0) {
                        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        mbar_expect_tx(gload_bar(warp - 2), STG_BYTES);
                        tma_load_2d(stage_buf, &map_g, gload_bar(warp - 2), (int)n0, (int)row_w);
                    }
                }
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * kTileN + cb), v);
                // per-column constants: explicit shared-space loads on a 32-bit address (every lane reads the same entry)
                const uint32_t cc_addr = colc_w + (uint32_t)(cb - cg * COLS) * 16u;
                auto ldcc = [&](int i) {
                    float4 r;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(cc_addr + 16u * (uint32_t)i));
                    return r;
                };
                if (EPI == EPI_GEO) {
                    // columns cb .. cb+31 hold <x,p_j>, columns 128 + cb .. hold <x,a_j> of the same planes; in two halves
                    // of 16 planes to bound the live registers
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        float w[16];
                        tmem_ld16x(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * kTileN + TN + cb + 16 * hf), w);
                        if (last && hf == 1) {   // the accumulator stage is in registers: hand it back before the math
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive_cluster(lead_tempty + 8u * as);
                        }
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float4 cc = ldcc(16 * hf + i);  // {p2, <p,a>, |a|, bias}
                            GyroPairCtx kk;
                            v[16 * hf + i] = gyro_pair_fwd(v[16 * hf + i], w[i], x2r, cc.x, cc.y, cc.z, prm.gp, kk) + cc.w;
                        }
                    }
                } else if (last) {   // the accumulator stage is in registers: hand it back before the math and the stores
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(lead_tempty + 8u * as);
                }
                if (EPI == EPI_PLAIN) {
                    if (prm.rowsq) {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (n0 + i < prm.N) accr = fmaf(v[i], v[i], accr);   // (columns past a ragged tile's MMA width are stale)
                    }
                    if (prm.rowscale) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] *= rs;
                    }
                    if (axpy && rok) {
                        const float* xr = prm.axpy_x + grow * prm.N + n0;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            if (n0 + 4 * i + 4 <= prm.N) {   // N % 4 == 0 on this path
                                const float4 xv = __ldg(reinterpret_cast<const float4*>(xr + 4 * i));
                                v[4 * i] = fmaf(cf, xv.x, v[4 * i]);
                                v[4 * i + 1] = fmaf(cf, xv.y, v[4 * i + 1]);
                                v[4 * i + 2] = fmaf(cf, xv.z, v[4 * i + 2]);
                                v[4 * i + 3] = fmaf(cf, xv.w, v[4 * i + 3]);
                            }
                        }
                    }
                } else if (EPI == EPI_MOBIUS || EPI == EPI_MOBIUS_F) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] *= rs;
                } else if (EPI == EPI_ROWDOT) {
                    if (rok) {
                        const float* xr = prm.xrow + grow * prm.N + n0;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            if (n0 + 4 * i + 4 <= prm.N) {
                                const float4 xv = __ldg(reinterpret_cast<const float4*>(xr + 4 * i));
                                accr = fmaf(v[4 * i], xv.x, fmaf(v[4 * i + 1], xv.y, fmaf(v[4 * i + 2], xv.z, fmaf(v[4 * i + 3], xv.w, accr))));
                            }
                        }
                    }
                } else if (EPI == EPI_GYRO || EPI == EPI_GYRO_LEAN) {
                    // geoopt signed distance with a == p.  With z = (-p) (+) x,
                    //   <z, p> = (Bc <x,p> - A |p|^2) / den   and   1 - c|z|^2 = Bc (1 - c|x|^2) / den     (Bc = 1 - c|p|^2)
                    // so den cancels and asinh's argument separates into row and column factors around <x,p>:
                    //   y = 2 sqrt(c) [<x,p>(1 + c|p|^2) - |p|^2 (1 + c|x|^2)] / (Bc |p| (1 - c|x|^2)) = a_b (px u_j - w_b v_j)
                    // (the clamps of the reference only bind for |p| ~ 1e-15, kept via the + MIN_NORM in u, v).  Other flag
                    // combinations take the general pair function.
                    constexpr bool lean = EPI == EPI_GYRO_LEAN;   // (two instantiations: the general pair function costs registers)
                    const float rsc_ln2 = prm.gp.rsc * 0.693147180559945f;
                    const float ar = 1.0f / fmaxf(1.0f - prm.gp.c * x2r, 1e-30f);
                    const float wr = -(1.0f + prm.gp.c * x2r);
                    if (lean) {
                        // row factors folded: out = log2(|y| + sqrt(y^2 + 1)) sign(y) (ln2 / sqrt(c)) + bias,  y = px (a_b u_j) + (a_b w_b) v_j
                        const float aw = ar * wr;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float4 cc = ldcc(i);  // {u, v, bias, p2}
                            const float y = fmaf(v[i] * ar, cc.x, aw * cc.y);
                            v[i] = fmaf(asinh_lg2(y), rsc_ln2, cc.z);
                        }
                    } else {
#pragma unroll 4
                        for (int i = 0; i < 32; ++i) {
                            const float4 cc = ldcc(i);
                            GyroPairCtx kk;
                            v[i] = gyro_pair_fwd(v[i], v[i], x2r, cc.w, cc.w, sqrtf(cc.w), prm.gp, kk) + cc.z;
                        }
                    }
                }
                if (EPI == EPI_GYRO_BWD) {
                    // Backward of the lean forward above (a == p, signed, no other flag): with y = a_b (px u_j + w_b v_j),
                    //   dL/dpx = g rsc a_b u_j / sqrt(1 + y^2)      (bf16: the A operand of the two gradient GEMMs).
                    // Everything else follows from CP = dL/dpx by linear algebra (tc_gemm.cu, hvae_gyroplane_tc_bwd_f32):
                    // the row term sum_j dL/dy (y - v_j) = <x_b, (CP p)_b> - 2 (CP (v/u))_b and the column terms
                    // sum_b dL/dy a_b px = <p_j, (CP^T x)_j> / u_j,  sum_b dL/dy a_b w_b = (CP^T w)_j / u_j  ride on the
                    // gradient GEMMs as augmented columns - the epilogue does no reductions.
                    uint32_t pk[16];
                    if (inb) {
                        mbar_wait(gload_bar(warp - 2), gphase);
                        gphase ^= 1u;
                        const float ar = 1.0f / fmaxf(1.0f - prm.gp.c * x2r, 1e-30f);
                        const float wr = -(1.0f + prm.gp.c * x2r);
                        const float ra = prm.gp.rsc * ar;
                        const uint32_t rowp = stage_buf + (uint32_t)lane * 128u;
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            float g0, g1, g2, g3;
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                         : "=f"(g0), "=f"(g1), "=f"(g2), "=f"(g3)
                                         : "r"(rowp + (uint32_t)((c ^ (lane & 7)) << 4))
                                         : "memory");
                            const float gg[4] = {g0, g1, g2, g3};
                            float o[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float4 cc = ldcc(4 * c + e);  // {u, v, bias, p2}
                                const float y = ar * fmaf(v[4 * c + e], cc.x, wr * cc.y);
                                o[e] = (gg[e] * rsqrtf(fmaf(y, y, 1.0f))) * (ra * cc.x);
                                accr = fmaf(o[e], cc.z, accr);        // row sum of CP v/u
                                v[4 * c + e] = o[e] * wr;             // column-sum term CP w_b (reduced across the warp below)
                            }
                            __nv_bfloat162 lo = __floats2bfloat162_rn(o[0], o[1]), hi = __floats2bfloat162_rn(o[2], o[3]);
                            pk[2 * c] = *reinterpret_cast<uint32_t*>(&lo);
                            pk[2 * c + 1] = *reinterpret_cast<uint32_t*>(&hi);
                        }
                        // column sums over the warp's 32 rows: butterfly transpose-reduce, lane l ends with column n0 + l
#pragma unroll
                        for (int sft = 16; sft >= 1; sft >>= 1) {
                            const bool up = (lane & sft) != 0;
#pragma unroll
                            for (int i = 0; i < sft; ++i) {
                                const float send = up ? v[i] : v[i + sft];
                                const float keep = up ? v[i + sft] : v[i];
                                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, sft);
                            }
                        }
                        if (n0 + lane < prm.N) prm.vcol[(row_w >> 5) * prm.N + n0 + lane] = v[0];
                        __syncwarp();   // every lane has read its g row: the buffer may be overwritten
                        // bf16 box: 32 rows x 64 B, 16-byte chunk c at position c ^ ((row / 2) % 4)  (SWIZZLE_64B)
                        const uint32_t rowq = stage_buf + (uint32_t)lane * 64u;
                        const int swz = (lane >> 1) & 3;
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowq + (uint32_t)((c ^ swz) << 4)), "r"(pk[4 * c]),
                                         "r"(pk[4 * c + 1]), "r"(pk[4 * c + 2]), "r"(pk[4 * c + 3])
                                         : "memory");
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) {
                            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                                         ::"l"(&map_d), "r"(stage_buf), "r"((int)n0), "r"((int)row_w), "r"(0)
                                         : "memory");
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                    }
                    continue;
                }
                if (EPI != EPI_ROWDOT) {
#ifdef HVAE_EXPERIMENT
                    if (prm.dbg & 1) {
                        float t = 0.0f;
#pragma unroll
                        for (int i = 0; i < 32; ++i) t += v[i];
                        if (t == 123.456f) prm.D[0] = t;
                        continue;
                    }
#endif
                    // SB_COLS-wide boxes: row `lane`, 16-byte chunk c at position c ^ swz(lane) - the pattern of
                    // CU_TENSOR_MAP_SWIZZLE_128B (32 columns: c ^ (row % 8)) / SWIZZLE_64B (16 columns: c ^ ((row / 2) % 4))
#pragma unroll
                    for (int sb = 0; sb < 32 / SB_COLS; ++sb) {
#ifdef HVAE_EXPERIMENT
                        if (prm.dbg & 8) {
                            // experiment: the same staging box, drained by coalesced st.global instead of a TMA store
                            constexpr int CPR = SB_COLS / 4, RPI = 32 / CPR;   // 16-byte chunks per row, rows per instruction
                            __syncwarp();
                            {
                                const uint32_t rowp = stage_buf + (uint32_t)lane * (uint32_t)(SB_COLS * 4);
                                const int swz = ARES ? ((lane >> 1) & 3) : (lane & 7);
#pragma unroll
                                for (int c = 0; c < CPR; ++c) {
                                    const int e0 = sb * SB_COLS + 4 * c;
                                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(rowp + (uint32_t)((c ^ swz) << 4)),
                                                 "f"(v[e0]), "f"(v[e0 + 1]), "f"(v[e0 + 2]), "f"(v[e0 + 3]) : "memory");
                                }
                            }
                            __syncwarp();
                            const int64_t nb = n0 + sb * SB_COLS;
                            float* dpl = prm.D + (int64_t)sp * prm.M * prm.N;
#pragma unroll
                            for (int it = 0; it < 32 / RPI; ++it) {
                                const int r = it * RPI + lane / CPR, c = lane % CPR;
                                const int swz = ARES ? ((r >> 1) & 3) : (r & 7);
                                float4 o;
                                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w)
                                             : "r"(stage_buf + (uint32_t)r * (uint32_t)(SB_COLS * 4) + (uint32_t)((c ^ swz) << 4)) : "memory");
                                const int64_t gr = row_w + r, gc = nb + 4 * c;
                                if (gr < prm.M && gc + 4 <= prm.N) {
                                    float4* dst = reinterpret_cast<float4*>(dpl + gr * prm.N + gc);
                                    if (prm.dbg & 32) __stcs(dst, o); else *dst = o;
                                }
                            }
                            continue;
                        }
#endif
                        // the previous box of this warp must have been read out of shared memory by its TMA store
#ifdef HVAE_EXPERIMENT
                        if (!(prm.dbg & 16))
#endif
                        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        __syncwarp();
                        const uint32_t rowp = stage_buf + (uint32_t)lane * (uint32_t)(SB_COLS * 4);
                        const int swz = ARES ? ((lane >> 1) & 3) : (lane & 7);
#pragma unroll
                        for (int c = 0; c < SB_COLS / 4; ++c) {
                            const uint32_t addr = rowp + (uint32_t)((c ^ swz) << 4);
                            const int e0 = sb * SB_COLS + 4 * c;
                            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v[e0]), "f"(v[e0 + 1]),
                                         "f"(v[e0 + 2]), "f"(v[e0 + 3])
                                         : "memory");
                        }
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the TMA engine
                        __syncwarp();
                        const int64_t nb = n0 + sb * SB_COLS;
                        if (lane == 0 && row_w < prm.M && nb < prm.N) {   // (a box wholly outside the matrix is not issued)
                            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                                         ::"l"(&map_d), "r"(stage_buf), "r"((int)nb), "r"((int)row_w), "r"(sp)
                                         : "memory");
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                    }
                }
            }
            // per-(row, n-tile, column-group) partials: [(nt * kCG + cg)][M]
            if ((EPI == EPI_PLAIN && prm.rowsq) || EPI == EPI_ROWDOT || EPI == EPI_GYRO_BWD) {
                float* dstp = (EPI == EPI_PLAIN) ? prm.rowsq : (EPI == EPI_ROWDOT ? prm.rowdot : prm.srow);
                if (rok) dstp[(nt * kCG + cg) * prm.M + grow] = accr;
            }
            if (++as == ACC_STAGES) { as = 0; aphase ^= 1u; }
          }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory stays valid until read
        __syncwarp();
    }
    tc_fence_before();
    cluster_sync_all();   // both CTAs are done with the pair's tensor memory and with each other's barriers
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
}

// fp32 (S, M, N) row-major output -> 3-D tensor map with a {32, 32, 1} box (one 128-byte row segment x 32 rows), 128B swizzle
static bool make_map_out(CUtensorMap* m, float* D, int64_t S, int64_t M, int64_t N, bool ares) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)M, (cuuint64_t)S};
    const cuuint64_t strides[2] = {(cuuint64_t)N * 4, (cuuint64_t)N * (cuuint64_t)M * 4};
    const cuuint32_t box[3] = {ares ? 16u : 32u, 32, 1};   // A-resident schedule: 2 KB boxes (64B swizzle)
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, D, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               ares ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// GYRO_BWD: bf16 (M, N) output through {32, 32, 1} boxes (64-byte rows, 64B swizzle) and the fp32 (M, N) upstream
// gradient read through {32, 32} boxes (128-byte rows, 128B swizzle)
static bool make_map_bwd(CUtensorMap* md, CUtensorMap* mg, __nv_bfloat16* D16, const float* g, int64_t M, int64_t N) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return false;
    const cuuint32_t estr[3] = {1, 1, 1};
    {
        const cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)M, 1};
        const cuuint64_t strides[2] = {(cuuint64_t)N * 2, (cuuint64_t)N * (cuuint64_t)M * 2};
        const cuuint32_t box[3] = {32, 32, 1};
        if (enc(md, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, D16, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
    const cuuint64_t strides[1] = {(cuuint64_t)N * 4};
    const cuuint32_t box[2] = {32, 32};
    return enc(mg, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(g), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int EPI>
static int launch_t(const __nv_bfloat16* A, const __nv_bfloat16* B, const __nv_bfloat16* B2, const Params2& prm, cudaStream_t s) {
    CUtensorMap ma, mb, mb2, md, mg;
    // K-major operand: (rows, K) buffer, {64 k, 128 rows} boxes; MN-major: (K, rows) buffer, {64, 64} boxes
    const bool okA = prm.a_mn ? make_map(&ma, A, prm.K, prm.M, 64, prm.a_pitch) : make_map(&ma, A, prm.M, prm.K, BMC, prm.a_pitch);
    const bool okB = prm.b_mn ? make_map(&mb, B, prm.K, prm.N, 64, prm.b_pitch) : make_map(&mb, B, prm.N, prm.K, BNH, prm.b_pitch);
    if (!okA || !okB) return HVAE_ELAUNCH;
    if (!make_map(&mb2, B2 ? B2 : B, prm.N, prm.K, BNH, prm.b_pitch)) return HVAE_ELAUNCH;
    constexpr int TN = (EPI == EPI_GEO) ? kTileN / 2 : kTileN;
    const int64_t m_tiles = (prm.M + kPairM - 1) / kPairM, n_tiles = (prm.N + TN - 1) / TN;
    const int S = prm.splits > 1 ? prm.splits : 1;
    // A-resident when the panel fits (K <= 512), there are several n-tiles to amortise it over and enough m-blocks to
    // keep every pair busy (the unit of work becomes a whole m-block)
    bool ares = S == 1 && !prm.a_mn && !prm.b_mn && EPI != EPI_GYRO_BWD && prm.K <= (int64_t)ARES_KB * BK && n_tiles >= 4 &&
                m_tiles >= 2 * (kNumSMs / 2);
#ifdef HVAE_EXPERIMENT
    if (prm.dbg & 4) ares = false;
#endif
    if (EPI == EPI_MOBIUS_F) {
        if (!ares) return HVAE_ESHAPE;
        if (!make_map(&mb2, B2, prm.K, prm.K, BNH, 0)) return HVAE_ELAUNCH;   // G: (K, K) bf16
    }
    // output: fp32 (S, M, N), boxes of 32 rows x 32 (16) columns through swizzled staging; ROWDOT writes no matrix
    mg = ma;
    if (EPI == EPI_GYRO_BWD) {
        if (!make_map_bwd(&md, &mg, prm.D16, prm.g, prm.M, prm.N)) return HVAE_ELAUNCH;
    } else if (EPI != EPI_ROWDOT) {
        if (!make_map_out(&md, prm.D, S, prm.M, prm.N, ares)) return HVAE_ELAUNCH;
    } else {
        md = ma;
    }
    // per-device attribute (a process may drive several GPUs): set on every launch, a cheap driver call
    if (EPI != EPI_MOBIUS_F)
        cudaFuncSetAttribute(k_tc_gemm2 < EPI == EPI_MOBIUS_F ? EPI_MOBIUS : EPI, false >, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)SMEM_BYTES);
    cudaFuncSetAttribute(k_tc_gemm2<EPI, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES_ARES);
    const int64_t units = ares ? m_tiles : m_tiles * n_tiles * S;
    const int pairs = (int)(units < kNumSMs / 2 ? units : kNumSMs / 2);
    if (ares)
        k_tc_gemm2<EPI, true><<<2 * pairs, THREADS, SMEM_BYTES_ARES, s>>>(ma, mb, mb2, md, mg, prm);
    else if (EPI != EPI_MOBIUS_F)
        k_tc_gemm2 < EPI == EPI_MOBIUS_F ? EPI_MOBIUS : EPI, false ><<<2 * pairs, THREADS, SMEM_BYTES, s>>>(ma, mb, mb2, md, mg, prm);
    return check_launch();
}

int launch_gemm2(int epi, const __nv_bfloat16* A, const __nv_bfloat16* B, const __nv_bfloat16* B2, const Params2& prm,
                 cudaStream_t s) {
    if (prm.M <= 0 || prm.N <= 0 || prm.K <= 0 || (prm.K % 8) != 0) return HVAE_ESHAPE;
    if (epi != EPI_ROWDOT && (prm.N % 4) != 0) return HVAE_ESHAPE;   // TMA store: 16-byte row pitch
    if ((prm.a_mn && (prm.M % 8)) || (prm.b_mn && (prm.N % 8))) return HVAE_ESHAPE;   // MN-major buffers: 16-byte row pitch
    const int k_blocks = (int)((prm.K + BK - 1) / BK);
    if (prm.splits > k_blocks) return HVAE_EARG;
    if (prm.splits > 1 && (epi != EPI_PLAIN || prm.rowsq || prm.axpy_x || prm.rowscale)) return HVAE_EARG;
    switch (epi) {
        case EPI_PLAIN: return launch_t<EPI_PLAIN>(A, B, nullptr, prm, s);
        case EPI_GYRO:
            if (prm.gp.flags == (uint32_t)HVAE_GYRO_SIGNED) return launch_t<EPI_GYRO_LEAN>(A, B, nullptr, prm, s);
            return launch_t<EPI_GYRO>(A, B, nullptr, prm, s);
        case EPI_ROWDOT: return launch_t<EPI_ROWDOT>(A, B, nullptr, prm, s);
        case EPI_MOBIUS: return prm.rowscale ? launch_t<EPI_MOBIUS>(A, B, nullptr, prm, s) : HVAE_EARG;
        case EPI_MOBIUS_F: return (B2 && prm.x2) ? launch_t<EPI_MOBIUS_F>(A, B, B2, prm, s) : HVAE_EARG;
        case EPI_GEO: return B2 ? launch_t<EPI_GEO>(A, B, B2, prm, s) : HVAE_EARG;
        case EPI_GYRO_BWD:
            return (prm.g && prm.D16 && prm.srow && prm.vcol && (prm.N % 8) == 0) ? launch_t<EPI_GYRO_BWD>(A, B, nullptr, prm, s) : HVAE_EARG;
    }
    return HVAE_EARG;
}

}  // namespace tc2
}  // namespace hvae

#ifdef HVAE_EXPERIMENT
// experiment build only (libhvae_b200_exp.so, scripts/tc2_probe.py): the pair kernel alone on bf16 operands
extern "C" int hvae_exp_gemm2_bf16(const void* A, const void* B, float* D, const float* rowscale, int64_t M, int64_t N, int64_t K,
                                   int epi, int dbg, void* stream) {
    hvae::tc2::Params2 q{};
    q.D = D; q.M = M; q.N = N; q.K = K; q.splits = 1; q.rowscale = rowscale; q.dbg = dbg;
    return hvae::tc2::launch_gemm2(epi, (const __nv_bfloat16*)A, (const __nv_bfloat16*)B, nullptr, q, (cudaStream_t)stream);
}
#endif
