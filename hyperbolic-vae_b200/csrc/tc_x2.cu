// f-3: the Euclidean trunk's dense layers - fp32 in, fp32 out, fp32 accuracy - on the fp16 tensor-core path.
//
// reference: every nn.Linear of the encoders / decoders (hyperbolic_vae/models/*.py; pvae Enc / Dec of config 2) and its
// autograd (input gradient gy W, weight gradient gy^T x).
//
// Arithmetic.  Each operand row r is scaled by a power of two 2^e_r that brings its largest magnitude into [2^14, 2^15)
// (exact), then split into TWO fp16 pieces: hi = rn(v), lo = rn(v - hi) (the subtraction is exact in fp32).  fp16 carries
// 11 significand bits, so hi + lo represents v to 2^-22 relative; elements more than 2^28 below their row's maximum fall
// into fp16's subnormal range and keep an ABSOLUTE error of 2^-39 of that maximum.  The product
//     a b  ~=  a_lo b_hi + a_hi b_lo + a_hi b_hi          (a_lo b_lo ~ 2^-22 |a b| dropped)
// is three tcgen05 kind::f16 MMAs per k-step, smallest first, accumulated in fp32 tensor memory, and the epilogue undoes
// the scales (2^-(e_m + e_n), exact).  Per-term error <= 3 * 2^-22 ~= 7e-7, random in sign: ~1e-7 of the output scale
// after the contraction - tighter than the three-way bf16 split (six products) this replaces, at half the MMA work and
// two thirds of the operand bytes.  The scale runs along the operand's NON-contracted index, so it factors out of the
// sum; a tensor used with the contraction along its rows (weight gradients: contraction over the batch) is split
// transposed with per-COLUMN scales (hvae_split2h_both_f32).
//
// The tensor core adds into its fp32 accumulator with truncation, an error that grows with the length of the chain
// (tc_gemm.cu measured 6e-6 relative at K = 4096), so the MMA warp hands the accumulator over every `hand` k-positions
// (2 x 64 contraction elements x 3 products = 384 terms) and the epilogue warps sum the hand-overs in registers with
// round-to-nearest adds; two TMEM stages pipeline chunk i + 1 under the drain of chunk i.
//
// Tiling.  cta_group::1, 128 rows x bn columns per CTA with bn in {128, 160, 192, 224, 256} chosen per problem together
// with the split-K factor so that the units fill the 148 SMs in as few waves as possible (config 2: 4096 x 600 output ->
// bn = 160: 128 units, ONE wave; the fixed 128 x 128 tiling made 160 tiles = two waves, the second 8 % full).  A stage
// of the TMA ring holds the hi and lo tiles of A and B for one k-position; the three products are issued from it.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>

#include "tc_common.cuh"

namespace hvae {
namespace x2 {

using namespace hvae::tc;

constexpr int BM = 128, BK = kTcBK, UMMA_K = 16;
constexpr int CG = 4, EPI_WARPS = 4 * CG, THREADS = 64 + 32 * EPI_WARPS;
constexpr int ACC_STAGES = 2, ACC_COLS = 256;            // two accumulator stages of up to 256 columns: all of tensor memory
constexpr uint32_t TILE_A = BM * BK * 2;                 // 16 KB
constexpr uint32_t SMEM_LIMIT = 232448, BAR_BYTES = 256;
constexpr uint32_t STG_WARP = 16 * 128, STG_BYTES = EPI_WARPS * STG_WARP;   // 16 rows x 32 fp32 per epilogue warp
constexpr int MAX_STAGES = 4, MAX_SPLITS = 8;

// 16 lanes x (4 repeats of 256 bits) -> 16 registers, complete on return
__device__ __forceinline__ void tmem_ld16_sync(uint32_t addr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(addr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

struct Params {
    float* D;                 // (M, N) row-major (split-K: S partial planes)
    int64_t M, N;
    int kpos;                 // k-positions of 64 contraction elements (Kp / 64)
    int bn, splits, nst, hand;
    int a_lo, b_lo;           // column offset of the lo piece in the operand buffers (= Kp)
    const float* sa;          // (M,) 2^-e of the A rows
    const float* sb;          // (N,) 2^-e of the B rows
    const float* bias;        // optional (N,), S == 1 only
    int relu;
#ifdef HVAE_EXPERIMENT
    int dbg;                  // experiment build: 1 = no global stores
    long long* ts;            // experiment build: clock64 stamps of CTA 0 ([0,64) producer, [64,128) MMA, [128,192) epilogue warp 2)
#endif
};
#ifdef HVAE_EXPERIMENT
#define X2_STAMP(slot) do { if (prm.ts && blockIdx.x == 0 && (slot) < 64) prm.ts[(slot) + tsb] = clock64(); } while (0)
#else
#define X2_STAMP(slot) do { } while (0)
#endif

__global__ void __launch_bounds__(THREADS, 1)
k_x2_gemm(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, Params prm) {
    extern __shared__ uint8_t smem_raw[];
#ifdef HVAE_EXPERIMENT
    if (prm.ts && threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        prm.ts[190] = (long long)t;       // first instruction of CTA 0
        prm.ts[191] = clock64();
    }
#endif
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t tile_b = (uint32_t)prm.bn * (BK * 2);
    const uint32_t stage_bytes = 2u * TILE_A + 2u * tile_b;
    const int NST = prm.nst;
    const uint32_t stg_base = base + (uint32_t)NST * stage_bytes;          // per-warp staging boxes of the finish
    const uint32_t bars = stg_base + STG_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (MAX_STAGES + s); };
    auto tfull_bar = [&](int s) { return bars + 8u * (2 * MAX_STAGES + s); };
    auto tempty_bar = [&](int s) { return bars + 8u * (2 * MAX_STAGES + ACC_STAGES + s); };
    const uint32_t tmem_slot = bars + 8u * (2 * MAX_STAGES + 2 * ACC_STAGES);
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));
    auto a_hi = [&](int s) { return base + (uint32_t)s * stage_bytes; };
    auto b_hi = [&](int s) { return base + (uint32_t)s * stage_bytes + 2u * TILE_A; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // (32-bit unit arithmetic: 64-bit divisions are subroutine calls, ~1.3k cycles of them sat in front of the first TMA load;
    //  the host checks that the unit count fits)
    const int m_tiles = (int)((prm.M + BM - 1) / BM), n_tiles = (int)((prm.N + prm.bn - 1) / prm.bn);
    const int S = prm.splits > 1 ? prm.splits : 1;
    const int units = m_tiles * n_tiles * S;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        for (int s = 0; s < NST; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < ACC_STAGES; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // the first ring of loads leaves now, under the tensor-memory allocation and the block-wide sync below (the
        // barriers it uses were initialised by this very thread; their other users arrive after the sync)
        const int u = blockIdx.x;
        if (u < units) {
            const int tile = u / S, sp = u - tile * S;
            const int p0 = sp * prm.kpos / S, p1 = (sp + 1) * prm.kpos / S;
            const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * prm.bn;
            for (int kk = p0; kk < p1 && kk < p0 + NST; ++kk) {
                const int stage = kk - p0;
                mbar_expect_tx(full_bar(stage), stage_bytes);
                tma_load_2d(a_hi(stage), &map_a, full_bar(stage), kk * BK, m0);
                tma_load_2d(a_hi(stage) + TILE_A, &map_a, full_bar(stage), prm.a_lo + kk * BK, m0);
                tma_load_2d(b_hi(stage), &map_b, full_bar(stage), kk * BK, n0);
                tma_load_2d(b_hi(stage) + tile_b, &map_b, full_bar(stage), prm.b_lo + kk * BK, n0);
            }
        }
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(ACC_STAGES * ACC_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
#ifdef HVAE_EXPERIMENT
    if (prm.ts && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        prm.ts[192 + 3 * blockIdx.x] = (long long)t;
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        prm.ts[192 + 3 * blockIdx.x + 2] = smid;
    }
#endif

    if (warp == 0) {
        // ===== TMA producer: {A hi, A lo, B hi, B lo} of one k-position per stage =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            [[maybe_unused]] const int tsb = 0;
            [[maybe_unused]] int tsn = 0;
            X2_STAMP(tsn++);
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                const int tile = u / S;
                const int sp = u - tile * S;
                const int p0 = sp * prm.kpos / S, p1 = (sp + 1) * prm.kpos / S;
                const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * prm.bn;
                for (int kk = p0; kk < p1; ++kk) {
                    if (u == (int)blockIdx.x && kk < p0 + NST) {   // issued in the prologue
                        if (++stage == NST) { stage = 0; phase ^= 1u; }
                        continue;
                    }
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    X2_STAMP(tsn++);
                    mbar_expect_tx(full_bar(stage), stage_bytes);
                    tma_load_2d(a_hi(stage), &map_a, full_bar(stage), kk * BK, m0);
                    tma_load_2d(a_hi(stage) + TILE_A, &map_a, full_bar(stage), prm.a_lo + kk * BK, m0);
                    tma_load_2d(b_hi(stage), &map_b, full_bar(stage), kk * BK, n0);
                    tma_load_2d(b_hi(stage) + tile_b, &map_b, full_bar(stage), prm.b_lo + kk * BK, n0);
                    if (++stage == NST) { stage = 0; phase ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            int stage = 0, as = 0;
            uint32_t phase = 0, aphase = 0;
            // kind::f16 instruction descriptor: D = f32, A = B = f16 (format 0), both K-major, M = 128, N = bn
            const uint32_t idesc = (1u << 4) | ((uint32_t)(prm.bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            [[maybe_unused]] const int tsb = 64;
            [[maybe_unused]] int tsn = 0;
            X2_STAMP(tsn++);
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                const int sp = u % S;
                const int p0 = sp * prm.kpos / S, p1 = (sp + 1) * prm.kpos / S;
                int g = 0;
                for (int kk = p0; kk < p1; ++kk) {
                    if (g == 0) {
                        mbar_wait(tempty_bar(as), aphase ^ 1u);   // the epilogue has drained this accumulator stage
                        tc_fence_after();
                    }
                    const uint32_t tmem_d = tmem_base + (uint32_t)(as * ACC_COLS);
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    X2_STAMP(tsn++);
                    const uint64_t dah = make_desc(a_hi(stage)), dal = make_desc(a_hi(stage) + TILE_A);
                    const uint64_t dbh = make_desc(b_hi(stage)), dbl = make_desc(b_hi(stage) + tile_b);
                    // smallest products first
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) umma(tmem_d, dal + (uint64_t)(2 * k), dbh + (uint64_t)(2 * k), idesc, (uint32_t)((g | k) != 0));
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) umma(tmem_d, dah + (uint64_t)(2 * k), dbl + (uint64_t)(2 * k), idesc, 1u);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) umma(tmem_d, dah + (uint64_t)(2 * k), dbh + (uint64_t)(2 * k), idesc, 1u);
                    umma_commit(empty_bar(stage));
                    if (++stage == NST) { stage = 0; phase ^= 1u; }
                    if (++g == prm.hand || kk == p1 - 1) {
                        umma_commit(tfull_bar(as));   // hand the chunk over
                        g = 0;
                        if (++as == ACC_STAGES) { as = 0; aphase ^= 1u; }
                    }
                }
            }
        }
        __syncwarp();
    } else {
        // ===== epilogue: warp w owns TMEM lane quarter (w % 4) and the 32-column chunks cg, cg + 4 of the tile =====
        // 16x256b loads: for each 8-column repeat i, lane t holds columns 8i + 2(t % 4) + {0, 1} of rows t / 4 and t / 4 + 8
        // of a 16-row half: four neighbouring lanes own one 32-byte sector of an output row, the stores leave
        // sector-complete straight from registers.
        const int q = warp & 3, cg = (warp - 2) >> 2;
        const int lr = lane >> 2, lc = (lane & 3) * 2;
        const int nchunks = prm.bn >> 5;
        const bool two = cg + CG < nchunks;   // (warp-uniform) this warp owns a second chunk
        int as = 0;
        uint32_t aphase = 0;
        [[maybe_unused]] const int tsb = 128;
        [[maybe_unused]] int tsn = (warp == 2 && lane == 0) ? 0 : 64;
        X2_STAMP(tsn++);
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            const int tile = u / S;
            const int sp = u - tile * S;
            const int p0 = sp * prm.kpos / S, p1 = (sp + 1) * prm.kpos / S;
            const int mt = tile / n_tiles, nt = tile - mt * n_tiles;
            float acc[2][2][16];
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[j][0][i] = acc[j][1][i] = 0.0f;
            // inverse scales of the thread's four rows (fragment layout), fetched while the first chunk is still being multiplied
            const bool fin = S == 1;
            float sar[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int64_t row = (int64_t)mt * BM + q * 32 + 16 * (k >> 1) + 8 * (k & 1) + lr;
                sar[k] = row < prm.M ? __ldg(prm.sa + row) : 0.0f;
            }
            for (int c0 = p0; c0 < p1; c0 += prm.hand) {
                mbar_wait(tfull_bar(as), aphase);
                tc_fence_after();
                X2_STAMP(tsn++);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (j == 1 && !two) break;
                    const uint32_t col = (uint32_t)(as * ACC_COLS + (cg + CG * j) * 32);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {   // one 16-row half at a time: 64 accumulators + 16 in flight fit the registers
                        float v[16];
                        tmem_ld16_sync(tmem_base + ((uint32_t)(q * 32 + 16 * h) << 16) + col, v);
#pragma unroll
                        for (int i = 0; i < 16; ++i) acc[j][h][i] += v[i];
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar(as));
                if (++as == ACC_STAGES) { as = 0; aphase ^= 1u; }
            }
            // finish: undo the operand scales, bias, ReLU, store.  element (k = 2h + g, i, e): acc[j][h][4i + 2g + e]  <->
            // row 16h + 8g + lr of the warp's 32, column 8i + lc + e of chunk j.  Each 16-row half of a chunk is scaled by its
            // rows' factors, goes through the warp's 2 KB staging box (16 rows x 128 B, 16-byte chunk c of row r at position
            // c ^ (r % 8): conflict-free both ways) and is read back with a lane owning 4 consecutive columns of one row: the
            // column factors and the bias are one float4 each per lane and chunk, the stores full 128-byte row segments.
            const int64_t col0 = (int64_t)nt * prm.bn, rowq = (int64_t)mt * BM + q * 32;
            float* __restrict__ Dp = prm.D + (int64_t)sp * prm.M * prm.N;
            // (uniform) nothing to clip and every float4 access 16-byte aligned
            const bool fast = (prm.N & 3) == 0 && (int64_t)mt * BM + BM <= prm.M && col0 + prm.bn <= prm.N &&
                              ((reinterpret_cast<uintptr_t>(prm.D) | reinterpret_cast<uintptr_t>(prm.sb) | reinterpret_cast<uintptr_t>(prm.bias)) & 15) == 0;
            uint8_t* stg_ptr = smem_raw + (stg_base - raw) + (uint32_t)(warp - 2) * STG_WARP;
            const int c4 = (lane & 7) * 4;     // read-back: this lane's first column inside the chunk
            X2_STAMP(tsn++);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (j == 1 && !two) break;
                const int64_t n0 = col0 + (cg + CG * j) * 32;
                float4 s4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f), b4 = s4;
                if (fast) {
                    s4 = __ldg(reinterpret_cast<const float4*>(prm.sb + n0 + c4));
                    if (fin && prm.bias) b4 = __ldg(reinterpret_cast<const float4*>(prm.bias + n0 + c4));
                } else {
                    float* sv = &s4.x;
                    float* bv = &b4.x;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (n0 + c4 + e < prm.N) {
                            sv[e] = __ldg(prm.sb + n0 + c4 + e);
                            if (fin && prm.bias) bv[e] = __ldg(prm.bias + n0 + c4 + e);
                        }
                    }
                }
                X2_STAMP(tsn++);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int g = 0; g < 2; ++g) {
                            const float rsc = sar[2 * h + g];
                            // (plain shared-memory accesses, not asm volatile: the compiler schedules them freely between the
                            //  __syncwarp barriers)
                            *reinterpret_cast<float2*>(stg_ptr + (8 * g + lr) * 128 + (((2 * i + (lc >> 2)) ^ lr) << 4) + (lc & 3) * 4) =
                                make_float2(acc[j][h][4 * i + 2 * g] * rsc, acc[j][h][4 * i + 2 * g + 1] * rsc);
                        }
                    __syncwarp();
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const int rr = 4 * kk + (lane >> 3), c = lane & 7;
                        float4 o = *reinterpret_cast<const float4*>(stg_ptr + rr * 128 + ((c ^ (rr & 7)) << 4));
                        o.x = fmaf(o.x, s4.x, b4.x); o.y = fmaf(o.y, s4.y, b4.y); o.z = fmaf(o.z, s4.z, b4.z); o.w = fmaf(o.w, s4.w, b4.w);
                        if (fin && prm.relu) { o.x = fmaxf(o.x, 0.0f); o.y = fmaxf(o.y, 0.0f); o.z = fmaxf(o.z, 0.0f); o.w = fmaxf(o.w, 0.0f); }
                        const int64_t row = rowq + 16 * h + rr, col = n0 + c4;
#ifdef HVAE_EXPERIMENT
                        if (prm.dbg & 1) { if (o.x == 123.456f) Dp[0] = o.y; continue; }
#endif
                        float* dst = Dp + row * prm.N + col;
                        if (fast) {
                            *reinterpret_cast<float4*>(dst) = o;
                        } else if (row < prm.M) {
                            if (col < prm.N) dst[0] = o.x;
                            if (col + 1 < prm.N) dst[1] = o.y;
                            if (col + 2 < prm.N) dst[2] = o.z;
                            if (col + 3 < prm.N) dst[3] = o.w;
                        }
                    }
                    __syncwarp();   // the box is rewritten by the next half
                }
            }
            X2_STAMP(tsn++);
        }
    }
    tc_fence_before();
    __syncthreads();
#ifdef HVAE_EXPERIMENT
    if (prm.ts && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        prm.ts[192 + 3 * blockIdx.x + 1] = (long long)t;
        if (blockIdx.x == 0) prm.ts[189] = clock64();
    }
#endif
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(ACC_STAGES * ACC_COLS));
    }
}

// ---- operand preparation ------------------------------------------------------------------------------------------------
// power-of-two scale that brings a row whose largest magnitude has the float bits `maxbits` into [2^14, 2^15); the
// exponent is clamped so that both the scale and its inverse are normal floats (an all-zero row: scale 1)
__device__ __forceinline__ float scale_of(uint32_t maxbits, float& inv) {
    int e = (int)((maxbits >> 23) & 0xffu) - 127;     // floor(log2(max)); denormal / zero -> -127
    if (maxbits == 0u) { inv = 1.0f; return 1.0f; }
    e = e < -100 ? -100 : e;
    inv = __uint_as_float((uint32_t)(127 + (e - 14)) << 23);    // 2^(e - 14)
    return __uint_as_float((uint32_t)(127 - (e - 14)) << 23);   // 2^(14 - e)
}
__device__ __forceinline__ void split2h(float v, __half& h, __half& l) {
    h = __float2half_rn(v);
    l = __float2half_rn(v - __half2float(h));   // (exact subtraction)
}

// (R, C) fp32 -> (R, 2*Cp) fp16 [hi | lo] of row r times 2^e_r, inv[r] = 2^-e_r; zero padding up to Cp.  One warp per row,
// two passes over the row (the second hits L1).
__global__ void __launch_bounds__(256) k_split2h_rows(const float* __restrict__ in, __half* __restrict__ out, float* __restrict__ inv,
                                                      int64_t R, int64_t C, int64_t Cp) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < R; r += nw) {
        const float* row = in + r * C;
        uint32_t mb = 0;
        for (int64_t c = lane; c < C; c += 32) mb = max(mb, __float_as_uint(row[c]) & 0x7fffffffu);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mb = max(mb, __shfl_xor_sync(0xffffffffu, mb, o));
        float iv;
        const float s = scale_of(mb, iv);
        if (lane == 0) inv[r] = iv;
        __half* o = out + r * 2 * Cp;
        for (int64_t c = 2 * lane; c < Cp; c += 64) {
            const float a = c < C ? row[c] * s : 0.0f, b = c + 1 < C ? row[c + 1] * s : 0.0f;
            __half2 h, l;
            split2h(a, h.x, l.x);
            split2h(b, h.y, l.y);
            *reinterpret_cast<__half2*>(o + c) = h;
            *reinterpret_cast<__half2*>(o + Cp + c) = l;
        }
    }
}

// largest magnitude (as float bits) of every row and every column of an (R, C) matrix: a block owns a 32-row x 128-column
// tile (8 warps x 4 rows each, a lane reads 4 consecutive columns), row maxima by warp shuffle, column maxima through
// shared memory, then one atomicMax per row / column and block (rowbits and colbits zeroed by the caller)
// mask (optional, same shape): elements whose mask value is not > 0 count as zero (the ReLU backward of a fused
// Linear + ReLU layer).  colpart (optional): the block's column sums of its 32 rows -> colpart[blockIdx.x][C] (the bias
// gradient = their sum over the row blocks, hvae_common.cuh k_reduce_slabs; fixed order, no atomics).
constexpr int kAmRows = 32, kAmCols = 128;
__global__ void __launch_bounds__(256) k_absmax_rc(const float* __restrict__ in, const float* __restrict__ mask,
                                                   uint32_t* __restrict__ rowbits, uint32_t* __restrict__ colbits,
                                                   float* __restrict__ colpart, int64_t R, int64_t C) {
    __shared__ uint32_t cmax[8][kAmCols];
    float csum[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t r0 = (int64_t)blockIdx.x * kAmRows, c0 = (int64_t)blockIdx.y * kAmCols + 4 * lane;
    const bool vec = (C & 3) == 0;
    uint32_t cm[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int i = 0; i < kAmRows / 8; ++i) {
        const int64_t r = r0 + warp + 8 * i;
        uint32_t b[4] = {0u, 0u, 0u, 0u};
        if (r < R) {
            if (vec && c0 + 4 <= C) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(in + r * C + c0));
                b[0] = __float_as_uint(v.x); b[1] = __float_as_uint(v.y); b[2] = __float_as_uint(v.z); b[3] = __float_as_uint(v.w);
                if (mask) {
                    const float4 m = __ldg(reinterpret_cast<const float4*>(mask + r * C + c0));
                    if (!(m.x > 0.0f)) b[0] = 0u;
                    if (!(m.y > 0.0f)) b[1] = 0u;
                    if (!(m.z > 0.0f)) b[2] = 0u;
                    if (!(m.w > 0.0f)) b[3] = 0u;
                }
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (c0 + e < C) {
                        b[e] = __float_as_uint(__ldg(in + r * C + c0 + e));
                        if (mask && !(__ldg(mask + r * C + c0 + e) > 0.0f)) b[e] = 0u;
                    }
            }
        }
        uint32_t rm = 0u;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            csum[e] += __uint_as_float(b[e]);
            b[e] &= 0x7fffffffu;
            cm[e] = max(cm[e], b[e]);
            rm = max(rm, b[e]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rm = max(rm, __shfl_xor_sync(0xffffffffu, rm, o));
        if (rowbits && lane == 0 && r < R && rm) atomicMax(rowbits + r, rm);
    }
    if (!colbits && !colpart) return;
    if (colbits) {
#pragma unroll
        for (int e = 0; e < 4; ++e) cmax[warp][4 * lane + e] = cm[e];
        __syncthreads();
        if (threadIdx.x < kAmCols) {
            uint32_t m = 0u;
#pragma unroll
            for (int w = 0; w < 8; ++w) m = max(m, cmax[w][threadIdx.x]);
            const int64_t c = (int64_t)blockIdx.y * kAmCols + threadIdx.x;
            if (c < C && m) atomicMax(colbits + c, m);
        }
    }
    if (colpart) {
        float* fsum = reinterpret_cast<float*>(&cmax[0][0]);   // reuse the staging array
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 4; ++e) fsum[warp * kAmCols + 4 * lane + e] = csum[e];
        __syncthreads();
        if (threadIdx.x < kAmCols) {
            float t = 0.0f;
#pragma unroll
            for (int w = 0; w < 8; ++w) t += fsum[w * kAmCols + threadIdx.x];
            const int64_t c = (int64_t)blockIdx.y * kAmCols + threadIdx.x;
            if (c < C) colpart[(int64_t)blockIdx.x * C + c] = t;
        }
    }
}

// Both layouts from one read: (R, C) fp32 -> rows split (R, 2*Cp) scaled per row AND transposed split (C, 2*Rp) scaled per
// column (either may be NULL); the inverse scales go to inv_r (R,) / inv_c (C,).  64 x 64 tiles through shared memory.
__global__ void __launch_bounds__(256) k_split2h_both(const float* __restrict__ in, const float* __restrict__ mask,
                                                      const uint32_t* __restrict__ rowbits,
                                                      const uint32_t* __restrict__ colbits, __half* __restrict__ out_r,
                                                      float* __restrict__ inv_r, __half* __restrict__ out_t, float* __restrict__ inv_c,
                                                      int64_t R, int64_t C, int64_t Cp, int64_t Rp) {
    __shared__ float tile[64][65];
    __shared__ float srow[64], scol[64];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t r0 = (int64_t)blockIdx.x * 64, c0 = (int64_t)blockIdx.y * 64;
    if (threadIdx.x < 64) {
        const int64_t r = r0 + threadIdx.x;
        float iv = 1.0f, s = 1.0f;
        if (out_r && r < R) {
            s = scale_of(rowbits[r], iv);
            if (blockIdx.y == 0) inv_r[r] = iv;
        }
        srow[threadIdx.x] = s;
    } else if (threadIdx.x < 128) {
        const int64_t c = c0 + threadIdx.x - 64;
        float iv = 1.0f, s = 1.0f;
        if (out_t && c < C) {
            s = scale_of(colbits[c], iv);
            if (blockIdx.x == 0) inv_c[c] = iv;
        }
        scol[threadIdx.x - 64] = s;
    }
    __syncthreads();
    const bool c_even = (C & 1) == 0;
    for (int i = ty; i < 64; i += 8) {
        const int64_t r = r0 + i, c = c0 + 2 * tx;
        float v0 = 0.0f, v1 = 0.0f;
        if (r < R) {
            if (c_even && c + 1 < C) {
                const float2 v = __ldg(reinterpret_cast<const float2*>(in + r * C + c));
                v0 = v.x; v1 = v.y;
            } else {
                if (c < C) v0 = __ldg(in + r * C + c);
                if (c + 1 < C) v1 = __ldg(in + r * C + c + 1);
            }
            if (mask) {
                if (c < C && !(__ldg(mask + r * C + c) > 0.0f)) v0 = 0.0f;
                if (c + 1 < C && !(__ldg(mask + r * C + c + 1) > 0.0f)) v1 = 0.0f;
            }
        }
        tile[i][2 * tx] = v0;
        tile[i][2 * tx + 1] = v1;
        if (out_r && r < R) {
            const float s = srow[i];
            __half2 h, l;
            split2h(v0 * s, h.x, l.x);
            split2h(v1 * s, h.y, l.y);
            __half* o = out_r + r * 2 * Cp + c;
            *reinterpret_cast<__half2*>(o) = h;
            *reinterpret_cast<__half2*>(o + Cp) = l;
        }
    }
    if (!out_t) return;
    __syncthreads();
    for (int i = ty; i < 64; i += 8) {
        const int64_t c = c0 + i, r = r0 + 2 * tx;
        if (c < C) {
            const float s = scol[i];
            __half2 h, l;
            split2h(tile[2 * tx][i] * s, h.x, l.x);
            split2h(tile[2 * tx + 1][i] * s, h.y, l.y);
            __half* o = out_t + c * 2 * Rp + r;
            *reinterpret_cast<__half2*>(o) = h;
            *reinterpret_cast<__half2*>(o + Rp) = l;
        }
    }
}

// split-K fix-up: C = sum_s part[s] (+ bias per column) (ReLU)
__global__ void __launch_bounds__(256) k_x2_reduce(const float* __restrict__ part, const float* __restrict__ bias, float* __restrict__ C,
                                                   int64_t M, int64_t N, int S, int relu) {
    const int64_t n = M * N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float a = 0.0f;
        for (int s = 0; s < S; ++s) a += part[(int64_t)s * n + i];
        if (bias) a += __ldg(bias + i % N);
        if (relu) a = fmaxf(a, 0.0f);
        C[i] = a;
    }
}

// ---- host side ----------------------------------------------------------------------------------------------------------
static int64_t kp_of(int64_t K) { return (K + BK - 1) / BK * BK; }

struct Plan { int bn, splits, nst; };
// tile width and split-K factor with the smallest modelled time: waves x (k-positions x 12 MMAs + fill / drain), plus the
// reduce pass when the contraction is split
static Plan pick_plan(int64_t M, int64_t N, int64_t K) {
    const int64_t kpos = kp_of(K) / BK, m_tiles = (M + BM - 1) / BM;
    Plan best{128, 1, 2};
    double best_cost = 1e30;
    for (int bn = 128; bn <= 256; bn += 32) {
        if (bn > 128 && bn - 32 >= N) break;   // wider than the matrix
        const int64_t n_tiles = (N + bn - 1) / bn;
        const double smem_clk = (double)(BM + bn) * UMMA_K * 2 / 128.0, mma_clk = bn / 2.0 > smem_clk ? bn / 2.0 : smem_clk;
        for (int S = 1; S <= MAX_SPLITS && S <= kpos; ++S) {
            const int64_t units = m_tiles * n_tiles * S;
            const int64_t waves = (units + kNumSMs - 1) / kNumSMs;
            const double kper = (double)((kpos + S - 1) / S);
            double cost = (double)waves * (kper * 12.0 * mma_clk + 4000.0);
            if (S > 1) cost += 6000.0 + (double)M * N * 4.0 * (S + 1) / 2500.0;   // extra launch + partial planes through L2
            if (cost < best_cost) { best_cost = cost; best.bn = bn; best.splits = S; }
        }
    }
#ifdef HVAE_EXPERIMENT
    if (getenv("HVAE_X2_BN")) best.bn = atoi(getenv("HVAE_X2_BN"));
    if (getenv("HVAE_X2_SPLITS")) best.splits = atoi(getenv("HVAE_X2_SPLITS"));
#endif
    const uint32_t stage = 2u * TILE_A + 2u * (uint32_t)best.bn * (BK * 2);
    int nst = (int)((SMEM_LIMIT - 1024u - BAR_BYTES - STG_BYTES) / stage);
    best.nst = nst > MAX_STAGES ? MAX_STAGES : nst;
#ifdef HVAE_EXPERIMENT
    if (getenv("HVAE_X2_NST") && atoi(getenv("HVAE_X2_NST")) < best.nst) best.nst = atoi(getenv("HVAE_X2_NST"));
#endif
    return best;
}

}  // namespace x2
}  // namespace hvae

using namespace hvae;

#ifdef HVAE_EXPERIMENT
static long long* g_x2_ts = nullptr;
// experiment build: a device buffer of 192 int64 that CTA 0 of every following k_x2_gemm launch stamps with clock64
extern "C" void hvae_exp_x2_timestamps(long long* dev_buf) { g_x2_ts = dev_buf; }
#endif

extern "C" size_t hvae_split2h_bytes(int64_t rows, int64_t cols) {
    if (rows <= 0 || cols <= 0) return 0;
    return (size_t)rows * 2 * (size_t)x2::kp_of(cols) * 2;
}
extern "C" size_t hvae_split2h_workspace_bytes(int64_t rows, int64_t cols) {
    if (rows <= 0 || cols <= 0) return 0;
    return (size_t)(rows + cols) * 4 + 256;
}

// src (rows, cols) fp32 -> dst (rows, 2*Cp) fp16 [hi | lo] of row r times 2^e_r (Cp = cols rounded up to 64) and
// inv_scale (rows,) = 2^-e_r.
extern "C" int hvae_split2h_rows_f32(const float* src, void* dst, float* inv_scale, int64_t rows, int64_t cols, void* stream) {
    if (rows <= 0 || cols <= 0) return HVAE_ESHAPE;
    if (!src || !dst || !inv_scale) return HVAE_EARG;
    const int64_t Cp = x2::kp_of(cols);
    const int64_t blocks = (rows + 7) / 8;
    const unsigned grid = (unsigned)(blocks < (int64_t)kNumSMs * 16 ? blocks : (int64_t)kNumSMs * 16);
    x2::k_split2h_rows<<<grid, 256, 0, (cudaStream_t)stream>>>(src, (__half*)dst, inv_scale, rows, cols, Cp);
    return check_launch();
}

// Both layouts of src (rows, cols): dst_rows (rows, 2*Cp) scaled per row with inv_rows (rows,), and dst_t (cols, 2*Rp)
// - the split of src^T - scaled per COLUMN of src with inv_cols (cols,).  Either layout may be NULL.  workspace:
// hvae_split2h_workspace_bytes.  (A maximum pass, then the split pass.)
extern "C" size_t hvae_split2h_ex_workspace_bytes(int64_t rows, int64_t cols) {
    if (rows <= 0 || cols <= 0) return 0;
    const size_t row_blocks = (size_t)((rows + x2::kAmRows - 1) / x2::kAmRows);
    return ((size_t)(rows + cols) * 4 + 255) / 256 * 256 + row_blocks * (size_t)cols * 4 + 256;
}

// The general form: src optionally masked (elements where mask <= 0 count as zero: the ReLU backward folded into the
// operand split of the gradient), either layout optional, and optionally the column sums of the (masked) src (the bias
// gradient, out of the maximum pass's read).  workspace: hvae_split2h_ex_workspace_bytes.
extern "C" int hvae_split2h_both_ex_f32(const float* src, const float* mask, void* dst_rows, float* inv_rows, void* dst_t,
                                        float* inv_cols, float* colsum, int64_t rows, int64_t cols, void* workspace,
                                        size_t workspace_bytes, void* stream) {
    if (rows <= 0 || cols <= 0) return HVAE_ESHAPE;
    if (!src || (!dst_rows && !dst_t && !colsum) || (dst_rows && !inv_rows) || (dst_t && !inv_cols)) return HVAE_EARG;
    if (!dst_t && !mask && !colsum) return hvae_split2h_rows_f32(src, dst_rows, inv_rows, rows, cols, stream);
    if (!workspace || workspace_bytes < hvae_split2h_ex_workspace_bytes(rows, cols)) return HVAE_EARG;
    const int64_t Cp = x2::kp_of(cols), Rp = x2::kp_of(rows);
    if (Cp / 64 > 65535) return HVAE_ESHAPE;
    cudaStream_t s = (cudaStream_t)stream;
    uint32_t* colbits = (uint32_t*)workspace;
    uint32_t* rowbits = colbits + cols;
    float* colpart = (float*)((char*)workspace + ((size_t)(rows + cols) * 4 + 255) / 256 * 256);
    if (cudaMemsetAsync(colbits, 0, (size_t)(cols + rows) * 4, s) != cudaSuccess) return HVAE_ELAUNCH;
    const int64_t row_blocks = (rows + x2::kAmRows - 1) / x2::kAmRows;
    {
        dim3 grid((unsigned)row_blocks, (unsigned)((cols + x2::kAmCols - 1) / x2::kAmCols));
        x2::k_absmax_rc<<<grid, 256, 0, s>>>(src, mask, dst_rows ? rowbits : nullptr, dst_t ? colbits : nullptr,
                                             colsum ? colpart : nullptr, rows, cols);
    }
    if (colsum) {
        SlabReducer red;
        red.add(colpart, colsum, cols, (int)row_blocks);
        red.launch(s);
    }
    if (dst_rows || dst_t) {
        dim3 grid((unsigned)(Rp / 64), (unsigned)(Cp / 64));
        x2::k_split2h_both<<<grid, 256, 0, s>>>(src, mask, rowbits, colbits, (__half*)dst_rows, inv_rows, (__half*)dst_t, inv_cols, rows,
                                                cols, Cp, Rp);
    }
    return check_launch();
}

extern "C" int hvae_split2h_both_f32(const float* src, void* dst_rows, float* inv_rows, void* dst_t, float* inv_cols, int64_t rows,
                                     int64_t cols, void* workspace, size_t workspace_bytes, void* stream) {
    if (rows <= 0 || cols <= 0) return HVAE_ESHAPE;
    if (!src || (!dst_rows && !dst_t) || (dst_rows && !inv_rows) || (dst_t && !inv_cols)) return HVAE_EARG;
    if (!dst_t) return hvae_split2h_rows_f32(src, dst_rows, inv_rows, rows, cols, stream);
    if (!workspace || workspace_bytes < hvae_split2h_workspace_bytes(rows, cols)) return HVAE_EARG;
    // the plain form needs only the row / column maxima: the first (rows + cols) words of the workspace
    const int64_t Cp = x2::kp_of(cols), Rp = x2::kp_of(rows);
    if (Cp / 64 > 65535) return HVAE_ESHAPE;
    cudaStream_t s = (cudaStream_t)stream;
    uint32_t* colbits = (uint32_t*)workspace;
    uint32_t* rowbits = colbits + cols;
    if (cudaMemsetAsync(colbits, 0, (size_t)(cols + rows) * 4, s) != cudaSuccess) return HVAE_ELAUNCH;
    {
        dim3 grid((unsigned)((rows + x2::kAmRows - 1) / x2::kAmRows), (unsigned)((cols + x2::kAmCols - 1) / x2::kAmCols));
        x2::k_absmax_rc<<<grid, 256, 0, s>>>(src, nullptr, dst_rows ? rowbits : nullptr, colbits, nullptr, rows, cols);
    }
    dim3 grid((unsigned)(Rp / 64), (unsigned)(Cp / 64));
    x2::k_split2h_both<<<grid, 256, 0, s>>>(src, nullptr, rowbits, colbits, (__half*)dst_rows, inv_rows, (__half*)dst_t, inv_cols, rows, cols,
                                            Cp, Rp);
    return check_launch();
}

extern "C" size_t hvae_gemm_x2s_workspace_bytes(int64_t M, int64_t N) {
    if (M <= 0 || N <= 0) return 0;
    return (size_t)x2::MAX_SPLITS * M * N * 4 + 256;
}
extern "C" int hvae_gemm_x2s_num_launches(int64_t M, int64_t N, int64_t K) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    return 1 + (x2::pick_plan(M, N, K).splits > 1 ? 1 : 0);
}
// the tile width / split-K factor / ring depth chosen for a problem (bench.py reports them)
extern "C" int hvae_gemm_x2s_plan(int64_t M, int64_t N, int64_t K, int* bn, int* splits, int* stages) {
    if (M <= 0 || N <= 0 || K <= 0) return HVAE_ESHAPE;
    const x2::Plan p = x2::pick_plan(M, N, K);
    if (bn) *bn = p.bn;
    if (splits) *splits = p.splits;
    if (stages) *stages = p.nst;
    return HVAE_OK;
}

// C (M, N) = A (M, K) . B (N, K)^T (+ bias[n]) (ReLU): As / Bs are hvae_split2h buffers of A / B (rows = the output index,
// contraction contiguous), inv_a (M,) / inv_b (N,) their inverse row scales.
extern "C" int hvae_gemm_x2s_f32(const void* As, const float* inv_a, const void* Bs, const float* inv_b, const float* bias, int relu,
                                 float* C, int64_t M, int64_t N, int64_t K, void* workspace, size_t workspace_bytes, void* stream) {
    if (M <= 0 || N <= 0 || K <= 0) return HVAE_ESHAPE;
    if (!As || !Bs || !inv_a || !inv_b || !C) return HVAE_EARG;
    const int64_t Kp = x2::kp_of(K);
    if (2 * Kp > 0x7fffffffLL) return HVAE_ESHAPE;
    const x2::Plan pl = x2::pick_plan(M, N, K);
    if (pl.splits > 1 && (!workspace || workspace_bytes < (size_t)pl.splits * M * N * 4)) return HVAE_EARG;
    cudaStream_t s = (cudaStream_t)stream;
    CUtensorMap ma, mb;
    if (!tc::make_map(&ma, As, M, 2 * Kp, x2::BM) || !tc::make_map(&mb, Bs, N, 2 * Kp, pl.bn)) return HVAE_ELAUNCH;
    x2::Params prm{};
    prm.M = M; prm.N = N; prm.kpos = (int)(Kp / x2::BK); prm.bn = pl.bn; prm.splits = pl.splits; prm.nst = pl.nst; prm.hand = 2;
#ifdef HVAE_EXPERIMENT
    if (getenv("HVAE_X2_HAND")) prm.hand = atoi(getenv("HVAE_X2_HAND"));
    prm.ts = g_x2_ts;
    prm.dbg = getenv("HVAE_X2_DBG") ? atoi(getenv("HVAE_X2_DBG")) : 0;
#endif
    prm.a_lo = (int)Kp; prm.b_lo = (int)Kp; prm.sa = inv_a; prm.sb = inv_b;
    const uint32_t smem = (uint32_t)pl.nst * (2u * x2::TILE_A + 2u * (uint32_t)pl.bn * (x2::BK * 2)) + x2::STG_BYTES + 1024u + x2::BAR_BYTES;
    cudaFuncSetAttribute(x2::k_x2_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)x2::SMEM_LIMIT);
    const int64_t units = ((M + x2::BM - 1) / x2::BM) * ((N + pl.bn - 1) / pl.bn) * pl.splits;
    if (units > 0x7fffffffLL || N > 0x3ffffffLL) return HVAE_ESHAPE;   // 32-bit unit / in-tile offset arithmetic in the kernel
    const unsigned grid = (unsigned)(units < kNumSMs ? units : kNumSMs);
    if (pl.splits == 1) {
        prm.D = C; prm.bias = bias; prm.relu = relu;
        x2::k_x2_gemm<<<grid, x2::THREADS, smem, s>>>(ma, mb, prm);
        return check_launch();
    }
    prm.D = (float*)workspace;
    x2::k_x2_gemm<<<grid, x2::THREADS, smem, s>>>(ma, mb, prm);
    int rc = check_launch();
    if (rc != HVAE_OK) return rc;
    const int64_t n = M * N;
    const unsigned rgrid = (unsigned)((n + 255) / 256 < (int64_t)kNumSMs * 16 ? (n + 255) / 256 : (int64_t)kNumSMs * 16);
    x2::k_x2_reduce<<<rgrid, 256, 0, s>>>((const float*)workspace, bias, C, M, N, pl.splits, relu);
    return check_launch();
}
