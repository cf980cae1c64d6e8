// K2: gyroplane (signed hyperplane distance) layer, forward and analytic backward — SIMT path for
// latent dims D <= 64 (the decoder-side shapes of configs 1-4: D = 2..64, P = 16..1024).
//
// reference: hyperbolic_vae/layers.py:193-210 (Distance2PoincareHyperplanes) and geoopt's
// Distance2StereographicHyperplanes -> math.dist2plane (App. A.1);  layers.py:96-121 (GeodesicLayer)
// -> manifolds.py:41-65 (normdist2plane) with HVAE_GYRO_PVAE.
//
// The reference broadcasts to (B, D, P) intermediates (~12 of them kept for autograd).  Here the
// Mobius subtraction is re-expressed through inner products:
//   A = 1 - 2c<p,x> + c|x|^2,  Bc = 1 - c|p|^2,  den = 1 - 2c<p,x> + c^2|p|^2|x|^2
//   <diff,a> = (-A<p,a> + Bc<x,a>)/den,   |diff|^2 = (A^2|p|^2 - 2 A Bc <p,x> + Bc^2|x|^2)/den^2
// so one pass over D per (row, plane) pair gives <p,x> (and <a,x>), and the rest is a scalar epilogue.
// Memory is O(B*P) for the output only.  Algorithmic bytes: fwd 4(BD + 2PD + BP), bwd 4(BP + 2BD + 4PD).
// At D <= 64 the kernels are bound by the pair function's instructions and by latency, not by bytes or FMAs:
//  * the pair function has two lean closed forms (gyro_pair.cuh: unprojected - one division; projected - the common
//    case for latent points beyond the fp32 projection radius - one rsqrt) and falls back to the general clamp-by-clamp
//    form only where a clamp binds;
//  * the flag word is a template constant for the two combinations the layers use;
//  * grids are sized in whole waves of the resident-CTA slots (gyro_fwd_rows_per_cta, gyro_x_plan);
//  * HVAE_GYRO_RELU fuses the decoder's ReLU (forward clamp, backward mask from the recomputed sign).
#include "hvae_common.cuh"
#include "gyro_pair.cuh"

namespace hvae {

// ---------------------------------------------------------------------------------------------------
// tiling: a CTA of 128 threads owns TB rows x TJ planes; thread = one plane, 8 rows at a time.
// smem: xs[TB][D4] (row-major, broadcast reads), ps[D4][TJ] / as[D4][TJ] (plane-minor, conflict-free).
// ---------------------------------------------------------------------------------------------------
constexpr int kGyroThreads = 128;
constexpr int kGyroRB = 8;    // rows per register block

// FL >= 0: the flag word as a compile-time constant (the two combinations the layers use: GeodesicLayer's
// PVAE|SIGNED with a != p, Distance2PoincareHyperplanes' SIGNED with a == p); FL < 0: flags read from prm at run time.
template <int D4, bool kAliased, int FL>
__global__ void __launch_bounds__(kGyroThreads)
k_gyro_fwd(const float* __restrict__ x, const float* __restrict__ p, const float* __restrict__ a,
           const float* __restrict__ bias, float* __restrict__ out, int B, int D, int P_, int kGyroTB /* rows per CTA */,
           GyroParams prm) {
    if (FL >= 0) prm.flags = (uint32_t)FL;
    constexpr int TJ = kGyroThreads;
    extern __shared__ float smem[];
    float* xs = smem;                       // [TB][D4]
    float* xs2 = xs + kGyroTB * D4;         // [TB]
    float* ps = xs2 + kGyroTB;              // [D4][TJ]
    float* as = ps + D4 * TJ;               // [D4][TJ] (unused when aliased)
    const int tid = threadIdx.x;
    const int j0 = blockIdx.x * TJ;
    const int b0 = blockIdx.y * kGyroTB;
    const int j = j0 + tid;
    // stage planes (plane-minor) and rows
    for (int i = tid; i < TJ * D4; i += kGyroThreads) {
        const int jj = i / D4, d = i - jj * D4;
        const bool ok = (j0 + jj) < P_ && d < D;
        ps[d * TJ + jj] = ok ? __ldg(p + (int64_t)(j0 + jj) * D + d) : 0.0f;
        if (!kAliased) as[d * TJ + jj] = ok ? __ldg(a + (int64_t)(j0 + jj) * D + d) : 0.0f;
    }
    for (int i = tid; i < kGyroTB * D4; i += kGyroThreads) {
        const int bb = i / D4, d = i - bb * D4;
        xs[i] = ((b0 + bb) < B && d < D) ? __ldg(x + (int64_t)(b0 + bb) * D + d) : 0.0f;
    }
    __syncthreads();
    if (tid < kGyroTB) {
        float s = 0.0f;
#pragma unroll 4
        for (int d = 0; d < D4; ++d) s = fmaf(xs[tid * D4 + d], xs[tid * D4 + d], s);
        xs2[tid] = s;
    }
    float p2 = 0.0f, pa = 0.0f, a2 = 0.0f;
#pragma unroll 4
    for (int d = 0; d < D4; ++d) {
        const float pv = ps[d * TJ + tid];
        const float av = kAliased ? pv : as[d * TJ + tid];
        p2 = fmaf(pv, pv, p2);
        pa = fmaf(pv, av, pa);
        a2 = fmaf(av, av, a2);
    }
    const float an_raw = sqrtf(a2);
    const GyroPlaneK pl = gyro_plane_consts(p2, pa, an_raw, prm);
    const float bj = (bias != nullptr && j < P_) ? __ldg(bias + j) : 0.0f;
    __syncthreads();
    for (int r0 = 0; r0 < kGyroTB; r0 += kGyroRB) {
        if (b0 + r0 >= B) break;
        float ee[kGyroRB], qq[kGyroRB], qa[kGyroRB];
#pragma unroll
        for (int r = 0; r < kGyroRB; ++r) ee[r] = qq[r] = qa[r] = 0.0f;
#pragma unroll 2
        for (int d = 0; d < D4; d += 4) {
            const float p0 = ps[(d + 0) * TJ + tid], p1 = ps[(d + 1) * TJ + tid];
            const float p2_ = ps[(d + 2) * TJ + tid], p3 = ps[(d + 3) * TJ + tid];
            float a0 = p0, a1 = p1, a2_ = p2_, a3 = p3;
            if (!kAliased) {
                a0 = as[(d + 0) * TJ + tid]; a1 = as[(d + 1) * TJ + tid];
                a2_ = as[(d + 2) * TJ + tid]; a3 = as[(d + 3) * TJ + tid];
            }
#pragma unroll
            for (int r = 0; r < kGyroRB; ++r) {
                const float4 xv = *reinterpret_cast<const float4*>(xs + (r0 + r) * D4 + d);
                const float t0 = p0 - xv.x, t1 = p1 - xv.y, t2 = p2_ - xv.z, t3 = p3 - xv.w;
                ee[r] = fmaf(t0, t0, fmaf(t1, t1, fmaf(t2, t2, fmaf(t3, t3, ee[r]))));
                qq[r] = fmaf(p0, t0, fmaf(p1, t1, fmaf(p2_, t2, fmaf(p3, t3, qq[r]))));
                if (!kAliased) qa[r] = fmaf(a0, t0, fmaf(a1, t1, fmaf(a2_, t2, fmaf(a3, t3, qa[r]))));
            }
        }
        if (j < P_) {
            float* op = out + (int64_t)(b0 + r0) * P_ + j;
            const int nr = min(kGyroRB, B - b0 - r0);
#pragma unroll
            for (int r = 0; r < kGyroRB; ++r) {
                if (r < nr) {
                    GyroDiff df;
                    df.e = ee[r]; df.q = qq[r]; df.qa = kAliased ? qq[r] : qa[r];
                    GyroLean L;
                    float o;
                    if (gyro_pair_lean_fwd(df, xs2[r0 + r], pl, prm, L, o) == GYRO_GENERAL)
                        o = gyro_pair_fwd_general(df.e, df.q, df.qa, xs2[r0 + r], p2, pa, an_raw, prm);
                    o += bj;
                    if (prm.flags & HVAE_GYRO_RELU) o = fmaxf(o, 0.0f);
                    *op = o;
                    op += P_;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// backward.  The pair math (forward recompute + scalar backward, ~400 instructions) runs ONCE per (row, plane):
//   G1  k_gyro_bwd_pairs : thread = one row of a 128-row block, planes of one chunk streamed through smem;
//                          accumulates gx in registers, writes the coefficient matrices
//                            CP[b][j] = dL/d<p_j,x_b> (+ dL/d<a_j,x_b> when a aliases p),  CA[b][j] = dL/d<a_j,x_b>
//                          and per-(row-block, plane) sums of the scalar terms (d|p|^2, d<p,a>, d|a|, g).
//   G2  k_gyro_bwd_planes: thread = one plane; gp_j = sum_b CP[b][j] x_b + (scalar terms) — a skinny GEMM over a
//                          slab of rows, partials per slab, then k_reduce_slabs (one launch for gx, gp, ga, gbias; deterministic, no atomics).
// ---------------------------------------------------------------------------------------------------
constexpr int kGyroBxThreads = 128;  // rows per CTA
constexpr int kGyroBxTJ = 32;        // planes per smem stage

// small D: cap registers so 5 CTAs fit an SM (the static smem allows 5): 115 -> 89 registers at D4 = 12, no spills
template <int D4, bool kAliased, int FL>
__global__ void __launch_bounds__(kGyroBxThreads, (D4 <= 16 ? 5 : 1))
k_gyro_bwd_pairs(const float* __restrict__ x, const float* __restrict__ p, const float* __restrict__ a,
                 const float* __restrict__ bias /* forward bias, read only with HVAE_GYRO_RELU */, const float* __restrict__ gout, float* __restrict__ gx, float* __restrict__ CP, float* __restrict__ CA,
                 float* __restrict__ wsum /* [rowblocks][P][4] */, int B, int D, int P_, int planes_per_chunk,
                 GyroParams prm) {
    constexpr int TJ = (D4 >= 64) ? 16 : kGyroBxTJ;  // keep the static smem under 48 KB at D4 = 64
    __shared__ float ps[TJ][D4];
    __shared__ float as[kAliased ? 1 : TJ][D4];
    __shared__ __align__(16) float pst[TJ][8];        // GyroPlaneK of the staged planes
    __shared__ float gs[kGyroBxThreads][TJ + 1];      // upstream grad tile [row][plane]; reused for CP
    __shared__ float cs[kAliased ? 1 : kGyroBxThreads][TJ + 1];  // CA tile
    __shared__ float psum[4][TJ][4];                  // per-warp partial sums of the scalar terms
    if (FL >= 0) prm.flags = (uint32_t)FL;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int b0 = blockIdx.x * kGyroBxThreads;
    const int b = b0 + tid;
    float* const grow = &gs[tid][0];                         // this thread's row of the gradient / CP tile
    float* const crow = &cs[kAliased ? 0 : tid][0];
    float* const prow = &psum[warp][0][(lane >> 3) & 3];    // lane 8q publishes sum q
    float xr[D4], acc[D4];
    float x2 = 0.0f, sdx2 = 0.0f;
#pragma unroll
    for (int d = 0; d < D4; ++d) {
        xr[d] = (b < B && d < D) ? __ldg(x + (int64_t)b * D + d) : 0.0f;
        x2 = fmaf(xr[d], xr[d], x2);
        acc[d] = 0.0f;
    }
    const int jlo = blockIdx.y * planes_per_chunk;
    const int jhi = min(P_, jlo + planes_per_chunk);
    gx += (int64_t)blockIdx.y * B * D;  // this chunk's (B,D) slab (the final gx when there is one chunk)
    for (int j0 = jlo; j0 < jhi; j0 += TJ) {
        __syncthreads();
        for (int i = tid; i < TJ * D4; i += kGyroBxThreads) {
            const int jj = i / D4, d = i - jj * D4;
            const bool ok = (j0 + jj) < jhi && d < D;
            ps[jj][d] = ok ? __ldg(p + (int64_t)(j0 + jj) * D + d) : 0.0f;
            if (!kAliased) as[jj][d] = ok ? __ldg(a + (int64_t)(j0 + jj) * D + d) : 0.0f;
        }
        for (int i = tid; i < kGyroBxThreads * TJ; i += kGyroBxThreads) {
            const int rr = i / TJ, jj = i - rr * TJ;
            gs[rr][jj] = ((b0 + rr) < B && (j0 + jj) < jhi) ? __ldg(gout + (int64_t)(b0 + rr) * P_ + j0 + jj) : 0.0f;
        }
        __syncthreads();
        if (tid < TJ) {
            float p2 = 0.0f, pa = 0.0f, a2 = 0.0f;
            for (int d = 0; d < D4; ++d) {
                const float pv = ps[tid][d];
                const float av = kAliased ? pv : as[tid][d];
                p2 = fmaf(pv, pv, p2);
                pa = fmaf(pv, av, pa);
                a2 = fmaf(av, av, a2);
            }
            const GyroPlaneK pk = gyro_plane_consts(p2, pa, sqrtf(a2), prm);
            *reinterpret_cast<float4*>(&pst[tid][0]) = make_float4(pk.p2, pk.pa, pk.an_raw, pk.an);
            const float bj = (bias != nullptr && (j0 + tid) < jhi) ? __ldg(bias + j0 + tid) : 0.0f;
            *reinterpret_cast<float4*>(&pst[tid][4]) = make_float4(pk.ran, pk.Bc, pk.kproj, bj);
        }
        __syncthreads();
        const int jn = min(TJ, jhi - j0);
        for (int jj = 0; jj < jn; ++jj) {
            GyroDiff df;
            df.e = df.q = df.qa = 0.0f;
#pragma unroll
            for (int d = 0; d < D4; d += 4) {
                const float4 pv = *reinterpret_cast<const float4*>(&ps[jj][d]);
                const float t0 = pv.x - xr[d], t1 = pv.y - xr[d + 1], t2 = pv.z - xr[d + 2], t3 = pv.w - xr[d + 3];
                df.e = fmaf(t0, t0, fmaf(t1, t1, fmaf(t2, t2, fmaf(t3, t3, df.e))));
                df.q = fmaf(pv.x, t0, fmaf(pv.y, t1, fmaf(pv.z, t2, fmaf(pv.w, t3, df.q))));
                if (!kAliased) {
                    const float4 av = *reinterpret_cast<const float4*>(&as[jj][d]);
                    df.qa = fmaf(av.x, t0, fmaf(av.y, t1, fmaf(av.z, t2, fmaf(av.w, t3, df.qa))));
                }
            }
            if (kAliased) df.qa = df.q;
            GyroPlaneK pl;
            {
                const float4 c0 = *reinterpret_cast<const float4*>(&pst[jj][0]);
                const float4 c1 = *reinterpret_cast<const float4*>(&pst[jj][4]);
                pl.p2 = c0.x; pl.pa = c0.y; pl.an_raw = c0.z; pl.an = c0.w; pl.ran = c1.x; pl.Bc = c1.y; pl.kproj = c1.z;
            }
            float g = grow[jj];   // zero for rows past B: every gradient term below is linear in g
            GyroPairGrad gr;
            GyroLean L;
            float o;
            const int mode = gyro_pair_lean_fwd(df, x2, pl, prm, L, o);
            if (mode != GYRO_GENERAL) {
                if ((prm.flags & HVAE_GYRO_RELU) && !(o + pst[jj][7] > 0.0f)) g = 0.0f;   // fused ReLU: mask by the recomputed sign
                gr = gyro_pair_lean_bwd(mode, g, df, x2, pl, prm, L);
            } else {
                gr = gyro_pair_grad_general(g, df.e, df.q, df.qa, x2, pl.p2, pl.pa, pl.an_raw, prm, pst[jj][7]);
                g = gr.g_used;
            }
            sdx2 += gr.dx2;
            const float cp = kAliased ? gr.dpx + gr.dxa : gr.dpx;
#pragma unroll
            for (int d = 0; d < D4; d += 4) {
                const float4 pv = *reinterpret_cast<const float4*>(&ps[jj][d]);
                acc[d] = fmaf(cp, pv.x, acc[d]); acc[d + 1] = fmaf(cp, pv.y, acc[d + 1]);
                acc[d + 2] = fmaf(cp, pv.z, acc[d + 2]); acc[d + 3] = fmaf(cp, pv.w, acc[d + 3]);
                if (!kAliased) {
                    const float4 av = *reinterpret_cast<const float4*>(&as[jj][d]);
                    acc[d] = fmaf(gr.dxa, av.x, acc[d]); acc[d + 1] = fmaf(gr.dxa, av.y, acc[d + 1]);
                    acc[d + 2] = fmaf(gr.dxa, av.z, acc[d + 2]); acc[d + 3] = fmaf(gr.dxa, av.w, acc[d + 3]);
                }
            }
            grow[jj] = cp;                          // own element: no hazard
            if (!kAliased) crow[jj] = gr.dxa;
            // per-plane sums over this warp's 32 rows
            // four sums over the warp by a transposing butterfly (6 shuffles instead of 20): the halves of the warp
            // swap two of the four values, then the quarters one, then three plain steps; lane 8q ends with sum q
            {
                const bool h16 = lane & 16, h8 = lane & 8;
                float k0 = h16 ? gr.dan : gr.dp2, k1 = h16 ? g : gr.dpa;
                const float t0 = h16 ? gr.dp2 : gr.dan, t1 = h16 ? gr.dpa : g;
                k0 += __shfl_xor_sync(0xffffffffu, t0, 16);
                k1 += __shfl_xor_sync(0xffffffffu, t1, 16);
                float k = h8 ? k1 : k0;
                const float t = h8 ? k0 : k1;
                k += __shfl_xor_sync(0xffffffffu, t, 8);
                k += __shfl_xor_sync(0xffffffffu, k, 4);
                k += __shfl_xor_sync(0xffffffffu, k, 2);
                k += __shfl_xor_sync(0xffffffffu, k, 1);
                if ((lane & 7) == 0) prow[jj * 4] = k;
            }
        }
        __syncthreads();
        // coalesced write-out of the coefficient tiles and the per-row-block sums
        for (int i = tid; i < kGyroBxThreads * TJ; i += kGyroBxThreads) {
            const int rr = i / TJ, jj = i - rr * TJ;
            if ((b0 + rr) < B && jj < jn) {
                CP[(int64_t)(b0 + rr) * P_ + j0 + jj] = gs[rr][jj];
                if (!kAliased) CA[(int64_t)(b0 + rr) * P_ + j0 + jj] = cs[rr][jj];
            }
        }
        if (tid < jn * 4) {
            const int jj = tid >> 2, q = tid & 3;
            wsum[((int64_t)blockIdx.x * P_ + j0 + jj) * 4 + q] = psum[0][jj][q] + psum[1][jj][q] + psum[2][jj][q] + psum[3][jj][q];
        }
    }
    if (b < B) {
#pragma unroll
        for (int d = 0; d < D4; ++d)
            if (d < D) gx[(int64_t)b * D + d] = acc[d] + 2.0f * sdx2 * xr[d];
    }
}

constexpr int kGyroBpThreads = 64;
constexpr int kGyroBpTB = 64;  // rows per smem stage
constexpr int kGyroBpSlabGran = 64;  // slab granularity in rows (finer than G1's 128-row blocks: more CTAs in flight)

template <int D4, bool kAliased>
__global__ void __launch_bounds__(kGyroBpThreads)
k_gyro_bwd_planes(const float* __restrict__ x, const float* __restrict__ p, const float* __restrict__ a,
                  const float* __restrict__ CP, const float* __restrict__ CA, const float* __restrict__ wsum,
                  float* __restrict__ wp, float* __restrict__ wa, float* __restrict__ wb, int B, int D, int P_,
                  int rows_per_slab /* multiple of kGyroBpSlabGran */) {
    constexpr int TJ = kGyroBpThreads;
    __shared__ float xs[kGyroBpTB][D4];
    const int tid = threadIdx.x;
    const int j = blockIdx.x * TJ + tid;
    const int slab = blockIdx.y;
    const int bs = slab * rows_per_slab;
    const int be = min(B, bs + rows_per_slab);
    float accp[D4], acca[kAliased ? 1 : D4];
#pragma unroll
    for (int d = 0; d < D4; ++d) {
        accp[d] = 0.0f;
        if (!kAliased) acca[d] = 0.0f;
    }
    for (int b0 = bs; b0 < be; b0 += kGyroBpTB) {
        __syncthreads();
        for (int i = tid; i < kGyroBpTB * D4; i += kGyroBpThreads) {
            const int bb = i / D4, d = i - bb * D4;
            xs[bb][d] = ((b0 + bb) < be && d < D) ? __ldg(x + (int64_t)(b0 + bb) * D + d) : 0.0f;
        }
        __syncthreads();
        const int bn = min(kGyroBpTB, be - b0);
        if (j < P_) {
            // the kernel is latency-bound (a few warps per SM, each load a trip to L2): issue the coefficient loads of
            // 16 rows before the first FMA (rows past bn read as zero: they add nothing)
            constexpr int U = 16;
            static_assert(kGyroBpTB % U == 0, "row stage must be a multiple of the load batch");
            for (int bb = 0; bb < bn; bb += U) {
                float cpv[U], cav[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const bool ok = bb + u < bn;
                    cpv[u] = ok ? __ldg(CP + (int64_t)(b0 + bb + u) * P_ + j) : 0.0f;  // coalesced across the CTA's planes
                    cav[u] = (!kAliased && ok) ? __ldg(CA + (int64_t)(b0 + bb + u) * P_ + j) : 0.0f;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const float cp = cpv[u], ca = cav[u];
#pragma unroll
                    for (int d = 0; d < D4; d += 4) {
                        const float4 xv = *reinterpret_cast<const float4*>(&xs[bb + u][d]);
                        accp[d] = fmaf(cp, xv.x, accp[d]); accp[d + 1] = fmaf(cp, xv.y, accp[d + 1]);
                        accp[d + 2] = fmaf(cp, xv.z, accp[d + 2]); accp[d + 3] = fmaf(cp, xv.w, accp[d + 3]);
                        if (!kAliased) {
                            acca[d] = fmaf(ca, xv.x, acca[d]); acca[d + 1] = fmaf(ca, xv.y, acca[d + 1]);
                            acca[d + 2] = fmaf(ca, xv.z, acca[d + 2]); acca[d + 3] = fmaf(ca, xv.w, acca[d + 3]);
                        }
                    }
                }
            }
        }
    }
    if (j < P_) {
        float sdp2 = 0.0f, sdpa = 0.0f, sdan = 0.0f, sg = 0.0f;
        // the scalar sums live per 128-row block of G1: the slab that contains a block's first row takes it
        for (int rb = (bs + kGyroBxThreads - 1) / kGyroBxThreads; rb * kGyroBxThreads < be; ++rb) {
            const float4 w = *reinterpret_cast<const float4*>(wsum + ((int64_t)rb * P_ + j) * 4);
            sdp2 += w.x; sdpa += w.y; sdan += w.z; sg += w.w;
        }
        float a2 = 0.0f;
        for (int d = 0; d < D; ++d) {
            const float av = __ldg((kAliased ? p : a) + (int64_t)j * D + d);
            a2 = fmaf(av, av, a2);
        }
        const float an_raw = sqrtf(a2);
        const float inv_an = an_raw > 0.0f ? 1.0f / an_raw : 0.0f;
        float* wpj = wp + ((int64_t)slab * P_ + j) * D;
        float* waj = kAliased ? nullptr : wa + ((int64_t)slab * P_ + j) * D;
#pragma unroll
        for (int d = 0; d < D4; ++d) {
            if (d < D) {
                const float pv = __ldg(p + (int64_t)j * D + d);
                if (kAliased) {
                    // a == p: <p,a> = |p|^2 and |a| = |p| all flow into the single parameter
                    wpj[d] = accp[d] + (2.0f * sdp2 + 2.0f * sdpa + sdan * inv_an) * pv;
                } else {
                    const float av = __ldg(a + (int64_t)j * D + d);
                    wpj[d] = accp[d] + 2.0f * sdp2 * pv + sdpa * av;
                    waj[d] = acca[d] + sdpa * pv + sdan * inv_an * av;
                }
            }
        }
        wb[(int64_t)slab * P_ + j] = sg;
    }
}

// G2 slabs of rows (multiples of the 128-row blocks of G1)
inline int gyro_slabs(int64_t B, int64_t P) {
    const int64_t jb = (P + kGyroBpThreads - 1) / kGyroBpThreads;
    int64_t want = (8 * kNumSMs + jb - 1) / jb;
    const int64_t maxs = (B + kGyroBpSlabGran - 1) / kGyroBpSlabGran;
    if (want > maxs) want = maxs;
    if (want < 1) want = 1;
    return (int)want;
}

// G1: planes per chunk (gridDim.y chunks).  The pair math is a long dependent chain, so the grid wants several CTAs per SM
// - but in WHOLE waves of the resident-CTA slots (5 per SM at D <= 16 by launch bounds, 2 assumed beyond): at 4096 x 600
// the old "about 8 CTAs per SM" rule made 800 CTAs of 24 planes on 740 slots, i.e. a second wave 8 % full.
// cost = waves x (planes per chunk + ~3 planes' worth of per-CTA prologue / epilogue); ties go to fewer chunks.
inline void gyro_x_plan(int64_t B, int64_t P, int64_t D, int* ppc_out, int* nch_out) {
    const int64_t rb = (B + kGyroBxThreads - 1) / kGyroBxThreads;
    const int64_t slots = (int64_t)kNumSMs * (D <= 16 ? 5 : 2);
    int64_t best_ppc = P, best_cost = -1;
    const int64_t max_ch = P < 512 ? P : 512;
    for (int64_t nch = 1; nch <= max_ch; ++nch) {
        const int64_t ppc = (P + nch - 1) / nch;
        const int64_t n = (P + ppc - 1) / ppc;
        const int64_t cost = ((rb * n + slots - 1) / slots) * (ppc + 3);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_ppc = ppc; }
    }
    *ppc_out = (int)best_ppc;
    *nch_out = (int)((P + best_ppc - 1) / best_ppc);
}

struct GyroWs {
    size_t cp, ca, wsum, wx, wp, wa, wb, total;  // offsets in floats
};

inline GyroWs gyro_ws_layout(int64_t B, int64_t D, int64_t P) {
    GyroWs w;
    const size_t rbs = (size_t)((B + kGyroBxThreads - 1) / kGyroBxThreads);
    int ppc_, nch_;
    gyro_x_plan(B, P, D, &ppc_, &nch_);
    const size_t slabs = (size_t)gyro_slabs(B, P), chunks = (size_t)nch_;
    size_t o = 0;
    auto take = [&](size_t n) { const size_t at = o; o += (n + 3) / 4 * 4; return at; };
    w.cp = take((size_t)B * P);
    w.ca = take((size_t)B * P);
    w.wsum = take(rbs * P * 4);
    w.wx = take(chunks * B * D);
    w.wp = take(slabs * P * D);
    w.wa = take(slabs * P * D);
    w.wb = take(slabs * P);
    w.total = o;
    return w;
}

inline GyroParams make_gyro_params(float c, uint32_t flags) {
    const Ball b = make_ball(c);
    GyroParams p;
    p.c = b.c; p.sc = b.sc; p.rsc = b.rsc; p.maxnorm = b.maxnorm; p.flags = flags;
    return p;
}

// rows per CTA of the forward kernel: the grid should fill the resident-CTA slots of the 148 SMs in whole waves (a
// 4096 x 600 problem at 16 rows per CTA was 1280 CTAs on ~1036 slots: a second wave a quarter full).  cost = waves x
// (rows + ~4 rows' worth of staging the 128 planes)
inline int gyro_fwd_rows_per_cta(int64_t B, int64_t P, int ctas_per_sm) {
    const int64_t jb = (P + kGyroThreads - 1) / kGyroThreads;
    const int64_t slots = (int64_t)kNumSMs * (ctas_per_sm < 1 ? 1 : ctas_per_sm);
    int best = kGyroRB;
    int64_t best_cost = -1;
    for (int tb = kGyroRB; tb <= 128; tb += kGyroRB) {
        const int64_t ctas = jb * ((B + tb - 1) / tb);
        const int64_t cost = ((ctas + slots - 1) / slots) * (tb + 4);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = tb; }
    }
    return best;
}

template <int D4, bool kAliased>
int gyro_fwd_launch_a(const float* x, const float* p, const float* a, const float* bias, float* out, int64_t B, int64_t D,
                      int64_t P, const GyroParams& prm, cudaStream_t s) {
    constexpr uint32_t kLayerFlags = kAliased ? HVAE_GYRO_SIGNED : (HVAE_GYRO_PVAE | HVAE_GYRO_SIGNED);
    auto kern = (prm.flags == kLayerFlags) ? k_gyro_fwd<D4, kAliased, (int)kLayerFlags>
              : (!kAliased && prm.flags == (kLayerFlags | HVAE_GYRO_RELU)) ? k_gyro_fwd<D4, kAliased, (int)(kLayerFlags | HVAE_GYRO_RELU)>
                                                                           : k_gyro_fwd<D4, kAliased, -1>;
    auto smem_for = [](int tb) { return sizeof(float) * ((size_t)tb * D4 + tb + (size_t)D4 * kGyroThreads * (kAliased ? 1 : 2)); };
    if (smem_for(128) > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(128));
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kGyroThreads, smem_for(32)) != cudaSuccess || occ < 1) occ = 4;
    const int tb = gyro_fwd_rows_per_cta(B, P, occ);
    dim3 grid((unsigned)((P + kGyroThreads - 1) / kGyroThreads), (unsigned)((B + tb - 1) / tb));
    kern<<<grid, kGyroThreads, smem_for(tb), s>>>(x, p, a, bias, out, (int)B, (int)D, (int)P, tb, prm);
    return check_launch();
}

template <int D4>
int gyro_fwd_launch(const float* x, const float* p, const float* a, const float* bias, float* out, int64_t B, int64_t D,
                    int64_t P, const GyroParams& prm, cudaStream_t s) {
    if (a == p) return gyro_fwd_launch_a<D4, true>(x, p, a, bias, out, B, D, P, prm, s);
    return gyro_fwd_launch_a<D4, false>(x, p, a, bias, out, B, D, P, prm, s);
}

template <int D4>
int gyro_bwd_launch(const float* x, const float* p, const float* a, const float* bias, const float* gout, float* gx, float* gp, float* ga,
                    float* gbias, int64_t B, int64_t D, int64_t P, const GyroParams& prm, float* ws, cudaStream_t s) {
    const bool aliased = (a == p);
    const GyroWs L = gyro_ws_layout(B, D, P);
    float *CP = ws + L.cp, *CA = ws + L.ca, *wsum = ws + L.wsum, *wx = ws + L.wx, *wp = ws + L.wp, *wa = ws + L.wa,
          *wb = ws + L.wb;
    int ppc, nch;
    gyro_x_plan(B, P, D, &ppc, &nch);
    float* dst = nch == 1 ? gx : wx;
    {
        dim3 grid((unsigned)((B + kGyroBxThreads - 1) / kGyroBxThreads), (unsigned)nch);
        constexpr uint32_t kGeo = HVAE_GYRO_PVAE | HVAE_GYRO_SIGNED;
        auto kern = aliased ? (prm.flags == HVAE_GYRO_SIGNED ? k_gyro_bwd_pairs<D4, true, (int)HVAE_GYRO_SIGNED> : k_gyro_bwd_pairs<D4, true, -1>)
                            : (prm.flags == kGeo ? k_gyro_bwd_pairs<D4, false, (int)kGeo>
                               : prm.flags == (kGeo | HVAE_GYRO_RELU) ? k_gyro_bwd_pairs<D4, false, (int)(kGeo | HVAE_GYRO_RELU)>
                                                                      : k_gyro_bwd_pairs<D4, false, -1>);
        kern<<<grid, kGyroBxThreads, 0, s>>>(x, p, a, bias, gout, dst, CP, CA, wsum, (int)B, (int)D, (int)P, ppc, prm);
    }
    SlabReducer red;   // gx over plane chunks, gp / ga / gbias over row slabs: one launch (hvae_common.cuh)
    if (nch > 1) red.add(wx, gx, B * D, nch);
    if (gp || ga || gbias) {
        const int slabs = gyro_slabs(B, P);
        const int rows_per_slab = (int)(((B + slabs - 1) / slabs + kGyroBpSlabGran - 1) / kGyroBpSlabGran * kGyroBpSlabGran);
        const int nsl = (int)((B + rows_per_slab - 1) / rows_per_slab);
        dim3 grid((unsigned)((P + kGyroBpThreads - 1) / kGyroBpThreads), (unsigned)nsl);
        if (aliased) k_gyro_bwd_planes<D4, true><<<grid, kGyroBpThreads, 0, s>>>(x, p, a, CP, CA, wsum, wp, wa, wb, (int)B, (int)D, (int)P, rows_per_slab);
        else         k_gyro_bwd_planes<D4, false><<<grid, kGyroBpThreads, 0, s>>>(x, p, a, CP, CA, wsum, wp, wa, wb, (int)B, (int)D, (int)P, rows_per_slab);
        const int64_t n = P * D;
        if (gp) red.add(wp, gp, n, nsl);
        if (ga && !aliased) red.add(wa, ga, n, nsl);
        if (gbias) red.add(wb, gbias, P, nsl);
    }
    red.launch(s);
    return check_launch();
}

}  // namespace hvae

using namespace hvae;

constexpr int64_t kGyroMaxD = 64;

extern "C" int hvae_gyroplane_fwd_f32(const float* x, const float* p, const float* a, const float* bias, float* out,
                                      int64_t B, int64_t D, int64_t P, float c, uint32_t flags, void* stream) {
    if (B < 0 || P < 0 || D <= 0 || D > kGyroMaxD) return HVAE_ESHAPE;
    if (B == 0 || P == 0) return HVAE_OK;
    if (!x || !p || !a || !out) return HVAE_EARG;
    const GyroParams prm = make_gyro_params(c, flags);
    cudaStream_t s = (cudaStream_t)stream;
    if (D <= 4) return gyro_fwd_launch<4>(x, p, a, bias, out, B, D, P, prm, s);
    if (D <= 8) return gyro_fwd_launch<8>(x, p, a, bias, out, B, D, P, prm, s);
    if (D <= 12) return gyro_fwd_launch<12>(x, p, a, bias, out, B, D, P, prm, s);  // latent 10 (config 2): 25 % less padding than 16
    if (D <= 16) return gyro_fwd_launch<16>(x, p, a, bias, out, B, D, P, prm, s);
    if (D <= 32) return gyro_fwd_launch<32>(x, p, a, bias, out, B, D, P, prm, s);
    return gyro_fwd_launch<64>(x, p, a, bias, out, B, D, P, prm, s);
}

extern "C" size_t hvae_gyroplane_bwd_workspace_bytes(int64_t B, int64_t D, int64_t P) {
    if (B <= 0 || P <= 0 || D <= 0) return 0;
    return sizeof(float) * gyro_ws_layout(B, D, P).total;
}

extern "C" int hvae_gyroplane_relu_bwd_f32(const float* x, const float* p, const float* a, const float* bias, const float* gout,
                                           float* gx, float* gp, float* ga, float* gbias, int64_t B, int64_t D, int64_t P,
                                           float c, uint32_t flags, void* workspace, size_t workspace_bytes, void* stream) {
    if (B < 0 || P < 0 || D <= 0 || D > kGyroMaxD) return HVAE_ESHAPE;
    if (B == 0 || P == 0) return HVAE_OK;
    if (!x || !p || !a || !gout) return HVAE_EARG;
    if (a != p && gp && !ga) return HVAE_EARG;
    if (!gx) return HVAE_EARG;
    if (!workspace || workspace_bytes < hvae_gyroplane_bwd_workspace_bytes(B, D, P)) return HVAE_EARG;
    const GyroParams prm = make_gyro_params(c, flags);
    cudaStream_t s = (cudaStream_t)stream;
    float* ws = (float*)workspace;
    if (D <= 4) return gyro_bwd_launch<4>(x, p, a, bias, gout, gx, gp, ga, gbias, B, D, P, prm, ws, s);
    if (D <= 8) return gyro_bwd_launch<8>(x, p, a, bias, gout, gx, gp, ga, gbias, B, D, P, prm, ws, s);
    if (D <= 12) return gyro_bwd_launch<12>(x, p, a, bias, gout, gx, gp, ga, gbias, B, D, P, prm, ws, s);
    if (D <= 16) return gyro_bwd_launch<16>(x, p, a, bias, gout, gx, gp, ga, gbias, B, D, P, prm, ws, s);
    if (D <= 32) return gyro_bwd_launch<32>(x, p, a, bias, gout, gx, gp, ga, gbias, B, D, P, prm, ws, s);
    return gyro_bwd_launch<64>(x, p, a, bias, gout, gx, gp, ga, gbias, B, D, P, prm, ws, s);
}

extern "C" int hvae_gyroplane_bwd_f32(const float* x, const float* p, const float* a, const float* gout, float* gx,
                                      float* gp, float* ga, float* gbias, int64_t B, int64_t D, int64_t P, float c,
                                      uint32_t flags, void* workspace, size_t workspace_bytes, void* stream) {
    return hvae_gyroplane_relu_bwd_f32(x, p, a, nullptr, gout, gx, gp, ga, gbias, B, D, P, c, flags, workspace, workspace_bytes, stream);
}

// the grid the backward's pair kernel runs for a problem (host-only; bench.py / tests report it)
extern "C" int hvae_gyroplane_bwd_plan(int64_t B, int64_t D, int64_t P, int* planes_per_chunk, int* chunks, int* ctas) {
    if (B <= 0 || P <= 0 || D <= 0 || D > kGyroMaxD) return HVAE_ESHAPE;
    int ppc = 0, nch = 0;
    gyro_x_plan(B, P, D, &ppc, &nch);
    if (planes_per_chunk) *planes_per_chunk = ppc;
    if (chunks) *chunks = nch;
    if (ctas) *ctas = (int)(((B + kGyroBxThreads - 1) / kGyroBxThreads) * nch);
    return HVAE_OK;
}
