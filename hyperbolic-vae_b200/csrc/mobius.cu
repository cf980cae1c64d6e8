// K1b (Riemannian-layer weight prep) and K1 (Mobius matvec) — SIMT path.
//
// reference: hyperbolic_vae/layers.py:58-67 (weight/bias properties, recomputed on EVERY forward),
//            layers.py:145-147 -> geoopt mobius_matvec + project (App. A.1).
//
// K1 here is the skinny-N regime of the encoder (P = latent dim 2..64, F = 512/600): it is bound by the
// single read of x (B*F*4 bytes), so the design is: one warp owns R rows of x in registers, the
// transported weight M sits in shared memory, and the norm / artanh / tanh rescale + projection is the
// warp's epilogue — x is read once, y (and mx for backward) written once.
// Algorithmic bytes: fwd 4(BF + PF + 2BP), bwd 4(BF + PF + 3BP) read + 4(BF + PF) written.
// Larger P runs the same kernel in chunks of 32 planes (correct, not fast); the tcgen05 path is K1-TC.
#include "hvae_common.cuh"
#include "row_maps.cuh"
#include "mobius_row.cuh"

namespace hvae {

// =================================================================================================
// K1b weight prep.  Row j:  t = W_j * beta_j ; bpt_j = expmap0(t) ; m_j = clamp_min(1 - c|bpt_j|^2) ; M_j = W_j m_j
// =================================================================================================
template <int G, int EPL>
__global__ void __launch_bounds__(kRowThreads)
k_weight_prep_fwd(const float* __restrict__ W, const float* __restrict__ beta, const float* __restrict__ bias_pt_in,
                  float* __restrict__ bpt, float* __restrict__ M, int64_t P, int F, Ball ball) {
    HVAE_ROW_PROLOGUE(G)
    for (int64_t r0 = warp_global * RPW; r0 < P; r0 += warps_total * RPW) {
        const int64_t row = r0 + sub;
        const bool valid = row < P;
        RowSlice<G, EPL> w, b;
        w.load(W, row, F, lg, valid);
        if (bias_pt_in) {
            b.load(bias_pt_in, row, F, lg, valid);
        } else {
            const float bj = valid ? __ldg(beta + row) : 0.0f;
            RowSlice<G, EPL> t;
#pragma unroll
            for (int i = 0; i < EPL; ++i) t.v[i] = w.v[i] * bj;
            float n_raw, n, th, pn;
            expmap0_row<G, EPL>(t, b, ball, n_raw, n, th);
            project_inplace<G, EPL>(b, ball, pn);
            if (bpt) b.store(bpt, row, F, lg, valid);
        }
        const float m = fmaxf(1.0f - ball.c * sqnorm<G, EPL>(b), kMinNorm);
#pragma unroll
        for (int i = 0; i < EPL; ++i) w.v[i] *= m;
        w.store(M, row, F, lg, valid);
    }
}

template <int G, int EPL>
__global__ void __launch_bounds__(kRowThreads)
k_weight_prep_bwd(const float* __restrict__ W, const float* __restrict__ beta, const float* __restrict__ bias_pt_in,
                  const float* __restrict__ gM, const float* __restrict__ gbpt, float* __restrict__ gW,
                  float* __restrict__ gbeta, float* __restrict__ gbias_pt, int64_t P, int F, Ball ball) {
    HVAE_ROW_PROLOGUE(G)
    for (int64_t r0 = warp_global * RPW; r0 < P; r0 += warps_total * RPW) {
        const int64_t row = r0 + sub;
        const bool valid = row < P;
        RowSlice<G, EPL> w, b, t, gm, gb;
        w.load(W, row, F, lg, valid);
        float bj = 0.0f;
        if (bias_pt_in) {
            b.load(bias_pt_in, row, F, lg, valid);
        } else {
            bj = valid ? __ldg(beta + row) : 0.0f;
#pragma unroll
            for (int i = 0; i < EPL; ++i) t.v[i] = w.v[i] * bj;
            float n_raw, n, th, pn;
            expmap0_row<G, EPL>(t, b, ball, n_raw, n, th);
            project_inplace<G, EPL>(b, ball, pn);
        }
        if (gM) gm.load(gM, row, F, lg, valid); else gm.zero();
        if (gbpt) gb.load(gbpt, row, F, lg, valid); else gb.zero();
        const float m0 = 1.0f - ball.c * sqnorm<G, EPL>(b);
        const float m = fmaxf(m0, kMinNorm);
        // M = W m(b):  gW += m gM ;  g_b += (gM . W) dm/db = (gM . W)(-2 c b) when unclamped
        const float gmw = dot<G, EPL>(gm, w);
        const float coef = (m0 >= kMinNorm) ? -2.0f * ball.c * gmw : 0.0f;
#pragma unroll
        for (int i = 0; i < EPL; ++i) gb.v[i] += coef * b.v[i];
        if (bias_pt_in) {
#pragma unroll
            for (int i = 0; i < EPL; ++i) gm.v[i] *= m;
            gm.store(gW, row, F, lg, valid);
            if (gbias_pt) gb.store(gbias_pt, row, F, lg, valid);
        } else {
            expmap0_row_bwd<G, EPL>(t, gb, ball);  // gb <- dL/dt
            const float gbeta_j = dot<G, EPL>(gb, w);
#pragma unroll
            for (int i = 0; i < EPL; ++i) gm.v[i] = m * gm.v[i] + bj * gb.v[i];
            gm.store(gW, row, F, lg, valid);
            if (gbeta && valid && lg == 0) gbeta[row] = gbeta_j;
        }
    }
}

// =================================================================================================
// K1 Mobius matvec, SIMT.  warp = R rows; lane i holds x[r][lane + 32 i]; M chunk (<=32 planes) in smem.
// =================================================================================================
constexpr int kMobThreads = 256;
constexpr int kMobWarps = kMobThreads / 32;
constexpr int kMobChunk = 32;  // planes per smem chunk

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int EPLF, int R>
__global__ void __launch_bounds__(kMobThreads)
k_mobius_fwd(const float* __restrict__ x, const float* __restrict__ M, float* __restrict__ y, float* __restrict__ mx_out,
             int64_t B, int F, int P, Ball ball) {
    extern __shared__ float Ms[];  // [chunk planes][F]
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nchunks = (P + kMobChunk - 1) / kMobChunk;
    const int64_t rows_per_iter = (int64_t)gridDim.x * kMobWarps * R;
    bool staged = false;
    for (int64_t base = 0; base < B; base += rows_per_iter) {
        const int64_t row0 = base + ((int64_t)blockIdx.x * kMobWarps + warp) * R;
        float xr[R][EPLF];
        float x2[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float s = 0.0f;
#pragma unroll
            for (int i = 0; i < EPLF; ++i) {
                const int f = lane + 32 * i;
                xr[r][i] = (row0 + r < B && f < F) ? __ldg(x + (row0 + r) * F + f) : 0.0f;
                s = fmaf(xr[r][i], xr[r][i], s);
            }
            x2[r] = warp_sum(s);
        }
        float mxv[R];   // lane j holds mx[r][chunk*32 + j] of the CURRENT chunk
        float mx2[R];
        bool nz[R];
#pragma unroll
        for (int r = 0; r < R; ++r) { mx2[r] = 0.0f; nz[r] = false; mxv[r] = 0.0f; }
        for (int ch = 0; ch < nchunks; ++ch) {
            const int j0 = ch * kMobChunk;
            const int jn = min(kMobChunk, P - j0);
            if (nchunks > 1 || !staged) {
                __syncthreads();
                for (int i = threadIdx.x; i < jn * F; i += kMobThreads) Ms[i] = __ldg(M + (int64_t)j0 * F + i);
                __syncthreads();
                staged = true;
            }
#pragma unroll
            for (int r = 0; r < R; ++r) mxv[r] = 0.0f;
            for (int jj = 0; jj < jn; ++jj) {
                float acc[R];
#pragma unroll
                for (int r = 0; r < R; ++r) acc[r] = 0.0f;
                const float* mrow = Ms + jj * F;
#pragma unroll
                for (int i = 0; i < EPLF; ++i) {
                    const int f = lane + 32 * i;
                    const float mv = (f < F) ? mrow[f] : 0.0f;
#pragma unroll
                    for (int r = 0; r < R; ++r) acc[r] = fmaf(xr[r][i], mv, acc[r]);
                }
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const float s = warp_sum(acc[r]);
                    if (lane == jj) mxv[r] = s;
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const bool on = lane < jn;
                mx2[r] += warp_sum(on ? mxv[r] * mxv[r] : 0.0f);
                nz[r] = nz[r] || (__ballot_sync(0xffffffffu, on && mxv[r] != 0.0f) != 0u);
                if (mx_out && on && row0 + r < B) mx_out[(row0 + r) * P + j0 + lane] = mxv[r];
            }
        }
        // epilogue: rescale + project, chunk by chunk (single chunk: straight from registers)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (row0 + r >= B) continue;   // warp-uniform
            MobRow rs;
            mob_row_scalars(x2[r], mx2[r], !nz[r], ball, rs);
            if (nchunks == 1) {
                if (lane < P) {
                    float v = rs.psi * mxv[r];
                    if (rs.hit) v = v / rs.ypn * ball.maxnorm;
                    y[(row0 + r) * P + lane] = rs.zero_row ? 0.0f : v;
                }
            } else {
                for (int j = lane; j < P; j += 32) {
                    const float m = mx_out[(row0 + r) * P + j];
                    float v = rs.psi * m;
                    if (rs.hit) v = v / rs.ypn * ball.maxnorm;
                    y[(row0 + r) * P + j] = rs.zero_row ? 0.0f : v;
                }
            }
        }
    }
}

// backward, kernel 1: per row -> gmx (B,P) to workspace, gx (B,F)
template <int EPLF, int R>
__global__ void __launch_bounds__(kMobThreads)
k_mobius_bwd_x(const float* __restrict__ x, const float* __restrict__ M, const float* __restrict__ mx,
               const float* __restrict__ gy, float* __restrict__ gx, float* __restrict__ gmx, int64_t B, int F, int P,
               Ball ball) {
    extern __shared__ float Ms[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nchunks = (P + kMobChunk - 1) / kMobChunk;
    const int64_t rows_per_iter = (int64_t)gridDim.x * kMobWarps * R;
    bool staged = false;
    for (int64_t base = 0; base < B; base += rows_per_iter) {
        const int64_t row0 = base + ((int64_t)blockIdx.x * kMobWarps + warp) * R;
        float xr[R][EPLF], acc[R][EPLF];
        float x2[R], mx2[R], gdm[R];
        bool nz[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float s = 0.0f;
#pragma unroll
            for (int i = 0; i < EPLF; ++i) {
                const int f = lane + 32 * i;
                xr[r][i] = (row0 + r < B && f < F) ? __ldg(x + (row0 + r) * F + f) : 0.0f;
                s = fmaf(xr[r][i], xr[r][i], s);
                acc[r][i] = 0.0f;
            }
            x2[r] = warp_sum(s);
            // row dots over P:  |mx|^2, <gy, mx>, any(mx != 0)
            float a = 0.0f, b = 0.0f;
            bool any = false;
            if (row0 + r < B) {
                for (int j = lane; j < P; j += 32) {
                    const float m = __ldg(mx + (row0 + r) * P + j);
                    const float g = __ldg(gy + (row0 + r) * P + j);
                    a = fmaf(m, m, a);
                    b = fmaf(g, m, b);
                    any = any || (m != 0.0f);
                }
            }
            mx2[r] = warp_sum(a);
            gdm[r] = warp_sum(b);
            nz[r] = __ballot_sync(0xffffffffu, any) != 0u;
        }
        // per-row coefficients: gmx_j = alpha * gy_j + beta * mx_j ;  gx += gxn_coef * x
        float alpha[R], beta[R], gxc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) mob_bwd_coefs(x2[r], mx2[r], gdm[r], !nz[r], ball, alpha[r], beta[r], gxc[r]);
        for (int ch = 0; ch < nchunks; ++ch) {
            const int j0 = ch * kMobChunk;
            const int jn = min(kMobChunk, P - j0);
            if (nchunks > 1 || !staged) {
                __syncthreads();
                for (int i = threadIdx.x; i < jn * F; i += kMobThreads) Ms[i] = __ldg(M + (int64_t)j0 * F + i);
                __syncthreads();
                staged = true;
            }
            float gm[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                gm[r] = 0.0f;
                if (lane < jn && row0 + r < B) {
                    const int64_t o = (row0 + r) * P + j0 + lane;
                    gm[r] = alpha[r] * __ldg(gy + o) + beta[r] * __ldg(mx + o);
                    gmx[o] = gm[r];
                }
            }
            for (int jj = 0; jj < jn; ++jj) {
                const float* mrow = Ms + jj * F;
                float gj[R];
#pragma unroll
                for (int r = 0; r < R; ++r) gj[r] = __shfl_sync(0xffffffffu, gm[r], jj);
#pragma unroll
                for (int i = 0; i < EPLF; ++i) {
                    const int f = lane + 32 * i;
                    const float mv = (f < F) ? mrow[f] : 0.0f;
#pragma unroll
                    for (int r = 0; r < R; ++r) acc[r][i] = fmaf(gj[r], mv, acc[r][i]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (row0 + r >= B) continue;
#pragma unroll
            for (int i = 0; i < EPLF; ++i) {
                const int f = lane + 32 * i;
                if (f < F) gx[(row0 + r) * F + f] = acc[r][i] + gxc[r] * xr[r][i];
            }
        }
    }
}

// backward, kernel 2: gM partials.  thread = one feature column f; holds acc[<=32 planes]; rows streamed.
constexpr int kMobGmThreads = 128;
constexpr int kMobGmTB = 32;  // rows per smem stage of gmx

template <int PC /* planes per CTA: 8, 16 or 32 */>
__global__ void __launch_bounds__(kMobGmThreads)
k_mobius_bwd_m(const float* __restrict__ x, const float* __restrict__ gmx, float* __restrict__ wM, int64_t B, int F, int P,
               int rows_per_slab) {
    __shared__ float gs[kMobGmTB][PC];
    const int f = blockIdx.x * kMobGmThreads + threadIdx.x;
    const int j0 = blockIdx.y * PC;
    const int jn = min(PC, P - j0);
    const int slab = blockIdx.z;
    const int64_t bs = (int64_t)slab * rows_per_slab;
    const int64_t be = min(B, bs + (int64_t)rows_per_slab);
    float acc[PC];
#pragma unroll
    for (int j = 0; j < PC; ++j) acc[j] = 0.0f;
    for (int64_t b0 = bs; b0 < be; b0 += kMobGmTB) {
        __syncthreads();
        for (int i = threadIdx.x; i < kMobGmTB * PC; i += kMobGmThreads) {
            const int bb = i / PC, jj = i - bb * PC;
            gs[bb][jj] = (b0 + bb < be && jj < jn) ? __ldg(gmx + (b0 + bb) * P + j0 + jj) : 0.0f;
        }
        __syncthreads();
        const int bn = (int)min((int64_t)kMobGmTB, be - b0);
        if (f < F) {
            // latency-bound (a few warps per SM, every x load a trip to L2): 16 row loads in flight before the FMAs
            constexpr int U = 16;
            static_assert(kMobGmTB % U == 0, "row stage must be a multiple of the load batch");
            for (int bb = 0; bb < bn; bb += U) {
                float xv[U];
#pragma unroll
                for (int u = 0; u < U; ++u) xv[u] = (bb + u < bn) ? __ldg(x + (b0 + bb + u) * F + f) : 0.0f;
#pragma unroll
                for (int u = 0; u < U; ++u) {
#pragma unroll
                    for (int j = 0; j < PC; j += 4) {
                        const float4 g4 = *reinterpret_cast<const float4*>(&gs[bb + u][j]);
                        acc[j] = fmaf(g4.x, xv[u], acc[j]); acc[j + 1] = fmaf(g4.y, xv[u], acc[j + 1]);
                        acc[j + 2] = fmaf(g4.z, xv[u], acc[j + 2]); acc[j + 3] = fmaf(g4.w, xv[u], acc[j + 3]);
                    }
                }
            }
        }
    }
    if (f < F) {
#pragma unroll
        for (int j = 0; j < PC; ++j)
            if (j < jn) wM[((int64_t)slab * P + j0 + j) * F + f] = acc[j];
    }
}

inline int mob_pc(int64_t P) { return P <= 8 ? 8 : (P <= 16 ? 16 : 32); }
inline int mob_slabs(int64_t B, int64_t F, int64_t P) {
    const int pc = mob_pc(P);
    const int64_t tiles = ((F + kMobGmThreads - 1) / kMobGmThreads) * ((P + pc - 1) / pc);
    int64_t want = (4 * kNumSMs + tiles - 1) / tiles;
    const int64_t maxs = (B + kMobGmTB - 1) / kMobGmTB;
    if (want > maxs) want = maxs;
    if (want > 256) want = 256;
    if (want < 1) want = 1;
    return (int)want;
}

template <int EPLF, int R>
int mob_fwd_launch(const float* x, const float* M, float* y, float* mx, int64_t B, int64_t F, int64_t P, const Ball& ball,
                   cudaStream_t s) {
    const int jn = (int)(P < kMobChunk ? P : kMobChunk);
    const size_t smem = sizeof(float) * (size_t)jn * F;
    auto kern = k_mobius_fwd<EPLF, R>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int64_t blocks = (B + kMobWarps * R - 1) / (kMobWarps * R);
    const int64_t cap = (int64_t)kNumSMs * (smem > 64 * 1024 ? 1 : 2);
    if (blocks > cap) blocks = cap;
    kern<<<(unsigned)blocks, kMobThreads, smem, s>>>(x, M, y, mx, B, (int)F, (int)P, ball);
    return check_launch();
}

template <int EPLF, int R>
int mob_bwd_launch(const float* x, const float* M, const float* mx, const float* gy, float* gx, float* gM, int64_t B,
                   int64_t F, int64_t P, const Ball& ball, float* ws, cudaStream_t s) {
    const int jn = (int)(P < kMobChunk ? P : kMobChunk);
    const size_t smem = sizeof(float) * (size_t)jn * F;
    auto kern = k_mobius_bwd_x<EPLF, R>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int64_t blocks = (B + kMobWarps * R - 1) / (kMobWarps * R);
    const int64_t cap = (int64_t)kNumSMs * (smem > 64 * 1024 ? 1 : 2);
    if (blocks > cap) blocks = cap;
    float* gmx = ws;                       // (B,P)
    float* wM = ws + B * P;                // [slabs][P][F]
    kern<<<(unsigned)blocks, kMobThreads, smem, s>>>(x, M, mx, gy, gx, gmx, B, (int)F, (int)P, ball);
    if (gM) {
        const int slabs = mob_slabs(B, F, P);
        const int rows_per_slab = (int)(((B + slabs - 1) / slabs + kMobGmTB - 1) / kMobGmTB * kMobGmTB);
        const int pc = mob_pc(P);
        dim3 grid((unsigned)((F + kMobGmThreads - 1) / kMobGmThreads), (unsigned)((P + pc - 1) / pc), (unsigned)slabs);
        if (pc == 8)       k_mobius_bwd_m<8><<<grid, kMobGmThreads, 0, s>>>(x, gmx, wM, B, (int)F, (int)P, rows_per_slab);
        else if (pc == 16) k_mobius_bwd_m<16><<<grid, kMobGmThreads, 0, s>>>(x, gmx, wM, B, (int)F, (int)P, rows_per_slab);
        else               k_mobius_bwd_m<32><<<grid, kMobGmThreads, 0, s>>>(x, gmx, wM, B, (int)F, (int)P, rows_per_slab);
        SlabReducer red;   // hvae_common.cuh: slab-parallel, coalesced, fixed summation order
        red.add(wM, gM, P * F, slabs);
        red.launch(s);
    }
    return check_launch();
}

}  // namespace hvae

using namespace hvae;

constexpr int64_t kMobMaxF = 1024;

extern "C" int hvae_weight_prep_fwd_f32(const float* W, const float* beta, const float* bias_pt_in, float* bpt, float* M,
                                        int64_t P, int64_t F, float c, void* stream) {
    if (P < 0 || F <= 0 || F > kMaxRowDim) return HVAE_ESHAPE;
    if (P == 0) return HVAE_OK;
    if (!W || !M || (!beta && !bias_pt_in)) return HVAE_EARG;
    HVAE_ROW_DISPATCH(k_weight_prep_fwd, F, P, (cudaStream_t)stream, W, beta, bias_pt_in, bpt, M, P, (int)F, make_ball(c));
    return check_launch();
}

extern "C" int hvae_weight_prep_bwd_f32(const float* W, const float* beta, const float* bias_pt_in, const float* gM,
                                        const float* gbpt, float* gW, float* gbeta, float* gbias_pt, int64_t P, int64_t F,
                                        float c, void* stream) {
    if (P < 0 || F <= 0 || F > kMaxRowDim) return HVAE_ESHAPE;
    if (P == 0) return HVAE_OK;
    if (!W || !gW || (!beta && !bias_pt_in)) return HVAE_EARG;
    HVAE_ROW_DISPATCH(k_weight_prep_bwd, F, P, (cudaStream_t)stream, W, beta, bias_pt_in, gM, gbpt, gW, gbeta, gbias_pt, P,
                      (int)F, make_ball(c));
    return check_launch();
}

extern "C" int hvae_mobius_matvec_fwd_f32(const float* x, const float* M, float* y, float* mx_out, int64_t B, int64_t F,
                                          int64_t P, float c, void* stream) {
    if (B < 0 || P <= 0 || F <= 0 || F > kMobMaxF) return HVAE_ESHAPE;
    if (B == 0) return HVAE_OK;
    if (!x || !M || !y) return HVAE_EARG;
    if (P > kMobChunk && !mx_out) return HVAE_EARG;  // the chunked path stages mx through its output buffer
    const Ball ball = make_ball(c);
    cudaStream_t s = (cudaStream_t)stream;
    if (F <= 128) return mob_fwd_launch<4, 4>(x, M, y, mx_out, B, F, P, ball, s);
    if (F <= 256) return mob_fwd_launch<8, 4>(x, M, y, mx_out, B, F, P, ball, s);
    if (F <= 512) return mob_fwd_launch<16, 4>(x, M, y, mx_out, B, F, P, ball, s);
    if (F <= 768) return mob_fwd_launch<24, 2>(x, M, y, mx_out, B, F, P, ball, s);
    return mob_fwd_launch<32, 2>(x, M, y, mx_out, B, F, P, ball, s);
}

extern "C" size_t hvae_mobius_matvec_bwd_workspace_bytes(int64_t B, int64_t F, int64_t P) {
    if (B <= 0 || F <= 0 || P <= 0) return 0;
    return sizeof(float) * ((size_t)B * P + (size_t)mob_slabs(B, F, P) * P * F);
}

extern "C" int hvae_mobius_matvec_bwd_f32(const float* x, const float* M, const float* mx, const float* gy, float* gx,
                                          float* gM, int64_t B, int64_t F, int64_t P, float c, void* workspace,
                                          size_t workspace_bytes, void* stream) {
    if (B < 0 || P <= 0 || F <= 0 || F > kMobMaxF) return HVAE_ESHAPE;
    if (B == 0) return HVAE_OK;
    if (!x || !M || !mx || !gy || !gx) return HVAE_EARG;
    if (!workspace || workspace_bytes < hvae_mobius_matvec_bwd_workspace_bytes(B, F, P)) return HVAE_EARG;
    const Ball ball = make_ball(c);
    cudaStream_t s = (cudaStream_t)stream;
    float* ws = (float*)workspace;
    if (F <= 128) return mob_bwd_launch<4, 4>(x, M, mx, gy, gx, gM, B, F, P, ball, ws, s);
    if (F <= 256) return mob_bwd_launch<8, 4>(x, M, mx, gy, gx, gM, B, F, P, ball, ws, s);
    if (F <= 512) return mob_bwd_launch<16, 2>(x, M, mx, gy, gx, gM, B, F, P, ball, ws, s);
    if (F <= 768) return mob_bwd_launch<24, 2>(x, M, mx, gy, gx, gM, B, F, P, ball, ws, s);
    return mob_bwd_launch<32, 1>(x, M, mx, gy, gx, gM, B, F, P, ball, ws, s);
}
