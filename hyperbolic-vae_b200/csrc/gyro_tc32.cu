// K2 for GEMM-sized latent dims in fp32 mode: hyperplane distances of ANY (a, p) - the a == p layer (layers.py:193-210,
// geoopt Distance2StereographicHyperplanes) and GeodesicLayer / normdist2plane (layers.py:96-121 -> manifolds.py:41-65,
// HVAE_GYRO_PVAE) - forward AND backward at fp32 accuracy on the tensor cores.
//
// The SIMT kernels (gyroplane.cu) hold a row of x and a tile of planes in registers / shared memory and stop at D = 64;
// the bf16 tensor-core kernels (tc_gemm2.cu) carry 4e-3 on the inner products.  Here the two inner products every pair
// needs, <x_b, p_j> and <x_b, a_j>, come from the fp32-accurate split-operand GEMM (tc_gemm.cu, EPI_X3), and everything else is
// the scalar pair function shared with the other paths (gyro_pair.cuh) applied elementwise:
//   forward   PX = x p^T, XA = x a^T  (2 GEMMs)  ->  out[b][j] = pair(PX, XA, |x_b|^2, |p_j|^2, <p_j,a_j>, |a_j|) + bias_j
//   backward  recompute PX, XA;  pair gradients  CP = dL/dPX, CA = dL/dXA  (in place), row sums of dL/d|x|^2, column sums
//             of dL/d|p|^2, dL/d<p,a>, dL/d|a|;  then four GEMMs (three-way bf16 split: see the note at the call)
//               gx = CP p + CA a + 2 (sum_j dx2) x          gp = CP^T x + 2 (sum_b dp2) p + (sum_b dpa) a
//                                                           ga = CA^T x + (sum_b dpa) p + (sum_b dan) a / |a|
//   (a == p: one inner product, CP = dPX + dXA, gp = CP^T x + [2 (dp2 + dpa) + dan / |p|] p.)
// The inner-product form of the pair function loses |p|^2 eps absolutely when x -> p (the SIMT path's difference form
// does not): results are accurate to 1e-5 times the pair conditioning, like every other path here.
// Reductions are two-level and ordered: bit-reproducible.
#include "gyro_pair.cuh"
#include "hvae_common.cuh"

namespace hvae {
namespace gtc32 {

__global__ void __launch_bounds__(256) k_rows_sq(const float* __restrict__ x, float* __restrict__ out, int64_t R, int64_t D) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < R; r += nw) {
        float s = 0.0f;
        for (int64_t i = lane; i < D; i += 32) { const float v = __ldg(x + r * D + i); s = fmaf(v, v, s); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) out[r] = s;
    }
}

// per plane: |p|^2, <p, a>, |a|  (a == p: |p|^2, |p|^2, |p|)
__global__ void __launch_bounds__(256) k_plane_stats(const float* __restrict__ p, const float* __restrict__ a, float* __restrict__ p2,
                                                     float* __restrict__ pa, float* __restrict__ an, int64_t P, int64_t D) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t j = warp; j < P; j += nw) {
        float s2 = 0.0f, spa = 0.0f, sa = 0.0f;
        for (int64_t i = lane; i < D; i += 32) {
            const float pv = __ldg(p + j * D + i), av = __ldg(a + j * D + i);
            s2 = fmaf(pv, pv, s2);
            spa = fmaf(pv, av, spa);
            sa = fmaf(av, av, sa);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            spa += __shfl_xor_sync(0xffffffffu, spa, o);
            sa += __shfl_xor_sync(0xffffffffu, sa, o);
        }
        if (lane == 0) { p2[j] = s2; pa[j] = spa; an[j] = sqrtf(sa); }
    }
}

__global__ void __launch_bounds__(256)
k_dots_fwd(const float* __restrict__ PX, const float* __restrict__ XA, const float* __restrict__ x2, const float* __restrict__ p2,
           const float* __restrict__ pa, const float* __restrict__ an, const float* __restrict__ bias, float* __restrict__ out,
           int64_t B, int64_t P, GyroParams prm) {
    const int64_t n = B * P;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / P, j = i - b * P;
        GyroPairCtx k;
        const float px = __ldg(PX + i);
        const float o = gyro_pair_fwd(px, XA ? __ldg(XA + i) : px, __ldg(x2 + b), __ldg(p2 + j), __ldg(pa + j), __ldg(an + j), prm, k);
        out[i] = o + (bias ? __ldg(bias + j) : 0.0f);
    }
}

// Pair gradients over a (64 rows x 256 planes) tile, thread = plane.  PX / XA are overwritten with CP = dL/dPX, CA = dL/dXA
// (a == p: XA == NULL and CP = dPX + dXA); rowpart[cb][b] = this tile's sum over planes of dL/d|x_b|^2; colpart[rb][q][j],
// q = 0..3: this tile's sums over rows of dL/d|p_j|^2, dL/d<p_j,a_j>, dL/d|a_j| and of the one part of dL/d|a_j| that does
// not cancel (HVAE_GYRO_SCALED: out = |a| out0), which k_gp_combine uses to pin the radial component of ga.
constexpr int kCols = 256, kRows = 64;
__global__ void __launch_bounds__(kCols)
k_dots_bwd(const float* __restrict__ g, float* __restrict__ PX, float* __restrict__ XA, const float* __restrict__ x2,
           const float* __restrict__ p2, const float* __restrict__ pa, const float* __restrict__ an, float* __restrict__ rowpart,
           float* __restrict__ colpart, int64_t B, int64_t P, GyroParams prm) {
    __shared__ float rs[kCols / 32][kRows];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t j = (int64_t)blockIdx.y * kCols + tid;   // row blocks on grid.x (can exceed 65535), plane blocks on grid.y
    const int64_t b0 = (int64_t)blockIdx.x * kRows;
    const bool jok = j < P;
    const float p2j = jok ? __ldg(p2 + j) : 0.0f, paj = jok ? __ldg(pa + j) : 0.0f, anj = jok ? __ldg(an + j) : 0.0f;
    float c0 = 0.0f, c1 = 0.0f, c2 = 0.0f, c3 = 0.0f;
    for (int r = 0; r < kRows; ++r) {
        const int64_t b = b0 + r;
        float dx2 = 0.0f;
        if (jok && b < B) {
            const int64_t i = b * P + j;
            const float pxv = PX[i], xav = XA ? XA[i] : pxv, gv = __ldg(g + i), x2v = __ldg(x2 + b);
            GyroPairCtx k;
            gyro_pair_fwd(pxv, xav, x2v, p2j, paj, anj, prm, k);
            const GyroPairGrad gr = gyro_pair_bwd(gv, pxv, xav, x2v, p2j, paj, anj, prm, k);
            if (XA) { PX[i] = gr.dpx; XA[i] = gr.dxa; } else { PX[i] = gr.dpx + gr.dxa; }
            dx2 = gr.dx2;
            c0 += gr.dp2; c1 += gr.dpa; c2 += gr.dan;
            if (prm.flags & HVAE_GYRO_SCALED) {   // (gyro_pair_bwd's g1 * out0)
                float g1 = gv;
                if (prm.flags & HVAE_GYRO_SQUARED) g1 = (prm.flags & HVAE_GYRO_SIGNED) ? gv * 2.0f * fabsf(k.out1) : gv * 2.0f * k.out1;
                c3 += g1 * k.out0;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dx2 += __shfl_xor_sync(0xffffffffu, dx2, o);
        if (lane == 0) rs[warp][r] = dx2;
    }
    __syncthreads();
    if (tid < kRows && b0 + tid < B) {
        float a = 0.0f;
#pragma unroll
        for (int w = 0; w < kCols / 32; ++w) a += rs[w][tid];
        rowpart[(int64_t)blockIdx.y * B + b0 + tid] = a;
    }
    if (jok) {
        float* cp = colpart + (int64_t)blockIdx.x * 4 * P;
        cp[j] = c0; cp[P + j] = c1; cp[2 * P + j] = c2; cp[3 * P + j] = c3;
    }
}

// out[i] = sum_k part[k * n + i]   (ordered)
__global__ void __launch_bounds__(256) k_sum_parts(const float* __restrict__ part, float* __restrict__ out, int64_t n, int64_t nparts) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float a = 0.0f;
        for (int64_t k = 0; k < nparts; ++k) a += part[k * n + i];
        out[i] = a;
    }
}

// gx = g1 (+ g2) + 2 rdx2_b x
__global__ void __launch_bounds__(256) k_gx_combine(const float* __restrict__ g1, const float* __restrict__ g2, const float* __restrict__ rdx2,
                                                    const float* __restrict__ x, float* __restrict__ gx, int64_t B, int64_t D) {
    const int64_t n = B * D;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / D;
        gx[i] = g1[i] + (g2 ? g2[i] : 0.0f) + 2.0f * __ldg(rdx2 + b) * __ldg(x + i);
    }
}

// gp = gp1 + 2 dp2_j p + dpa_j a,  ga = ga1 + dpa_j p + dan_j a / |a_j|   (a == p: gp = gp1 + [2 (dp2 + dpa) + dan / |p|] p).
// One warp per plane.  a != p: the distance depends on a only through a / |a| (times |a| when SCALED), so <ga_j, a_j> is
// known in closed form - |a_j| * (sum_b g out0 when SCALED, else 0) - while the three terms above each carry O(B) along a_j
// and cancel to it; their rounding (2^-12 absolute at B = 1024) would otherwise land in the layer's `_bias` gradient, which
// reads ga along a.  The radial component is therefore replaced by its closed form.
__global__ void __launch_bounds__(256) k_gp_combine(const float* __restrict__ gp1, const float* __restrict__ ga1, const float* __restrict__ cd,
                                                    const float* __restrict__ p, const float* __restrict__ a, const float* __restrict__ an,
                                                    float* __restrict__ gp, float* __restrict__ ga, int64_t P, int64_t D, uint32_t flags) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t j = warp; j < P; j += nw) {
        const float dp2 = __ldg(cd + j), dpa = __ldg(cd + P + j), dan = __ldg(cd + 2 * P + j), rad = __ldg(cd + 3 * P + j), nj = __ldg(an + j);
        const float rn = nj > 0.0f ? 1.0f / nj : 0.0f;
        float dot = 0.0f;
        for (int64_t d = lane; d < D; d += 32) {
            const int64_t i = j * D + d;
            const float pv = __ldg(p + i);
            if (ga) {
                const float av = __ldg(a + i);
                gp[i] = gp1[i] + 2.0f * dp2 * pv + dpa * av;
                const float g = ga1[i] + dpa * pv + dan * rn * av;
                ga[i] = g;
                dot = fmaf(g, av, dot);
            } else {
                gp[i] = gp1[i] + (2.0f * (dp2 + dpa) + dan * rn) * pv;
            }
        }
        if (!ga) continue;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        const bool clamped = (flags & HVAE_GYRO_PVAE) && nj < kMinNorm;   // (|a| clamped: no dependence on it at all)
        const float want = clamped ? 0.0f : rad * nj;
        const float corr = (dot - want) * rn * rn;
        for (int64_t d = lane; d < D; d += 32) {
            const int64_t i = j * D + d;
            ga[i] -= corr * __ldg(a + i);
        }
    }
}

struct Ws {
    size_t x2, p2, pa, an, PX, XA, gemm, gemm_bytes, total;          // forward
    size_t g1, g2, gw1, gw2, rowpart, colpart, rd, cd, x3, x3_bytes;   // backward extras
};
static Ws layout(int64_t B, int64_t D, int64_t P, bool two, bool bwd) {
    Ws w{};
    size_t o = 0;
    auto take = [&](size_t n) { const size_t at = o; o += (n + 255) / 256 * 256; return at; };
    w.x2 = take((size_t)B * 4); w.p2 = take((size_t)P * 4); w.pa = take((size_t)P * 4); w.an = take((size_t)P * 4);
    w.PX = take((size_t)B * P * 4);
    w.XA = take(two ? (size_t)B * P * 4 : 0);
    const size_t gw = hvae_gemm_x3_workspace_bytes(B, P, D);
    w.gemm = take(gw);
    w.gemm_bytes = gw;
    if (bwd) {
        w.g1 = take((size_t)B * D * 4); w.g2 = take(two ? (size_t)B * D * 4 : 0);
        w.gw1 = take((size_t)P * D * 4); w.gw2 = take(two ? (size_t)P * D * 4 : 0);
        w.rowpart = take((size_t)((P + kCols - 1) / kCols) * B * 4);
        w.colpart = take((size_t)((B + kRows - 1) / kRows) * 4 * P * 4);
        w.rd = take((size_t)B * 4); w.cd = take((size_t)4 * P * 4);
        const size_t s1 = hvae_gemm_x3_workspace_bytes(B, D, P), s2 = hvae_gemm_x3_workspace_bytes(P, D, B);
        w.x3_bytes = s1 > s2 ? s1 : s2;
        w.x3 = take(w.x3_bytes);
    }
    w.total = o;
    return w;
}

static GyroParams make_params(float c, uint32_t flags) {
    const Ball bl = make_ball(c);
    GyroParams g;
    g.c = bl.c; g.sc = bl.sc; g.rsc = bl.rsc; g.maxnorm = bl.maxnorm; g.flags = flags;
    return g;
}
static unsigned grid_for(int64_t n) {
    const int64_t b = (n + 255) / 256;
    return (unsigned)(b < (int64_t)kNumSMs * 16 ? (b < 1 ? 1 : b) : (int64_t)kNumSMs * 16);
}

// PX (and XA) and the row / plane statistics, shared by forward and backward.  The inner products feed a pair function
// whose inner-product form amplifies their error on badly conditioned planes (|p| near the ball's edge), so they are formed
// with the 24-bit three-way split (2^-24 per term) rather than the two-piece fp16 GEMM (3 * 2^-22).
static int dots(const float* x, const float* p, const float* a, int64_t B, int64_t D, int64_t P, uint8_t* ws, const Ws& L, cudaStream_t s) {
    const bool two = a != p;
    k_rows_sq<<<grid_for(B * 32), 256, 0, s>>>(x, (float*)(ws + L.x2), B, D);
    k_plane_stats<<<grid_for(P * 32), 256, 0, s>>>(p, a, (float*)(ws + L.p2), (float*)(ws + L.pa), (float*)(ws + L.an), P, D);
    int rc = hvae_gemm_x3_f32(x, 0, p, 0, nullptr, 0, (float*)(ws + L.PX), B, P, D, ws + L.gemm, L.gemm_bytes, s);
    if (rc == HVAE_OK && two) rc = hvae_gemm_x3_f32(x, 0, a, 0, nullptr, 0, (float*)(ws + L.XA), B, P, D, ws + L.gemm, L.gemm_bytes, s);
    return rc;
}

}  // namespace gtc32
}  // namespace hvae

using namespace hvae;

extern "C" size_t hvae_gyroplane_tc32_fwd_workspace_bytes(int64_t B, int64_t D, int64_t P, int two) {
    if (B <= 0 || D <= 0 || P <= 0) return 0;
    return gtc32::layout(B, D, P, two != 0, false).total;
}
extern "C" size_t hvae_gyroplane_tc32_bwd_workspace_bytes(int64_t B, int64_t D, int64_t P, int two) {
    if (B <= 0 || D <= 0 || P <= 0) return 0;
    return gtc32::layout(B, D, P, two != 0, true).total;
}

extern "C" int hvae_gyroplane_tc32_fwd_f32(const float* x, const float* p, const float* a, const float* bias, float* out, int64_t B,
                                         int64_t D, int64_t P, float c, uint32_t flags, void* workspace, size_t workspace_bytes,
                                         void* stream) {
    if (B <= 0 || D <= 0 || P <= 0) return HVAE_ESHAPE;
    if (!x || !p || !out || !workspace) return HVAE_EARG;
    if (!a) a = p;
    const bool two = a != p;
    const gtc32::Ws L = gtc32::layout(B, D, P, two, false);
    if (workspace_bytes < L.total) return HVAE_EARG;
    cudaStream_t s = (cudaStream_t)stream;
    uint8_t* ws = (uint8_t*)workspace;
    int rc = gtc32::dots(x, p, a, B, D, P, ws, L, s);
    if (rc != HVAE_OK) return rc;
    gtc32::k_dots_fwd<<<gtc32::grid_for(B * P), 256, 0, s>>>((const float*)(ws + L.PX), two ? (const float*)(ws + L.XA) : nullptr,
                                                         (const float*)(ws + L.x2), (const float*)(ws + L.p2), (const float*)(ws + L.pa),
                                                         (const float*)(ws + L.an), bias, out, B, P, gtc32::make_params(c, flags));
    return check_launch();
}

// gp / ga: (P, D); ga must be NULL exactly when a aliases p (or is NULL); gx, gp may not be NULL.
extern "C" int hvae_gyroplane_tc32_bwd_f32(const float* x, const float* p, const float* a, const float* gout, float* gx, float* gp,
                                         float* ga, int64_t B, int64_t D, int64_t P, float c, uint32_t flags, void* workspace,
                                         size_t workspace_bytes, void* stream) {
    if (B <= 0 || D <= 0 || P <= 0) return HVAE_ESHAPE;
    if (!x || !p || !gout || !gx || !gp || !workspace) return HVAE_EARG;
    if (!a) a = p;
    const bool two = a != p;
    if (two != (ga != nullptr)) return HVAE_EARG;
    const gtc32::Ws L = gtc32::layout(B, D, P, two, true);
    if (workspace_bytes < L.total) return HVAE_EARG;
    cudaStream_t s = (cudaStream_t)stream;
    uint8_t* ws = (uint8_t*)workspace;
    int rc = gtc32::dots(x, p, a, B, D, P, ws, L, s);
    if (rc != HVAE_OK) return rc;
    float* PX = (float*)(ws + L.PX);
    float* XA = two ? (float*)(ws + L.XA) : nullptr;
    const int64_t rblk = (B + gtc32::kRows - 1) / gtc32::kRows, cblk = (P + gtc32::kCols - 1) / gtc32::kCols;
    {
        dim3 grid((unsigned)rblk, (unsigned)cblk);
        gtc32::k_dots_bwd<<<grid, gtc32::kCols, 0, s>>>(gout, PX, XA, (const float*)(ws + L.x2), (const float*)(ws + L.p2), (const float*)(ws + L.pa),
                                                    (const float*)(ws + L.an), (float*)(ws + L.rowpart), (float*)(ws + L.colpart), B, P,
                                                    gtc32::make_params(c, flags));
    }
    gtc32::k_sum_parts<<<gtc32::grid_for(B), 256, 0, s>>>((const float*)(ws + L.rowpart), (float*)(ws + L.rd), B, cblk);
    gtc32::k_sum_parts<<<gtc32::grid_for(4 * P), 256, 0, s>>>((const float*)(ws + L.colpart), (float*)(ws + L.cd), 4 * P, rblk);
    // The four gradient GEMMs run on the three-way bf16 split (2^-24 per term, tc_gemm.cu): the GeodesicLayer parameters'
    // gradients are cancelling sums (the distance does not depend on |a|, so <ga, a> vanishes) that magnify the per-term
    // error of the contraction; the two-piece fp16 products (3 * 2^-22 per term) left them 6x off the fp32 reference's own
    // error.  Operands are read as they lie: (B, P) . (P, D) and (B, P)^T . (B, D).
    for (int t = 0; t < (two ? 2 : 1); ++t) {
        const float* W = t ? a : p;
        const float* Cm = t ? XA : PX;
        float* g1 = (float*)(ws + (t ? L.g2 : L.g1));
        float* gw = (float*)(ws + (t ? L.gw2 : L.gw1));
        rc = hvae_gemm_x3_f32(Cm, 0, W, 1, nullptr, 0, g1, B, D, P, ws + L.x3, L.x3_bytes, s);      // contraction over the planes
        if (rc == HVAE_OK) rc = hvae_gemm_x3_f32(Cm, 1, x, 1, nullptr, 0, gw, P, D, B, ws + L.x3, L.x3_bytes, s);   // over the batch
        if (rc != HVAE_OK) return rc;
    }
    gtc32::k_gx_combine<<<gtc32::grid_for(B * D), 256, 0, s>>>((const float*)(ws + L.g1), two ? (const float*)(ws + L.g2) : nullptr,
                                                           (const float*)(ws + L.rd), x, gx, B, D);
    gtc32::k_gp_combine<<<gtc32::grid_for(P * 32), 256, 0, s>>>((const float*)(ws + L.gw1), two ? (const float*)(ws + L.gw2) : nullptr,
                                                            (const float*)(ws + L.cd), p, a, (const float*)(ws + L.an), gp, ga, P, D, flags);
    return check_launch();
}
