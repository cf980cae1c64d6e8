// K6 / K7: HyperbolicRadius — the radial part of the Riemannian normal on the Poincare ball — and
// expmap_polar.   reference: hyperbolic_vae/distributions/old_pvae_riemannian_normal.py:31,44-52 over pvae
// (hyperbolic_radius.py / ars.py / poincareball.py @ c04ec2149).
//
// density on r > 0:  rho(r) = exp(-r^2/(2 s^2)) (sinh(sqrt(c) r)/sqrt(c))^n / Z,  n = dim-1.
// With b_k = (n-2k) sqrt(c), C_k = binom(n,k), E_k = e^{b_k^2 s^2/2}(1 + erf(b_k s/sqrt2)):
//   Z = c^{-n/2} 2^{-n} s sqrt(pi/2) sum_k (-1)^k C_k E_k                      (K7, float64 inside)
//   F(r) = sum_k (-1)^k C_k e^{b_k^2 s^2/2}(erf((r-b_k s^2)/(s sqrt2)) + erf(b_k s/sqrt2)) / sum_k (-1)^k C_k E_k
// K6 samples r EXACTLY by rejection from a 3-tangent upper hull of the log-concave log-density (tangents at
// mode - w, mode, mode + w): a piecewise-exponential proposal like pvae's ARS (fixed hull, never refined),
// but with the knots placed by a Newton solve for the mode instead of the moment formulas, and Philox4x32-10
// counters instead of torch.rand — one thread per sample, no host loop.
// K6' is the implicit reparameterisation gradient dr/ds = -(dF/ds)/rho(r) in float64.
#include "hvae_common.cuh"

namespace hvae {

constexpr double kSqrt2 = 1.4142135623730950488;
constexpr double kSqrt2OverPi = 0.79788456080286535588;  // sqrt(2/pi)
constexpr int kMaxRadiusDim = 256;

// log(1 + erf(x)) without cancellation for x < 0
__device__ __forceinline__ double log1p_erf(double x) {
    if (x >= 0.0) return log1p(erf(x));
    return log(erfcx(-x)) - x * x;
}

__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One WARP per row: lane l evaluates the series terms k = l, l+32, ...  (the float64 erf/erfcx/log/exp chain of one term
// is ~1 us of latency; a thread-per-row loop over dim terms was pure latency: 16 us for B = 1).
// v_k = logC_k + b_k^2 s^2/2 + log(1+erf(b_k s/sqrt2));  logZ = const + log s + m + log sum_k (-1)^k e^{v_k - m}
__global__ void k_hradius_lognorm(const float* __restrict__ sigma, float* __restrict__ logZ, float* __restrict__ dlogZ,
                                  int64_t B, int dim, double c) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= B) return;  // warp-uniform
    const int n = dim - 1;
    const double s = (double)sigma[row];
    const double sc = sqrt(c);
    const double lgd = lgamma((double)dim);
    double m = -1e300;
    for (int k = lane; k <= n; k += 32) {
        const double b = (n - 2 * k) * sc;
        const double v = lgd - lgamma((double)k + 1.0) - lgamma((double)(dim - k)) + 0.5 * b * b * s * s + log1p_erf(b * s / kSqrt2);
        m = fmax(m, v);
    }
    m = warp_max_d(m);
    double S = 0.0, dS = 0.0;
    for (int k = lane; k <= n; k += 32) {
        const double b = (n - 2 * k) * sc;
        const double sg = (k & 1) ? -1.0 : 1.0;
        const double lc = lgd - lgamma((double)k + 1.0) - lgamma((double)(dim - k));
        const double v = lc + 0.5 * b * b * s * s + log1p_erf(b * s / kSqrt2);
        const double e = exp(v - m);
        S += sg * e;
        // d/ds [C e^{b^2 s^2/2}(1+erf(b s/sqrt2))] = b^2 s (.) + C b sqrt(2/pi)
        dS += sg * (b * b * s * e + exp(lc - m) * b * kSqrt2OverPi);
    }
    S = warp_sum_d(S);
    dS = warp_sum_d(dS);
    if (lane == 0) {
        const double lz = 0.5 * (log(3.14159265358979323846) - log(2.0)) + log(s) - n * (0.5 * log(c) + log(2.0)) + m + log(S);
        logZ[row] = (float)lz;
        if (dlogZ) dlogZ[row] = (float)(1.0 / s + dS / S);
    }
}

// ---- Philox4x32-10 ------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}

__device__ __forceinline__ float u01(uint32_t x) {  // (0, 1]
    return ((float)(x >> 8) + 1.0f) * (1.0f / 16777216.0f);
}

// ---- HypersphericalUniform.sample in the kernel (pvae: normalise a standard normal vector; App. A.2) -------------------
// Row i (sample s of batch row b, i = s * B + b) draws its D normals from Philox counters (cnt_i, chunk, 0xa1fa) - a
// stream disjoint from the radius sampler's (.., iter, 0x5ad1) - so a row's direction depends only on (seed, cnt_i):
// a data-parallel shard reproduces the rows a single-GPU run would draw.  One thread per row.
__global__ void k_sphere_sample(float* __restrict__ out, int64_t rows, int D, uint64_t seed, uint64_t offset,
                                const int64_t* __restrict__ offset_dev) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    const uint64_t cnt = offset + (offset_dev ? (uint64_t)__ldg(offset_dev) : 0ull) + (uint64_t)i;
    float* o = out + i * D;
    float n2 = 0.0f;
    for (int d0 = 0; d0 < D; d0 += 4) {
        const uint4 rn = philox4x32_10(make_uint4((uint32_t)cnt, (uint32_t)(cnt >> 32), (uint32_t)(d0 >> 2), 0xa1fau), key);
        // two Box-Muller pairs
        const float r0 = sqrtf(-2.0f * logf(u01(rn.x))), r1 = sqrtf(-2.0f * logf(u01(rn.z)));
        float s0, c0, s1, c1;
        sincospif(2.0f * u01(rn.y), &s0, &c0);
        sincospif(2.0f * u01(rn.w), &s1, &c1);
        const float z[4] = {r0 * c0, r0 * s0, r1 * c1, r1 * s1};
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (d0 + e < D) { o[d0 + e] = z[e]; n2 = fmaf(z[e], z[e], n2); }
    }
    const float rn_ = rsqrtf(fmaxf(n2, 1e-30f));
    for (int d = 0; d < D; ++d) o[d] *= rn_;
}

// log-density (unnormalised) and derivative, float32 is enough for the accept test
__device__ __forceinline__ float rad_h(float r, float inv2s2, float sc, float n) {
    const float x = sc * r;
    // log sinh x = x + log(1 - e^{-2x}) - log 2 ; for tiny x use log x
    const float ls = (x < 1e-3f) ? logf(x) : x + log1pf(-expf(-2.0f * x)) - 0.69314718f;
    return -r * r * inv2s2 + n * ls;
}
__device__ __forceinline__ float rad_hp(float r, float inv2s2, float sc, float n) {
    const float x = sc * r;
    const float coth = (x < 1e-3f) ? 1.0f / x + x * (1.0f / 3.0f) : (1.0f + expf(-2.0f * x)) / (1.0f - expf(-2.0f * x));
    return -2.0f * r * inv2s2 + n * sc * coth;
}
__device__ __forceinline__ float rad_hpp(float r, float inv2s2, float sc, float n) {
    const float x = sc * r;
    const float sh = (x < 1e-3f) ? x : sinhf(fminf(x, 40.0f));
    return -2.0f * inv2s2 - n * sc * sc / (sh * sh);
}

__global__ void k_hradius_sample(const float* __restrict__ sigma, float* __restrict__ r_out, int64_t S, int64_t B, int dim,
                                 float c, uint64_t seed, uint64_t offset, const int64_t* __restrict__ offset_dev) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S * B) return;
    const float s = sigma[i % B];
    const float sc = sqrtf(c);
    const float n = (float)(dim - 1);
    const float inv2s2 = 0.5f / (s * s);
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    const uint64_t cnt = offset + (offset_dev ? (uint64_t)__ldg(offset_dev) : 0ull) + (uint64_t)i;
    uint32_t iter = 0;
    if (dim == 1) {  // half normal
        const uint4 rn = philox4x32_10(make_uint4((uint32_t)cnt, (uint32_t)(cnt >> 32), 0u, 0u), key);
        const float u1 = u01(rn.x), u2 = u01(rn.y);
        r_out[i] = s * fabsf(sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2));
        return;
    }
    // mode: h'(r) = 0, Newton from the large-x asymptote n s^2 sqrt(c) / small-x sqrt(n) s
    float mode = fmaxf(n * s * s * sc, sqrtf(n) * s);
#pragma unroll 1
    for (int it = 0; it < 30; ++it) {
        const float step = rad_hp(mode, inv2s2, sc, n) / rad_hpp(mode, inv2s2, sc, n);
        float nm = mode - step;
        if (!(nm > 0.0f)) nm = 0.5f * mode;
        const bool done = fabsf(nm - mode) <= 1e-6f * mode;
        mode = nm;
        if (done) break;
    }
    const float w = rsqrtf(-rad_hpp(mode, inv2s2, sc, n));  // local std
    // three tangents: x0 < x1 = mode < x2
    float x0 = mode - w;
    if (x0 < 0.25f * mode) x0 = 0.25f * mode;
    const float x1 = mode, x2 = mode + w;
    const float hoff = rad_h(x1, inv2s2, sc, n);
    const float h0 = rad_h(x0, inv2s2, sc, n) - hoff, h2 = rad_h(x2, inv2s2, sc, n) - hoff;  // h1 = 0
    const float d0 = rad_hp(x0, inv2s2, sc, n), d2 = rad_hp(x2, inv2s2, sc, n);             // d0 > 0 > d2; d1 = 0
    // intersections of consecutive tangents: z1 between x0,x1 ; z2 between x1,x2
    const float z1 = x0 - h0 / d0;   // h0 + d0 (z - x0) = 0
    const float z2 = x2 - h2 / d2;   // h2 + d2 (z - x2) = 0
    // segment masses: [0,z1] slope d0 ; [z1,z2] flat at 0 ; [z2,inf) slope d2
    const float m0 = (1.0f - expf(h0 + d0 * (0.0f - x0))) / d0;   // (e^{u(z1)} - e^{u(0)})/d0, u(z1) = 0
    const float m1 = z2 - z1;
    const float m2 = -1.0f / d2;                                    // int_{z2}^inf e^{d2 (r - z2)} dr
    const float tot = m0 + m1 + m2;
    float r = mode;
#pragma unroll 1
    for (; iter < 1000u; ++iter) {
        const uint4 rn = philox4x32_10(make_uint4((uint32_t)cnt, (uint32_t)(cnt >> 32), iter, 0x5ad1u), key);
        const float u = u01(rn.x) * tot;
        float env;
        if (u < m0) {
            // invert  (e^{d0 (r - z1)} - e^{-d0 z1})/d0 = u
            r = z1 + logf(fmaf(u, d0, expf(-d0 * z1))) / d0;
            env = d0 * (r - z1);
        } else if (u < m0 + m1) {
            r = z1 + (u - m0);
            env = 0.0f;
        } else {
            const float v = (u - m0 - m1) / m2;       // in [0,1)
            r = z2 + logf(fmaxf(1.0f - v, 1e-30f)) / d2;
            env = d2 * (r - z2);
        }
        if (!(r > 0.0f)) continue;
        const float lh = rad_h(r, inv2s2, sc, n) - hoff;
        if (logf(u01(rn.y)) <= lh - env) break;
    }
    r_out[i] = r;
}

// dr/dsigma by implicit differentiation of F(r; sigma) = U.  float64, one warp per sample (lanes stride over terms).
__global__ void k_hradius_rgrad(const float* __restrict__ sigma, const float* __restrict__ r_in, float* __restrict__ dr_ds,
                                float* __restrict__ cdf, int64_t S, int64_t B, int dim, double c) {
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= S * B) return;  // warp-uniform
    const int n = dim - 1;
    const double s = (double)sigma[i % B];
    const double r = (double)r_in[i];
    const double sc = sqrt(c);
    const double lgd = lgamma((double)dim);
    double m = -1e300;
    for (int k = lane; k <= n; k += 32) {
        const double b = (n - 2 * k) * sc;
        m = fmax(m, lgd - lgamma((double)k + 1.0) - lgamma((double)(dim - k)) + 0.5 * b * b * s * s);
    }
    m = warp_max_d(m);
    double num = 0.0, den = 0.0, dnum = 0.0, dden = 0.0, dfr = 0.0;
    const double gauss_r = -r * r / (2.0 * s * s);
    for (int k = lane; k <= n; k += 32) {
        const double b = (n - 2 * k) * sc;
        const double sg = (k & 1) ? -1.0 : 1.0;
        const double lc = lgd - lgamma((double)k + 1.0) - lgamma((double)(dim - k));
        const double w = exp(lc + 0.5 * b * b * s * s - m);   // C e^{b^2 s^2/2}, scaled
        const double w0 = exp(lc - m);                        // C, scaled
        const double eA = erf((r - b * s * s) / (s * kSqrt2));
        const double eB = erf(b * s / kSqrt2);
        num += sg * w * (eA + eB);
        den += sg * w * (1.0 + eB);
        // d/ds: b^2 s w (.) + sqrt(2/pi) C [(-r/s^2 - b) e^{-r^2/2s^2 + r b} + b]
        const double ex = exp(lc + gauss_r + r * b - m);
        dnum += sg * (b * b * s * w * (eA + eB) + kSqrt2OverPi * ((-r / (s * s) - b) * ex + b * w0));
        dden += sg * (b * b * s * w * (1.0 + eB) + kSqrt2OverPi * b * w0);
        dfr += sg * ex;   // rho(r) numerator: sum_k sg C e^{-r^2/2s^2 + r b}
    }
    num = warp_sum_d(num); den = warp_sum_d(den); dnum = warp_sum_d(dnum); dden = warp_sum_d(dden); dfr = warp_sum_d(dfr);
    if (lane == 0) {
        const double F = num / den;
        const double dF_ds = (dnum - F * dden) / den;
        const double dF_dr = dfr * kSqrt2OverPi / s / den;
        dr_ds[i] = (float)(-dF_ds / dF_dr);
        if (cdf) cdf[i] = (float)F;
    }
}

// ---- expmap_polar: z = project(mu (+) tanh(sc r/2) alpha/(sc |alpha|)) ----------------------------------------
template <int G, int EPL>
__global__ void __launch_bounds__(kRowThreads)
k_expmap_polar_fwd(const float* __restrict__ mu, const float* __restrict__ alpha, const float* __restrict__ r,
                   float* __restrict__ z, int64_t S, int64_t B, int D, Ball ball) {
    HVAE_ROW_PROLOGUE(G)
    const int64_t rows = S * B;
    for (int64_t r0 = warp_global * RPW; r0 < rows; r0 += warps_total * RPW) {
        const int64_t row = r0 + sub;
        const bool valid = row < rows;
        RowSlice<G, EPL> m, a, w, o;
        m.load(mu, valid ? row % B : 0, D, lg, valid);
        a.load(alpha, row, D, lg, valid);
        const float rr = valid ? __ldg(r + row) : 0.0f;
        const float an = fmaxf(sqrtf(sqnorm<G, EPL>(a)), kMinNorm);
        const float q = tanh_c(ball.sc * 0.5f * rr) / (ball.sc * an);
#pragma unroll
        for (int i = 0; i < EPL; ++i) w.v[i] = q * a.v[i];
        mobius_add_raw<G, EPL>(m, w, o, ball);
        float pn;
        project_inplace<G, EPL>(o, ball, pn);
        o.store(z, row, D, lg, valid);
    }
}

template <int G, int EPL>
__global__ void __launch_bounds__(kRowThreads)
k_expmap_polar_bwd(const float* __restrict__ mu, const float* __restrict__ alpha, const float* __restrict__ r,
                   const float* __restrict__ gz, float* __restrict__ gmu, float* __restrict__ gr, int64_t S, int64_t B, int D,
                   Ball ball) {
    HVAE_ROW_PROLOGUE(G)
    for (int64_t r0 = warp_global * RPW; r0 < B; r0 += warps_total * RPW) {
        const int64_t b = r0 + sub;
        const bool valid = b < B;
        RowSlice<G, EPL> m, gm;
        m.load(mu, b, D, lg, valid);
        gm.zero();
        for (int64_t s = 0; s < S; ++s) {
            const int64_t row = s * B + b;
            RowSlice<G, EPL> a, w, o, g, gx, gw;
            a.load(alpha, row, D, lg, valid);
            g.load(gz, row, D, lg, valid);
            const float rr = valid ? __ldg(r + row) : 0.0f;
            const float an = fmaxf(sqrtf(sqnorm<G, EPL>(a)), kMinNorm);
            const float th = ball.sc * 0.5f * rr;
            const float t = tanh_c(th);
            const float q = t / (ball.sc * an);
#pragma unroll
            for (int i = 0; i < EPL; ++i) w.v[i] = q * a.v[i];
            const MAddCtx ma = mobius_add_raw<G, EPL>(m, w, o, ball);
            RowSlice<G, EPL> op = o;
            float pn;
            const bool hit = project_inplace<G, EPL>(op, ball, pn);
            project_bwd<G, EPL>(g, o, pn, hit, ball);
            mobius_add_raw_bwd<G, EPL>(m, w, ma, g, gx, gw, ball);
#pragma unroll
            for (int i = 0; i < EPL; ++i) gm.v[i] += gx.v[i];
            // w = tanh(sc r/2)/(sc |a|) a :  dw/dr = sech^2(sc r/2)/2 * a/|a|
            const float gwa = dot<G, EPL>(gw, a);
            if (valid && lg == 0 && gr) gr[row] = gwa * (1.0f - t * t) * tanh_mask(th) * 0.5f / an;
        }
        if (gmu) gm.store(gmu, b, D, lg, valid);
    }
}

}  // namespace hvae

using namespace hvae;

extern "C" int hvae_hradius_lognorm_fwd_f32(const float* sigma, float* logZ, float* dlogZ_dsigma, int64_t B, int64_t dim,
                                            float c, void* stream) {
    if (B < 0 || dim < 1 || dim > kMaxRadiusDim) return HVAE_ESHAPE;
    if (B == 0) return HVAE_OK;
    if (!sigma || !logZ) return HVAE_EARG;
    k_hradius_lognorm<<<(unsigned)((B + 3) / 4), 128, 0, (cudaStream_t)stream>>>(sigma, logZ, dlogZ_dsigma, B, (int)dim,
                                                                                     (double)c);
    return check_launch();
}

extern "C" int hvae_hradius_sample_f32(const float* sigma, float* r, int64_t S, int64_t B, int64_t dim, float c,
                                       uint64_t seed, uint64_t offset, const int64_t* offset_dev, void* stream) {
    if (S < 0 || B < 0 || dim < 1 || dim > kMaxRadiusDim) return HVAE_ESHAPE;
    if (S == 0 || B == 0) return HVAE_OK;
    if (!sigma || !r) return HVAE_EARG;
    const int64_t n = S * B;
    k_hradius_sample<<<(unsigned)((n + 63) / 64), 64, 0, (cudaStream_t)stream>>>(sigma, r, S, B, (int)dim, c, seed, offset,
                                                                                     offset_dev);
    return check_launch();
}

extern "C" int hvae_hradius_rgrad_f32(const float* sigma, const float* r, float* dr_dsigma, float* cdf, int64_t S, int64_t B,
                                      int64_t dim, float c, void* stream) {
    if (S < 0 || B < 0 || dim < 1 || dim > kMaxRadiusDim) return HVAE_ESHAPE;
    if (S == 0 || B == 0) return HVAE_OK;
    if (!sigma || !r || !dr_dsigma) return HVAE_EARG;
    const int64_t n = S * B;
    k_hradius_rgrad<<<(unsigned)((n + 3) / 4), 128, 0, (cudaStream_t)stream>>>(sigma, r, dr_dsigma, cdf, S, B, (int)dim,
                                                                                   (double)c);
    return check_launch();
}

extern "C" int hvae_expmap_polar_fwd_f32(const float* mu, const float* alpha, const float* r, float* z, int64_t S, int64_t B,
                                         int64_t D, float c, void* stream) {
    if (S < 0 || B < 0 || D <= 0 || D > kMaxRowDim) return HVAE_ESHAPE;
    if (S == 0 || B == 0) return HVAE_OK;
    if (!mu || !alpha || !r || !z) return HVAE_EARG;
    HVAE_ROW_DISPATCH(k_expmap_polar_fwd, D, S * B, (cudaStream_t)stream, mu, alpha, r, z, S, B, (int)D, make_ball(c));
    return check_launch();
}

extern "C" int hvae_expmap_polar_bwd_f32(const float* mu, const float* alpha, const float* r, const float* gz, float* gmu,
                                         float* gr, int64_t S, int64_t B, int64_t D, float c, void* stream) {
    if (S < 0 || B < 0 || D <= 0 || D > kMaxRowDim) return HVAE_ESHAPE;
    if (S == 0 || B == 0) return HVAE_OK;
    if (!mu || !alpha || !r || !gz) return HVAE_EARG;
    HVAE_ROW_DISPATCH(k_expmap_polar_bwd, D, B, (cudaStream_t)stream, mu, alpha, r, gz, gmu, gr, S, B, (int)D, make_ball(c));
    return check_launch();
}

// alpha (rows, D) ~ U(S^{D-1}): normalised standard normals, Philox4x32-10 in the kernel.  Counters as for
// hvae_hradius_sample_f32: row i uses offset + *offset_dev + i (offset_dev may be NULL).
extern "C" int hvae_sphere_sample_f32(float* out, int64_t rows, int64_t D, uint64_t seed, uint64_t offset,
                                      const int64_t* offset_dev, void* stream) {
    if (rows <= 0 || D <= 0 || D > 65536) return HVAE_ESHAPE;
    if (!out) return HVAE_EARG;
    k_sphere_sample<<<(unsigned)((rows + 127) / 128), 128, 0, (cudaStream_t)stream>>>(out, rows, (int)D, seed, offset, offset_dev);
    return check_launch();
}
