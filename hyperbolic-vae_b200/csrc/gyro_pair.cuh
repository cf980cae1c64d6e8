// Scalar epilogue of the gyroplane (hyperplane-distance) op, shared by the SIMT kernels (gyroplane.cu) and the
// tcgen05 GEMM epilogue (tc_gemm.cu).  reference: geoopt math.dist2plane (App. A.1) and
// hyperbolic_vae/manifolds.py:41-65 (normdist2plane) behind HVAE_GYRO_PVAE.
#pragma once
#include "hvae_common.cuh"

namespace hvae {

struct GyroParams {
    float c, sc, rsc, maxnorm;
    uint32_t flags;
};

// asinh with branch-free fast intrinsics: |err| <= ~2e-6 relative (the parity budget is 1e-5).
//   |y| <  0.25 : odd Taylor polynomial to y^9 (next term 2e-8 relative at 0.25)
//   |y| >= 0.25 : log(|y| + sqrt(y^2+1)) via lg2.approx (abs err 2^-21.4 on a value >= 0.247)
__device__ __forceinline__ float asinh_fast(float y) {
    const float ay = fabsf(y);
    const float y2 = ay * ay;
    const float poly = ay * fmaf(y2, fmaf(y2, fmaf(y2, fmaf(y2, 105.0f / 3456.0f, -15.0f / 336.0f), 3.0f / 40.0f), -1.0f / 6.0f), 1.0f);
    const float h = fmaf(ay, ay, 1.0f);
    const float big = __logf(ay + h * rsqrtf(h));
    return copysignf(ay < 0.25f ? poly : big, y);
}

__device__ __forceinline__ float rcp_fast(float v) { return __fdividef(1.0f, v); }

struct GyroPairCtx {
    float A, Bc, den, N1, N2, da, dn2, an, w, denom, y, out0, out1, rho_n;  // rho_n: |diff| when projected
    bool den_ok, dn2_ok, w_ok, projected;
};

// scalar epilogue shared by forward and backward (and, later, by the tensor-core path)
// Difference form (SIMT path): with e = |x-p|^2, q = <p,p-x>, qa = <a,p-x> accumulated from elementwise
// differences (exact when x -> p), A = Bc + c e and
//   N1 = -A<p,a> + Bc<x,a> = -Bc qa - c e <p,a>,   N2 = e (Bc^2 + 2 Bc c q + c^2 e |p|^2)
// so the x -> p cancellation that the plain inner-product form suffers (abs error eps*|p|^2) is gone.
struct GyroDiff {
    float e, q, qa;
};

__device__ __forceinline__ float gyro_pair_fwd(float px, float xa, float x2, float p2, float pa, float an_raw,
                                               const GyroParams& P, GyroPairCtx& k, const GyroDiff* df = nullptr) {
    const float c = P.c;
    const bool pvae = P.flags & HVAE_GYRO_PVAE;
    k.Bc = 1.0f - c * p2;
    const float den0 = 1.0f - 2.0f * c * px + c * c * p2 * x2;
    k.den_ok = den0 >= kMinNorm;
    k.den = fmaxf(den0, kMinNorm);
    if (df) {
        k.A = k.Bc + c * df->e;
        k.N1 = -k.Bc * df->qa - c * df->e * pa;
        k.N2 = fmaxf(df->e * (k.Bc * k.Bc + 2.0f * k.Bc * c * df->q + c * c * df->e * p2), 0.0f);
    } else {
        k.A = 1.0f - 2.0f * c * px + c * x2;
        k.N1 = -k.A * pa + k.Bc * xa;
        k.N2 = fmaxf(k.A * k.A * p2 - 2.0f * k.A * k.Bc * px + k.Bc * k.Bc * x2, 0.0f);
    }
    const float rden = 1.0f / k.den;   // IEEE reciprocals on the forward value path: the approximate ones (2 ulp each) put a
    k.da = k.N1 * rden;                // handful of well-conditioned outputs at 1.1-2.2e-5 of the reference (strict audit)
    float dn2r = k.N2 * rden * rden;
    k.projected = false;
    if (pvae) {
        const float n = fmaxf(sqrtf(dn2r), kMinNorm);
        if (n > P.maxnorm) {  // (-p)(+)x was projected back into the ball
            k.projected = true;
            k.rho_n = n;
            k.da = k.da / n * P.maxnorm;
            dn2r = P.maxnorm * P.maxnorm;
        }
    }
    k.dn2_ok = dn2r >= kMinNorm;
    k.dn2 = fmaxf(dn2r, kMinNorm);
    const float s = (P.flags & HVAE_GYRO_SIGNED) ? k.da : fabsf(k.da);
    k.an = pvae ? fmaxf(an_raw, kMinNorm) : an_raw;
    k.w = (1.0f - c * k.dn2) * k.an;
    if (pvae) {
        k.w_ok = k.w >= kMinNorm;
        k.denom = fmaxf(k.w, kMinNorm);
    } else {
        k.w_ok = true;
        k.denom = (k.w >= 0.0f ? 1.0f : -1.0f) * (fabsf(k.w) + kMinNorm);  // clamp_abs, sign(0) = +1
    }
    k.y = 2.0f * P.sc * s / k.denom;
    k.out0 = asinh_fast(k.y) * P.rsc;
    k.out1 = (P.flags & HVAE_GYRO_SCALED) ? k.out0 * k.an : k.out0;
    float o = k.out1;
    if (P.flags & HVAE_GYRO_SQUARED) {
        const float sg = (o > 0.0f) ? 1.0f : ((o < 0.0f) ? -1.0f : 0.0f);
        o = (P.flags & HVAE_GYRO_SIGNED) ? o * o * sg : o * o;
    }
    return o;
}

struct GyroPairGrad {
    float dpx, dxa, dx2, dp2, dpa, dan;  // dan: gradient wrt the RAW ||a||
};

__device__ __forceinline__ GyroPairGrad gyro_pair_bwd(float g, float px, float xa, float x2, float p2, float pa,
                                                      float an_raw, const GyroParams& P, const GyroPairCtx& k) {
    GyroPairGrad r;
    const float c = P.c;
    const bool pvae = P.flags & HVAE_GYRO_PVAE;
    float g1 = g;
    if (P.flags & HVAE_GYRO_SQUARED) g1 = (P.flags & HVAE_GYRO_SIGNED) ? g * 2.0f * fabsf(k.out1) : g * 2.0f * k.out1;
    float gan = 0.0f;  // wrt the (possibly clamped) an
    float g0 = g1;
    if (P.flags & HVAE_GYRO_SCALED) {
        gan += g1 * k.out0;
        g0 = g1 * k.an;
    }
    const float dy = g0 * P.rsc * rsqrtf(fmaf(k.y, k.y, 1.0f));
    const float rdenom = rcp_fast(k.denom);
    const float ds = dy * 2.0f * P.sc * rdenom;
    const float ddenom = -dy * k.y * rdenom;
    const float dw = k.w_ok ? ddenom : 0.0f;
    const float ddn2 = dw * (-c * k.an);
    gan += dw * (1.0f - c * k.dn2);
    float dda = ds;
    if (!(P.flags & HVAE_GYRO_SIGNED)) dda = (k.da > 0.0f) ? ds : ((k.da < 0.0f) ? -ds : 0.0f);
    float dN1, dN2, dden;
    if (k.projected) {
        // da = maxnorm * N1 / sqrt(N2); |diff|^2 == maxnorm^2 carries no gradient
        const float rs = rsqrtf(k.N2);
        dN1 = dda * P.maxnorm * rs;
        dN2 = -dda * P.maxnorm * k.N1 * 0.5f * rs * rs * rs;
        dden = 0.0f;
    } else {
        const float ddn2r = k.dn2_ok ? ddn2 : 0.0f;
        const float rden = rcp_fast(k.den);
        dN1 = dda * rden;
        dN2 = ddn2r * rden * rden;
        dden = -dda * k.N1 * rden * rden - 2.0f * ddn2r * k.N2 * rden * rden * rden;
    }
    const float dden0 = k.den_ok ? dden : 0.0f;
    const float dA = -pa * dN1 + (2.0f * k.A * p2 - 2.0f * k.Bc * px) * dN2;
    const float dBc = xa * dN1 + (-2.0f * k.A * px + 2.0f * k.Bc * x2) * dN2;
    r.dpa = -k.A * dN1;
    r.dxa = k.Bc * dN1;
    r.dp2 = k.A * k.A * dN2 - c * dBc + c * c * x2 * dden0;
    r.dx2 = k.Bc * k.Bc * dN2 + c * dA + c * c * p2 * dden0;
    r.dpx = -2.0f * k.A * k.Bc * dN2 - 2.0f * c * dA - 2.0f * c * dden0;
    r.dan = (pvae && an_raw < kMinNorm) ? 0.0f : gan;
    return r;
}

}  // namespace hvae
