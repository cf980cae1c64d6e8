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
    float g_used;                        // set by gyro_pair_grad_general only: the upstream gradient after the fused-ReLU mask
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


// ---------------------------------------------------------------------------------------------------
// Lean pair path (SIMT kernels).  Two closed forms cover every pair on which no MIN_NORM clamp binds:
//  * unprojected: da = N1/den, |diff|^2 = N2/den^2, w = (1 - c|diff|^2)|a|, y = 2 sc da / w collapse to ONE division,
//        y = 2 sc N1 den / (|a| (den^2 - c N2));
//  * projected (pvae: |(-p)(+)x| > maxnorm, the COMMON case for a 10-d RiemannianNormal posterior, whose radius
//    concentrates near (D-1) sigma^2 - far beyond the fp32 projection radius): diff is rescaled to norm maxnorm, so
//    |diff|^2 and w are constants and  y = K_j N1 / sqrt(N2),  K_j = 2 sc maxnorm / ((1 - c maxnorm^2)|a_j|)  (den cancels).
// Reciprocals are MUFU + one Newton step (~1 ulp, no IEEE slow-path branch); the backward differentiates these forms
// directly.  gyro_pair_lean_fwd returns GYRO_GENERAL when one of the reference's clamps (den, |diff|^2, w >= MIN_NORM;
// geoopt's clamp_abs) would change the value or mask a gradient: the caller then takes gyro_pair_fwd / _bwd.
// ---------------------------------------------------------------------------------------------------
struct GyroPlaneK {
    float p2, pa, an_raw, an, ran, Bc, kproj;   // |p|^2, <p,a>, |a|, clamped |a| (pvae), 1/an, 1 - c|p|^2, K_j
};

__device__ __forceinline__ GyroPlaneK gyro_plane_consts(float p2, float pa, float an_raw, const GyroParams& P) {
    GyroPlaneK k;
    k.p2 = p2; k.pa = pa; k.an_raw = an_raw;
    k.an = (P.flags & HVAE_GYRO_PVAE) ? fmaxf(an_raw, kMinNorm) : an_raw;
    k.ran = 1.0f / k.an;
    k.Bc = 1.0f - P.c * p2;
    const float w = (1.0f - P.c * (P.maxnorm * P.maxnorm)) * k.an;      // w of a projected pair
    k.kproj = (w >= kMinNorm) ? 2.0f * P.sc * P.maxnorm / w : 0.0f;     // 0: the w clamp binds -> general path
    return k;
}

// MUFU wrappers without the denormal pre-/post-scaling the plain intrinsics carry in non-ftz builds: every argument on
// the lean path is >= 1 (1 + y^2, |y| + sqrt(1 + y^2)) or a checked normal number
__device__ __forceinline__ float rsqrt_ftz(float v) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float lg2_ftz(float v) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
// asinh_fast with those wrappers (same polynomial / log split, same error bound)
__device__ __forceinline__ float asinh_lean(float y, float h /* 1 + y^2 */, float rsh /* rsqrt(h) */) {
    const float ay = fabsf(y);
    const float y2 = ay * ay;
    const float poly = ay * fmaf(y2, fmaf(y2, fmaf(y2, fmaf(y2, 105.0f / 3456.0f, -15.0f / 336.0f), 3.0f / 40.0f), -1.0f / 6.0f), 1.0f);
    const float big = lg2_ftz(fmaf(h, rsh, ay)) * 0.693147180559945309f;
    return copysignf(ay < 0.25f ? poly : big, y);
}

__device__ __forceinline__ float rcp_nr(float v) {   // v normal and positive
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return fmaf(r, fmaf(-v, r, 1.0f), r);
}

enum : int { GYRO_GENERAL = 0, GYRO_UNPROJECTED = 1, GYRO_PROJECTED = 2 };

struct GyroLean {
    // y = s * t:  unprojected s = +-2 sc N1 den, t = 1/(|a| (den^2 - c N2));  projected s = +-N1, t = K_j rsqrt(N2)
    float A, N1, N2, den, t, rs, y, rsh, out0, out1;   // rs = rsqrt(N2) (projected), rsh = rsqrt(1 + y^2)
};

__device__ __forceinline__ int gyro_pair_lean_fwd(const GyroDiff& df, float x2, const GyroPlaneK& pl, const GyroParams& P,
                                                  GyroLean& L, float& out) {
    const float c = P.c;
    const bool pvae = P.flags & HVAE_GYRO_PVAE;
    const float px = pl.p2 - df.q;
    const float den = fmaf(c * c * pl.p2, x2, fmaf(-2.0f * c, px, 1.0f));
    const float ce = c * df.e;
    L.A = pl.Bc + ce;
    L.N1 = -fmaf(pl.Bc, df.qa, ce * pl.pa);
    L.N2 = df.e * fmaf(c * pl.p2, ce, fmaf(2.0f * c * pl.Bc, df.q, pl.Bc * pl.Bc));
    L.den = den;
    const float dd = den * den;
    const float lo = kMinNorm * dd;
    if (!(den >= kMinNorm) || !(L.N2 >= lo)) return GYRO_GENERAL;
    int mode;
    float s;
    if (pvae && L.N2 > P.maxnorm * P.maxnorm * dd) {
        if (!(pl.kproj > 0.0f)) return GYRO_GENERAL;
        mode = GYRO_PROJECTED;
        L.rs = rsqrt_ftz(L.N2);
        L.t = pl.kproj * L.rs;
        s = L.N1;
    } else {
        // w = denom/dd above its clamp (geoopt: |w| + MIN_NORM == |w| in fp32)
        const float denom = pl.an * fmaf(-c, L.N2, dd);
        if (!(denom >= (pvae ? lo : 1e-6f * dd))) return GYRO_GENERAL;
        mode = GYRO_UNPROJECTED;
        L.t = rcp_nr(denom);
        s = (2.0f * P.sc) * L.N1 * den;
    }
    L.y = ((P.flags & HVAE_GYRO_SIGNED) ? s : fabsf(s)) * L.t;
    const float h = fmaf(L.y, L.y, 1.0f);
    L.rsh = rsqrt_ftz(h);
    L.out0 = asinh_lean(L.y, h, L.rsh) * P.rsc;
    L.out1 = (P.flags & HVAE_GYRO_SCALED) ? L.out0 * pl.an : L.out0;
    float o = L.out1;
    if (P.flags & HVAE_GYRO_SQUARED) {
        const float sg = (o > 0.0f) ? 1.0f : ((o < 0.0f) ? -1.0f : 0.0f);
        o = (P.flags & HVAE_GYRO_SIGNED) ? o * o * sg : o * o;
    }
    out = o;
    return mode;
}

__device__ __forceinline__ GyroPairGrad gyro_pair_lean_bwd(int mode, float g, const GyroDiff& df, float x2, const GyroPlaneK& pl,
                                                           const GyroParams& P, const GyroLean& L) {
    GyroPairGrad r;
    const float c = P.c;
    float g1 = g;
    if (P.flags & HVAE_GYRO_SQUARED) g1 = (P.flags & HVAE_GYRO_SIGNED) ? g * 2.0f * fabsf(L.out1) : g * 2.0f * L.out1;
    float gan = 0.0f, g0 = g1;
    if (P.flags & HVAE_GYRO_SCALED) {
        gan = g1 * L.out0;
        g0 = g1 * pl.an;
    }
    const float dy = g0 * P.rsc * L.rsh;
    float u = dy * L.t;                                       // dL/ds up to the sign of s
    if (!(P.flags & HVAE_GYRO_SIGNED)) u = (L.N1 > 0.0f) ? u : ((L.N1 < 0.0f) ? -u : 0.0f);   // den > 0: sign(s) = sign(N1)
    const float yv = dy * L.y;
    gan = fmaf(-yv, pl.ran, gan);
    float dN1, dN2, dden;
    if (mode == GYRO_PROJECTED) {
        dN1 = u;
        dN2 = -0.5f * yv * L.rs * L.rs;                       // y ~ N2^(-1/2)
        dden = 0.0f;
    } else {
        u *= 2.0f * P.sc;
        const float v = yv * L.t * pl.an;                     // -dL/d(den^2 - c N2)
        dN1 = u * L.den;
        dN2 = c * v;
        dden = fmaf(u, L.N1, -2.0f * L.den * v);
    }
    const float px = pl.p2 - df.q, xa = pl.pa - df.qa;
    const float A = L.A, Bc = pl.Bc;
    const float dA = fmaf(-pl.pa, dN1, 2.0f * (A * pl.p2 - Bc * px) * dN2);
    const float dBc = fmaf(xa, dN1, 2.0f * (Bc * x2 - A * px) * dN2);
    r.dpa = -A * dN1;
    r.dxa = Bc * dN1;
    r.dp2 = fmaf(A * A, dN2, fmaf(c * c * x2, dden, -c * dBc));
    r.dx2 = fmaf(Bc * Bc, dN2, fmaf(c * c * pl.p2, dden, c * dA));
    r.dpx = -2.0f * fmaf(A * Bc, dN2, c * (dA + dden));
    r.dan = ((P.flags & HVAE_GYRO_PVAE) && pl.an_raw < kMinNorm) ? 0.0f : gan;
    return r;
}

// the general path out of line: it runs for a handful of pairs (if any), and inlining it next to the lean path in an
// 8-row unrolled loop would multiply the kernel's code size for nothing
#pragma nv_diag_suppress 177   // used by the SIMT kernels only; the tensor-core translation units include this header too
static __device__ __noinline__ float gyro_pair_fwd_general(float e, float q, float qa, float x2, float p2, float pa, float an_raw,
                                                    GyroParams P) {
    GyroPairCtx k;
    GyroDiff df;
    df.e = e; df.q = q; df.qa = qa;
    return gyro_pair_fwd(p2 - q, pa - qa, x2, p2, pa, an_raw, P, k, &df);
}

static __device__ __noinline__ GyroPairGrad gyro_pair_grad_general(float g, float e, float q, float qa, float x2, float p2, float pa,
                                                            float an_raw, GyroParams P, float bias) {
    GyroPairCtx k;
    GyroDiff df;
    df.e = e; df.q = q; df.qa = qa;
    const float px = p2 - q, xa = pa - qa;
    const float o = gyro_pair_fwd(px, xa, x2, p2, pa, an_raw, P, k, &df);
    if ((P.flags & HVAE_GYRO_RELU) && !(o + bias > 0.0f)) g = 0.0f;

    GyroPairGrad r = gyro_pair_bwd(g, px, xa, x2, p2, pa, an_raw, P, k);
    r.g_used = g;
    return r;
}

#pragma nv_diag_default 177

}  // namespace hvae
