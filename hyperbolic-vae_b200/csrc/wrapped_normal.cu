// K4 / K5 / fused latent head: WrappedNormal reparameterised sample, log-density, Monte-Carlo KL.
// reference: hyperbolic_vae/distributions/wrapped_normal.py:66-89, manifolds.py:25-35 (logdetexp),
//            models/vae_hyperbolic.py:126-127,191-216 (posterior built twice + prior per step).
//
// Each kernel is one pass: a row (length D) sits in the registers of G lanes, all norms / inner
// products are warp-shuffle reductions, nothing is staged in HBM between sample and KL.
// Algorithmic bytes per row (fp32): sample fwd 16D, bwd 24D; log_prob fwd 12D+4, bwd 24D+4;
// fused head fwd 16D+4 (eps injected), bwd 20D+4.
#include "hvae_common.cuh"

namespace hvae {

constexpr float kHalfLog2Pi = 0.91893853320467274178f;  // log(sqrt(2 pi))

// ---- reparameterised sample --------------------------------------------------------------------------
template <int G, int EPL>
struct SampleCtx {
    RowSlice<G, EPL> u, w, zpre;
    float mu2, m, lam, un_raw, un, th, t, sech2, pn;
    bool m_clamped, hit;
    MAddCtx ma;
};

// z = project(mu (+) tanh(sc lam ||u||/2)/sc * u/||u||),  u = (sigma*eps)/lam   (lam = lambda_mu; lambda_0 = 2 cancels)
// Per-element divisions of the reference are hoisted into per-row reciprocals (1-2 ulp different rounding).
template <int G, int EPL>
__device__ __forceinline__ void sample_row(const RowSlice<G, EPL>& mu, const RowSlice<G, EPL>& sig,
                                           const RowSlice<G, EPL>& eps, RowSlice<G, EPL>& z, SampleCtx<G, EPL>& k,
                                           const Ball& ball) {
    k.mu2 = sqnorm<G, EPL>(mu);
    const float m = 1.0f - ball.c * k.mu2;
    k.m_clamped = m < kMinNorm;
    k.m = fmaxf(m, kMinNorm);
    k.lam = 2.0f * rcpf(k.m);
    const float inv_lam = 0.5f * k.m;
#pragma unroll
    for (int i = 0; i < EPL; ++i) k.u.v[i] = (sig.v[i] * eps.v[i]) * inv_lam;
    k.un_raw = sqrt_fast(sqnorm<G, EPL>(k.u));
    k.un = fmaxf(k.un_raw, kMinNorm);
    k.th = ball.sc * ((k.lam * 0.5f) * k.un);
    tanh_sech2(k.th, k.t, k.sech2);
    const float q = ball.rsc * k.t * rcpf(k.un);
#pragma unroll
    for (int i = 0; i < EPL; ++i) k.w.v[i] = q * k.u.v[i];
    k.ma = mobius_add_raw<G, EPL>(mu, k.w, k.zpre, ball);
    z = k.zpre;
    k.hit = project_inplace<G, EPL>(z, ball, k.pn);
}

// g: dL/dz (post-projection). Adds into gmu, gsig.
template <int G, int EPL>
__device__ __forceinline__ void sample_row_bwd(const RowSlice<G, EPL>& mu, const RowSlice<G, EPL>& eps,
                                               const SampleCtx<G, EPL>& k, RowSlice<G, EPL> g,
                                               RowSlice<G, EPL>& gmu_acc, RowSlice<G, EPL>& gsig_acc, const Ball& ball) {
    project_bwd<G, EPL>(g, k.zpre, k.pn, k.hit, ball);
    RowSlice<G, EPL> gx, gw;
    mobius_add_raw_bwd<G, EPL>(mu, k.w, k.ma, g, gx, gw, ball);
    const float run = rcpf(k.un);
    const float q = ball.rsc * k.t * run;
    const float dq_dun = (k.sech2 * (k.lam * 0.5f) - q) * run;
    const float dq_dlam = k.sech2 * 0.5f;
    const float gwu = dot<G, EPL>(gw, k.u);
    const float coef = (k.un_raw >= kMinNorm) ? dq_dun * gwu * rcpf(k.un_raw) : 0.0f;
    RowSlice<G, EPL> gu;
#pragma unroll
    for (int i = 0; i < EPL; ++i) gu.v[i] = q * gw.v[i] + coef * k.u.v[i];
    // u = v/lam: gv = gu/lam ; glam += -(gu.u)/lam
    const float guu = dot<G, EPL>(gu, k.u);
    const float inv_lam = 0.5f * k.m;
    const float glam = dq_dlam * gwu - guu * inv_lam;
    const float rm = rcpf(k.m);
    const float gmu2 = k.m_clamped ? 0.0f : glam * (2.0f * rm * rm) * ball.c;
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
        gmu_acc.v[i] += gx.v[i] + 2.0f * gmu2 * mu.v[i];
        gsig_acc.v[i] += (gu.v[i] * inv_lam) * eps.v[i];
    }
}

// ---- log-density of WrappedNormal(mu, sigma) at z -----------------------------------------------------
template <int G, int EPL>
struct LogProbCtx {
    RowSlice<G, EPL> xn, s;  // xn = -mu, s = (-mu)(+)z
    float r, sn, phi, a, d;
    MAddCtx ma;
};

template <int G, int EPL, bool kScalarSigma>
__device__ __forceinline__ float logprob_row(const RowSlice<G, EPL>& mu, const RowSlice<G, EPL>& sig, float sigma0,
                                             const RowSlice<G, EPL>& z, LogProbCtx<G, EPL>& k, const Ball& ball, int D,
                                             int lg) {
    if (kScalarSigma) {
        // prior at the origin: (-0)(+)z = z exactly
        k.s = z;
        k.xn.zero();
        k.ma.A = 1.0f; k.ma.B = 1.0f; k.ma.den = 1.0f; k.ma.den_clamped = false; k.ma.x2 = 0.0f; k.ma.xy = 0.0f;
        k.ma.y2 = 0.0f;
    } else {
#pragma unroll
        for (int i = 0; i < EPL; ++i) k.xn.v[i] = -mu.v[i];
        k.ma = mobius_add_raw<G, EPL>(k.xn, z, k.s, ball);
    }
    k.r = sqrt_fast(sqnorm<G, EPL>(k.s));
    k.sn = fmaxf(k.r, kMinNorm);
    k.a = ball.sc * k.sn;
    const float at = artanh_c(k.a);
    k.phi = 2.0f * at * rcpf(k.a);  // u = phi * s  (lambda_mu cancels: logmap /lam, transp *lam/2, *lambda_0)
    float acc = 0.0f;
    if (kScalarSigma) {
        // sum_i -(phi s_i)^2/(2 s0^2) - D log s0 - D log sqrt(2 pi)
        const float ss = k.r * k.r;
        acc = -(k.phi * k.phi * ss) * (0.5f * rcpf(sigma0 * sigma0)) - (float)D * (__logf(sigma0) + kHalfLog2Pi);
    } else {
#pragma unroll
        for (int i = 0; i < EPL; ++i) {
            const int idx = RowSlice<G, EPL>::index(lg, i, D);
            if (idx < D) {
                const float sg = sig.v[i];
                const float ui = k.phi * k.s.v[i] * rcpf(sg);
                acc += -0.5f * (ui * ui) - __logf(sg) - kHalfLog2Pi;
            }
        }
        acc = group_sum<G>(acc);
    }
    // d = 2 artanh(clamp(sc r))/sc (no clamp_min on r in the reference's dist): identical to 2 at/sc unless r < 1e-15
    k.d = (k.r >= kMinNorm) ? 2.0f * ball.rsc * at : 2.0f * ball.rsc * artanh_c(ball.sc * k.r);
    // logdetexp = (D-1) (log sinh(sc d) - log sc - log d) = (D-1) log(sinh(sc d)/(sc d))
    return acc - (float)(D - 1) * log_sinhc(ball.sc * k.d);
}

// g: upstream scalar. Accumulates gmu, gsig, gz.
template <int G, int EPL, bool kScalarSigma>
__device__ __forceinline__ void logprob_row_bwd(const RowSlice<G, EPL>& sig, float sigma0, const RowSlice<G, EPL>& z,
                                                const LogProbCtx<G, EPL>& k, float g, RowSlice<G, EPL>& gmu_acc,
                                                RowSlice<G, EPL>& gsig_acc, RowSlice<G, EPL>& gz_acc, const Ball& ball,
                                                int D, int lg) {
    RowSlice<G, EPL> gu;
    const float rs0 = kScalarSigma ? rcpf(sigma0 * sigma0) : 0.0f;
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
        const int idx = RowSlice<G, EPL>::index(lg, i, D);
        const bool on = idx < D;
        const float ui = k.phi * k.s.v[i];
        if (kScalarSigma) {
            gu.v[i] = on ? -g * ui * rs0 : 0.0f;
        } else {
            const float rsg = on ? rcpf(sig.v[i]) : 0.0f;
            const float un = ui * rsg;  // u_i / sigma_i
            gu.v[i] = -g * un * rsg;
            if (on) gsig_acc.v[i] += g * (un * un - 1.0f) * rsg;
        }
    }
    const float rsn = rcpf(k.sn);
    const float dphi = (2.0f * artanh_grad(k.a) - k.phi) * rsn;
    const float gus = dot<G, EPL>(gu, k.s);
    const float rr = (k.r > 0.0f) ? rcpf(k.r) : 0.0f;
    float coef = (k.r >= kMinNorm) ? dphi * gus * rr : 0.0f;
    // -(D-1) d/dr log_sinhc(sc d(r)),  dd/dr = 2 artanh'(sc r)
    {
        const float dL = (float)(D - 1) * dlog_sinhc(ball.sc * k.d) * ball.sc * (2.0f * artanh_grad(ball.sc * k.r));
        coef -= g * dL * rr;
    }
    RowSlice<G, EPL> gs;
#pragma unroll
    for (int i = 0; i < EPL; ++i) gs.v[i] = k.phi * gu.v[i] + coef * k.s.v[i];
    if (kScalarSigma) {
#pragma unroll
        for (int i = 0; i < EPL; ++i) gz_acc.v[i] += gs.v[i];
    } else {
        RowSlice<G, EPL> gx, gy;
        mobius_add_raw_bwd<G, EPL>(k.xn, z, k.ma, gs, gx, gy, ball);
#pragma unroll
        for (int i = 0; i < EPL; ++i) {
            gmu_acc.v[i] -= gx.v[i];
            gz_acc.v[i] += gy.v[i];
        }
    }
}

// ---- posterior log-density AT ITS OWN SAMPLE, closed form -------------------------------------------------
// For z = rsample(mu, sigma; eps) with no clamp / projection binding, left-cancellation gives (-mu)(+)z = w, the
// tangent vector recovered by log_prob is exactly v = sigma*eps and dist(mu, z) = |v|, so
//   log q(z) = sum_i [-eps_i^2/2 - log sigma_i - log sqrt(2 pi)] - (D-1) log(sinh(sc |v|)/(sc |v|)),
// independent of mu.  The reference evaluates the same quantity the long way round (logmap, transport, Normal
// log-pdf); the two agree to rounding, this one is the better conditioned.  Rows where a clamp or the projection
// binds take the general path (logprob_row).
template <int G, int EPL>
__device__ __forceinline__ bool sample_is_regular(const SampleCtx<G, EPL>& k) {
    return !k.hit && !k.m_clamped && !k.ma.den_clamped && k.un_raw >= kMinNorm && k.th <= 8.0f;
}

template <int G, int EPL>
__device__ __forceinline__ float logq_at_sample(const RowSlice<G, EPL>& sig, const RowSlice<G, EPL>& eps,
                                                const SampleCtx<G, EPL>& k, const Ball& ball, int D, int lg) {
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
        const int idx = RowSlice<G, EPL>::index(lg, i, D);
        if (idx < D) acc += -0.5f * eps.v[i] * eps.v[i] - __logf(sig.v[i]) - kHalfLog2Pi;
    }
    acc = group_sum<G>(acc);
    const float nv = k.lam * k.un;  // |sigma*eps|
    return acc - (float)(D - 1) * log_sinhc(ball.sc * nv);
}

// d log q / d sigma_i on the regular path (d/d mu = 0)
template <int G, int EPL>
__device__ __forceinline__ void logq_at_sample_bwd(const RowSlice<G, EPL>& sig, const RowSlice<G, EPL>& eps,
                                                   const SampleCtx<G, EPL>& k, float g, RowSlice<G, EPL>& gsig_acc,
                                                   const Ball& ball, int D, int lg) {
    const float nv = k.lam * k.un;
    const float cL = (float)(D - 1) * dlog_sinhc(ball.sc * nv) * ball.sc * rcpf(nv);
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
        const int idx = RowSlice<G, EPL>::index(lg, i, D);
        if (idx < D) gsig_acc.v[i] += g * (-rcpf(sig.v[i]) - cL * sig.v[i] * eps.v[i] * eps.v[i]);
    }
}

// =================================================================================================
// kernels
// =================================================================================================
template <int G, int EPL>
__global__ void __launch_bounds__(kRowThreads) k_wrapped_sample_fwd(const float* __restrict__ mu,
                                                                     const float* __restrict__ sigma,
                                                                     const float* __restrict__ eps, float* __restrict__ z,
                                                                     int64_t S, int64_t B, int D, Ball ball) {
    HVAE_ROW_PROLOGUE(G)
    const int64_t rows = S * B;
    constexpr int U = HVAE_ROW_UNROLL(EPL);
    for (int64_t r0 = warp_global * (RPW * U); r0 < rows; r0 += warps_total * (RPW * U)) {
        RowSlice<G, EPL> m[U], sg[U], e[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const int64_t row = r0 + j * RPW + sub;
            const bool valid = row < rows;
            const int64_t b = valid ? row % B : 0;
            m[j].load(mu, b, D, lg, valid);
            sg[j].load(sigma, b, D, lg, valid);
            e[j].load(eps, row, D, lg, valid);
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const int64_t row = r0 + j * RPW + sub;
            RowSlice<G, EPL> zr;
            SampleCtx<G, EPL> k;
            sample_row<G, EPL>(m[j], sg[j], e[j], zr, k, ball);
            zr.store(z, row, D, lg, row < rows);
        }
    }
}

// one lane group per b; loops over the sample dim so that gmu/gsigma are complete sums
template <int G, int EPL>
__global__ void __launch_bounds__(kRowThreads) k_wrapped_sample_bwd(const float* __restrict__ mu,
                                                                     const float* __restrict__ sigma,
                                                                     const float* __restrict__ eps,
                                                                     const float* __restrict__ gz, float* __restrict__ gmu,
                                                                     float* __restrict__ gsigma, int64_t S, int64_t B, int D,
                                                                     Ball ball) {
    HVAE_ROW_PROLOGUE(G)
    for (int64_t r0 = warp_global * RPW; r0 < B; r0 += warps_total * RPW) {
        const int64_t b = r0 + sub;
        const bool valid = b < B;
        RowSlice<G, EPL> m, sg, gm, gs;
        m.load(mu, b, D, lg, valid);
        sg.load(sigma, b, D, lg, valid);
        gm.zero();
        gs.zero();
        for (int64_t s = 0; s < S; ++s) {
            RowSlice<G, EPL> e, g, zr;
            e.load(eps, s * B + b, D, lg, valid);
            g.load(gz, s * B + b, D, lg, valid);
            SampleCtx<G, EPL> k;
            sample_row<G, EPL>(m, sg, e, zr, k, ball);
            sample_row_bwd<G, EPL>(m, e, k, g, gm, gs, ball);
        }
        gm.store(gmu, b, D, lg, valid);
        gs.store(gsigma, b, D, lg, valid);
    }
}

template <int G, int EPL, bool kPrior>
__global__ void __launch_bounds__(kRowThreads) k_wrapped_logprob_fwd(const float* __restrict__ mu,
                                                                      const float* __restrict__ sigma, float sigma0,
                                                                      const float* __restrict__ z, float* __restrict__ logp,
                                                                      int64_t S, int64_t B, int D, Ball ball) {
    HVAE_ROW_PROLOGUE(G)
    const int64_t rows = S * B;
    for (int64_t r0 = warp_global * RPW; r0 < rows; r0 += warps_total * RPW) {
        const int64_t row = r0 + sub;
        const bool valid = row < rows;
        const int64_t b = valid ? row % B : 0;
        RowSlice<G, EPL> m, sg, zr;
        if (kPrior) {
            m.zero();
            sg.zero();
        } else {
            m.load(mu, b, D, lg, valid);
            sg.load(sigma, b, D, lg, valid);
            if (!valid) {
#pragma unroll
                for (int i = 0; i < EPL; ++i) sg.v[i] = 1.0f;
            }
        }
        zr.load(z, row, D, lg, valid);
        LogProbCtx<G, EPL> k;
        const float lp = logprob_row<G, EPL, kPrior>(m, sg, sigma0, zr, k, ball, D, lg);
        if (valid && lg == 0) logp[row] = lp;
    }
}

template <int G, int EPL, bool kPrior>
__global__ void __launch_bounds__(kRowThreads) k_wrapped_logprob_bwd(const float* __restrict__ mu,
                                                                      const float* __restrict__ sigma, float sigma0,
                                                                      const float* __restrict__ z,
                                                                      const float* __restrict__ glogp, float* __restrict__ gmu,
                                                                      float* __restrict__ gsigma, float* __restrict__ gz,
                                                                      int64_t S, int64_t B, int D, Ball ball) {
    HVAE_ROW_PROLOGUE(G)
    for (int64_t r0 = warp_global * RPW; r0 < B; r0 += warps_total * RPW) {
        const int64_t b = r0 + sub;
        const bool valid = b < B;
        RowSlice<G, EPL> m, sg, gm, gs;
        if (kPrior) {
            m.zero();
            sg.zero();
        } else {
            m.load(mu, b, D, lg, valid);
            sg.load(sigma, b, D, lg, valid);
            if (!valid) {
#pragma unroll
                for (int i = 0; i < EPL; ++i) sg.v[i] = 1.0f;
            }
        }
        gm.zero();
        gs.zero();
        for (int64_t s = 0; s < S; ++s) {
            const int64_t row = s * B + b;
            RowSlice<G, EPL> zr, gzr;
            zr.load(z, row, D, lg, valid);
            gzr.zero();
            const float g = valid ? __ldg(glogp + row) : 0.0f;
            LogProbCtx<G, EPL> k;
            logprob_row<G, EPL, kPrior>(m, sg, sigma0, zr, k, ball, D, lg);
            logprob_row_bwd<G, EPL, kPrior>(sg, sigma0, zr, k, g, gm, gs, gzr, ball, D, lg);
            if (gz) gzr.store(gz, row, D, lg, valid);
        }
        if (!kPrior) {
            if (gmu) gm.store(gmu, b, D, lg, valid);
            if (gsigma) gs.store(gsigma, b, D, lg, valid);
        }
    }
}

// ---- fused latent head: z = rsample, kl = log q(z) - log p(z) ------------------------------------------
// Lean evaluation: with v = sigma*eps, every norm / inner product the sample needs follows from FOUR row
// reductions (|mu|^2, |v|^2, <mu,v>, sum_i[-eps_i^2/2 - log sigma_i]):  w = qv*v, |w|^2 = qv^2|v|^2,
// <mu,w> = qv<mu,v>, |z_pre|^2 = (A^2|mu|^2 + 2AB<mu,w> + B^2|w|^2)/den^2, |z| = min(|z_pre|, maxnorm) — so the
// projection test and the whole prior density (a function of |z| only) cost no further reductions.
template <int G, int EPL>
struct HeadScalars {
    float mu2, vv, muv, lqa;
    float vn, qv, inv_lam, rz, rho, at_p;
    bool regular;
};

template <int G, int EPL>
__device__ __forceinline__ void head_forward(const RowSlice<G, EPL>& mu, const RowSlice<G, EPL>& sig,
                                             const RowSlice<G, EPL>& eps, RowSlice<G, EPL>& z, SampleCtx<G, EPL>& k,
                                             HeadScalars<G, EPL>& h, const Ball& ball, int D, int lg) {
    float mu2 = 0.0f, vv = 0.0f, muv = 0.0f, lqa = 0.0f;
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
        const int idx = RowSlice<G, EPL>::index(lg, i, D);
        const float v = sig.v[i] * eps.v[i];
        k.u.v[i] = v;  // holds v for now
        mu2 = fmaf(mu.v[i], mu.v[i], mu2);
        vv = fmaf(v, v, vv);
        muv = fmaf(mu.v[i], v, muv);
        if (idx < D) lqa += -0.5f * eps.v[i] * eps.v[i] - __logf(sig.v[i]) - kHalfLog2Pi;
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
        mu2 += __shfl_xor_sync(0xffffffffu, mu2, o);
        vv += __shfl_xor_sync(0xffffffffu, vv, o);
        muv += __shfl_xor_sync(0xffffffffu, muv, o);
        lqa += __shfl_xor_sync(0xffffffffu, lqa, o);
    }
    h.mu2 = mu2; h.vv = vv; h.muv = muv; h.lqa = lqa;
    const float c = ball.c;
    k.mu2 = mu2;
    const float m0 = 1.0f - c * mu2;
    k.m_clamped = m0 < kMinNorm;
    k.m = fmaxf(m0, kMinNorm);
    const float rm = rcpf(k.m);
    k.lam = 2.0f * rm;
    h.inv_lam = 0.5f * k.m;
    h.vn = sqrt_fast(vv);
    k.un_raw = h.inv_lam * h.vn;
    k.un = fmaxf(k.un_raw, kMinNorm);
    k.th = ball.sc * (k.un * rm);
    tanh_sech2(k.th, k.t, k.sech2);
    const float q = ball.rsc * k.t * rcpf(k.un);
    h.qv = q * h.inv_lam;
    MAddCtx& ma = k.ma;
    ma.x2 = mu2;
    ma.y2 = h.qv * h.qv * vv;
    ma.xy = h.qv * muv;
    ma.A = 1.0f + 2.0f * c * ma.xy + c * ma.y2;
    ma.B = 1.0f - c * mu2;
    const float den0 = 1.0f + 2.0f * c * ma.xy + c * c * mu2 * ma.y2;
    ma.den_clamped = den0 < kMinNorm;
    ma.den = fmaxf(den0, kMinNorm);
    const float rden = rcpf(ma.den);
    const float ar = ma.A * rden, br = ma.B * rden * h.qv;
    const float nz2 = fmaxf(ma.A * ma.A * mu2 + 2.0f * ma.A * ma.B * ma.xy + ma.B * ma.B * ma.y2, 0.0f) * rden * rden;
    const float nz = sqrt_fast(nz2);
    k.pn = fmaxf(nz, kMinNorm);
    k.hit = k.pn > ball.maxnorm;
    const float scale = k.hit ? ball.maxnorm * rcpf(k.pn) : 1.0f;
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
        const float v = k.u.v[i];
        k.u.v[i] = v * h.inv_lam;
        k.w.v[i] = v * h.qv;
        k.zpre.v[i] = fmaf(ar, mu.v[i], br * v);
        z.v[i] = k.zpre.v[i] * scale;
    }
    h.rz = k.hit ? ball.maxnorm * (nz * rcpf(k.pn)) : nz;
    // prior WrappedNormal(0, sigma0): u_p = rho * z/|z|, rho = 2 artanh(sc |z|)/sc = dist(0, z)
    const float rzc = fmaxf(h.rz, kMinNorm);
    h.at_p = artanh_c(ball.sc * rzc);
    h.rho = (h.rz >= kMinNorm) ? 2.0f * ball.rsc * h.at_p : 2.0f * h.at_p * rcpf(ball.sc * rzc) * h.rz;
    h.regular = !k.hit && !k.m_clamped && !ma.den_clamped && k.un_raw >= kMinNorm && k.th <= 8.0f;
}

template <int G, int EPL>
__device__ __forceinline__ float head_logp(const HeadScalars<G, EPL>& h, float sigma0, const Ball& ball, int D) {
    return -(h.rho * h.rho) * (0.5f * rcpf(sigma0 * sigma0)) - (float)D * (__logf(sigma0) + kHalfLog2Pi) -
           (float)(D - 1) * log_sinhc(ball.sc * h.rho);
}

template <int G, int EPL>
__global__ void __launch_bounds__(kRowThreads, (EPL >= 8) ? 3 : 1) k_latent_head_fwd(const float* __restrict__ mu, const float* __restrict__ sigma,
                                                                  const float* __restrict__ eps, float prior_scale,
                                                                  float* __restrict__ z, float* __restrict__ kl, int64_t B,
                                                                  int D, Ball ball) {
    HVAE_ROW_PROLOGUE(G)
    constexpr int U = 1;
    for (int64_t r0 = warp_global * (RPW * U); r0 < B; r0 += warps_total * (RPW * U)) {
        RowSlice<G, EPL> mm[U], ss[U], ee[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const int64_t b = r0 + j * RPW + sub;
            const bool valid = b < B;
            mm[j].load(mu, b, D, lg, valid);
            ss[j].load(sigma, b, D, lg, valid);
            if (!valid) {
#pragma unroll
                for (int i = 0; i < EPL; ++i) ss[j].v[i] = 1.0f;
            }
            ee[j].load(eps, b, D, lg, valid);
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const int64_t b = r0 + j * RPW + sub;
            const bool valid = b < B;
            const RowSlice<G, EPL>& m = mm[j];
            const RowSlice<G, EPL>& sg = ss[j];
            const RowSlice<G, EPL>& e = ee[j];
            RowSlice<G, EPL> zr;
            SampleCtx<G, EPL> k;
            HeadScalars<G, EPL> h;
            head_forward<G, EPL>(m, sg, e, zr, k, h, ball, D, lg);
            zr.store(z, b, D, lg, valid);
            float lq = h.lqa - (float)(D - 1) * log_sinhc(ball.sc * h.vn);
            if (__any_sync(0xffffffffu, valid && !h.regular)) {  // warp-uniform: the general path shuffles
                LogProbCtx<G, EPL> kq;
                const float lq_gen = logprob_row<G, EPL, false>(m, sg, 0.0f, zr, kq, ball, D, lg);
                if (!h.regular) lq = lq_gen;
            }
            if (valid && lg == 0) kl[b] = lq - head_logp<G, EPL>(h, prior_scale, ball, D);
        }
    }
}

// Backward of the head in closed form.  z_pre = a mu + b v with a = A/den, b = B qv/den and every scalar a
// function of (|mu|^2, |v|^2, <mu,v>), so with H = dL/dz the pull-back needs only two more reductions (<H,mu>, <H,v>):
//   h' = project^T H;  g_a = <h',mu>, g_b = <h',v>;  scalar chain -> g_|mu|^2, g_|v|^2, g_<mu,v>;
//   g_mu = a h' + 2 g_mu2 mu + g_muv v ;  g_v = b h' + 2 g_vv v + g_muv mu ;  g_sigma = g_v * eps.
// qv = tanh(sc|v|/2)/(sc|v|) depends on |v| only (lambda_mu cancels).  Rows where a clamp binds (m, |u|, den)
// take the general path.
template <int G, int EPL>
__global__ void __launch_bounds__(kRowThreads, (EPL >= 8) ? 2 : 1) k_latent_head_bwd(const float* __restrict__ mu, const float* __restrict__ sigma,
                                                                  const float* __restrict__ eps, float prior_scale,
                                                                  const float* __restrict__ gz, const float* __restrict__ gkl,
                                                                  float* __restrict__ gmu, float* __restrict__ gsigma,
                                                                  int64_t B, int D, Ball ball) {
    HVAE_ROW_PROLOGUE(G)
    constexpr int U = 1;
    for (int64_t r0 = warp_global * (RPW * U); r0 < B; r0 += warps_total * (RPW * U)) {
      RowSlice<G, EPL> mm[U], ss[U], ee[U], gg[U];
#pragma unroll
      for (int j = 0; j < U; ++j) {
          const int64_t bj = r0 + j * RPW + sub;
          const bool vj = bj < B;
          mm[j].load(mu, bj, D, lg, vj);
          ss[j].load(sigma, bj, D, lg, vj);
          if (!vj) {
#pragma unroll
              for (int i = 0; i < EPL; ++i) ss[j].v[i] = 1.0f;
          }
          ee[j].load(eps, bj, D, lg, vj);
          if (gz) gg[j].load(gz, bj, D, lg, vj); else gg[j].zero();
      }
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const int64_t b = r0 + j * RPW + sub;
        const bool valid = b < B;
        const RowSlice<G, EPL>& m = mm[j];
        const RowSlice<G, EPL>& sg = ss[j];
        const RowSlice<G, EPL>& e = ee[j];
        RowSlice<G, EPL> zr, gm, gs;
        RowSlice<G, EPL>& gzt = gg[j];
        gm.zero();
        gs.zero();
        const float g = (gkl && valid) ? __ldg(gkl + b) : 0.0f;
        SampleCtx<G, EPL> k;
        HeadScalars<G, EPL> h;
        head_forward<G, EPL>(m, sg, e, zr, k, h, ball, D, lg);
        const float c = ball.c;
        // H = gz - g * d log p / dz :  log p = f(|z|),  f' = [-rho/s0^2 - (D-1) sc L'(sc rho)] * 2 artanh'(sc |z|)
        {
            const float fp = (-h.rho * rcpf(prior_scale * prior_scale) - (float)(D - 1) * ball.sc * dlog_sinhc(ball.sc * h.rho)) *
                             (2.0f * artanh_grad(ball.sc * h.rz));
            const float coef = (h.rz >= kMinNorm) ? -g * fp * rcpf(h.rz) : 0.0f;
#pragma unroll
            for (int i = 0; i < EPL; ++i) gzt.v[i] = fmaf(coef, zr.v[i], gzt.v[i]);
        }
        if (__any_sync(0xffffffffu, valid && !h.regular)) {  // warp-uniform: the general path shuffles
            LogProbCtx<G, EPL> kq;
            logprob_row<G, EPL, false>(m, sg, 0.0f, zr, kq, ball, D, lg);
            logprob_row_bwd<G, EPL, false>(sg, 0.0f, zr, kq, h.regular ? 0.0f : g, gm, gs, gzt, ball, D, lg);
        }
        if (h.regular) {
            // d log q / d sigma_i = -1/sigma_i - (D-1) sc L'(sc |v|) sigma_i eps_i^2 / |v|   (d/d mu = 0)
            const float cL = (float)(D - 1) * dlog_sinhc(ball.sc * h.vn) * ball.sc * rcpf(h.vn);
#pragma unroll
            for (int i = 0; i < EPL; ++i) {
                const int idx = RowSlice<G, EPL>::index(lg, i, D);
                if (idx < D) gs.v[i] += g * (-rcpf(sg.v[i]) - cL * sg.v[i] * e.v[i] * e.v[i]);
            }
        }
        const bool clamp_free = !k.m_clamped && !k.ma.den_clamped && k.un_raw >= kMinNorm;
        if (__any_sync(0xffffffffu, valid && !clamp_free)) {
            RowSlice<G, EPL> gm2, gs2;
            gm2.zero();
            gs2.zero();
            sample_row_bwd<G, EPL>(m, e, k, gzt, gm2, gs2, ball);
            if (!clamp_free) {
#pragma unroll
                for (int i = 0; i < EPL; ++i) { gm.v[i] += gm2.v[i]; gs.v[i] += gs2.v[i]; }
            }
        }
        {
            // two reductions: <H, mu>, <H, v>   (v = sigma*eps = u / inv_lam; use w = qv v to avoid recomputing)
            float hm = 0.0f, hv = 0.0f;
#pragma unroll
            for (int i = 0; i < EPL; ++i) {
                hm = fmaf(gzt.v[i], m.v[i], hm);
                hv = fmaf(gzt.v[i], sg.v[i] * e.v[i], hv);
            }
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) {
                hm += __shfl_xor_sync(0xffffffffu, hm, o);
                hv += __shfl_xor_sync(0xffffffffu, hv, o);
            }
            if (clamp_free) {
                const MAddCtx& ma = k.ma;
                const float rden = rcpf(ma.den);
                const float a = ma.A * rden, bq = ma.B * h.qv * rden;
                // projection: h' = s (H - (<H,zpre>/n^2) zpre)
                float s = 1.0f, pr = 0.0f;
                if (k.hit) {
                    const float rn = rcpf(k.pn);
                    s = ball.maxnorm * rn;
                    pr = (a * hm + bq * hv) * rn * rn;
                }
                const float zm = a * h.mu2 + bq * h.muv;   // <zpre, mu>
                const float zv = a * h.muv + bq * h.vv;    // <zpre, v>
                const float ga = s * (hm - pr * zm);       // <h', mu>
                const float gb = s * (hv - pr * zv);       // <h', v>
                const float gA = ga * rden;
                const float gB = gb * h.qv * rden;
                float gqv = gb * ma.B * rden;
                const float gden = -(ga * a + gb * bq) * rden;
                const float gxy = 2.0f * c * (gA + gden);
                const float gy2 = c * gA + c * c * h.mu2 * gden;
                const float gmu2 = -c * gB + c * c * ma.y2 * gden;
                gqv += gxy * h.muv + gy2 * 2.0f * h.qv * h.vv;
                const float gmuv = gxy * h.qv;
                // qv = tanh(sc|v|/2)/(sc|v|):  d qv / d vv = (sech^2/2 - qv) / (2 vv)
                const float gvv = gy2 * h.qv * h.qv + gqv * (0.5f * k.sech2 - h.qv) * (0.5f * rcpf(h.vv));
#pragma unroll
                for (int i = 0; i < EPL; ++i) {
                    const float v = sg.v[i] * e.v[i];
                    const float hp = s * (gzt.v[i] - pr * k.zpre.v[i]);
                    gm.v[i] += a * hp + 2.0f * gmu2 * m.v[i] + gmuv * v;
                    gs.v[i] += (bq * hp + 2.0f * gvv * v + gmuv * m.v[i]) * e.v[i];
                }
            }
        }
        gm.store(gmu, b, D, lg, valid);
        gs.store(gsigma, b, D, lg, valid);
      }
    }
}

}  // namespace hvae

using namespace hvae;

#define HVAE_CHECK_SBD(S, B, D)                                                   \
    if ((S) < 0 || (B) < 0 || (D) <= 0 || (D) > kMaxRowDim) return HVAE_ESHAPE;    \
    if ((S) == 0 || (B) == 0) return HVAE_OK;

#define HVAE_DISPATCH_T(KERN, TARG, D, rows, s, ...)                                                                 \
    do {                                                                                                             \
        if ((D) <= 2)        KERN<1, 2, TARG><<<row_grid((rows), 1), kRowThreads, 0, (s)>>>(__VA_ARGS__);             \
        else if ((D) <= 4)   KERN<1, 4, TARG><<<row_grid((rows), 1), kRowThreads, 0, (s)>>>(__VA_ARGS__);             \
        else if ((D) <= 8)   KERN<1, 8, TARG><<<row_grid((rows), 1), kRowThreads, 0, (s)>>>(__VA_ARGS__);             \
        else if ((D) <= 16)  KERN<2, 8, TARG><<<row_grid((rows), 2), kRowThreads, 0, (s)>>>(__VA_ARGS__);             \
        else if ((D) <= 32)  KERN<4, 8, TARG><<<row_grid((rows), 4), kRowThreads, 0, (s)>>>(__VA_ARGS__);             \
        else if ((D) <= 64)  KERN<8, 8, TARG><<<row_grid((rows), 8), kRowThreads, 0, (s)>>>(__VA_ARGS__);           \
        else if ((D) <= 128) KERN<16, 8, TARG><<<row_grid((rows), 16), kRowThreads, 0, (s)>>>(__VA_ARGS__);           \
        else if ((D) <= 256) KERN<32, 8, TARG><<<row_grid((rows), 32), kRowThreads, 0, (s)>>>(__VA_ARGS__);           \
        else if ((D) <= 512) KERN<32, 16, TARG><<<row_grid((rows), 32), kRowThreads, 0, (s)>>>(__VA_ARGS__);          \
        else                 KERN<32, 32, TARG><<<row_grid((rows), 32), kRowThreads, 0, (s)>>>(__VA_ARGS__);          \
    } while (0)

extern "C" int hvae_wrapped_sample_fwd_f32(const float* mu, const float* sigma, const float* eps, float* z, int64_t S,
                                           int64_t B, int64_t D, float c, void* stream) {
    HVAE_CHECK_SBD(S, B, D)
    if (!mu || !sigma || !eps || !z) return HVAE_EARG;
    HVAE_ROW_DISPATCH(k_wrapped_sample_fwd, D, S * B, (cudaStream_t)stream, mu, sigma, eps, z, S, B, (int)D, make_ball(c));
    return check_launch();
}

extern "C" int hvae_wrapped_sample_bwd_f32(const float* mu, const float* sigma, const float* eps, const float* gz,
                                           float* gmu, float* gsigma, int64_t S, int64_t B, int64_t D, float c,
                                           void* stream) {
    HVAE_CHECK_SBD(S, B, D)
    if (!mu || !sigma || !eps || !gz || !gmu || !gsigma) return HVAE_EARG;
    HVAE_ROW_DISPATCH(k_wrapped_sample_bwd, D, B, (cudaStream_t)stream, mu, sigma, eps, gz, gmu, gsigma, S, B, (int)D,
                      make_ball(c));
    return check_launch();
}

extern "C" int hvae_wrapped_logprob_fwd_f32(const float* mu, const float* sigma, float sigma0, const float* z, float* logp,
                                            int64_t S, int64_t B, int64_t D, float c, void* stream) {
    HVAE_CHECK_SBD(S, B, D)
    if (!z || !logp || ((mu == nullptr) != (sigma == nullptr))) return HVAE_EARG;
    cudaStream_t s = (cudaStream_t)stream;
    const Ball ball = make_ball(c);
    if (mu) HVAE_DISPATCH_T(k_wrapped_logprob_fwd, false, D, S * B, s, mu, sigma, sigma0, z, logp, S, B, (int)D, ball);
    else    HVAE_DISPATCH_T(k_wrapped_logprob_fwd, true, D, S * B, s, mu, sigma, sigma0, z, logp, S, B, (int)D, ball);
    return check_launch();
}

extern "C" int hvae_wrapped_logprob_bwd_f32(const float* mu, const float* sigma, float sigma0, const float* z,
                                            const float* glogp, float* gmu, float* gsigma, float* gz, int64_t S, int64_t B,
                                            int64_t D, float c, void* stream) {
    HVAE_CHECK_SBD(S, B, D)
    if (!z || !glogp || ((mu == nullptr) != (sigma == nullptr))) return HVAE_EARG;
    cudaStream_t s = (cudaStream_t)stream;
    const Ball ball = make_ball(c);
    if (mu) HVAE_DISPATCH_T(k_wrapped_logprob_bwd, false, D, B, s, mu, sigma, sigma0, z, glogp, gmu, gsigma, gz, S, B, (int)D, ball);
    else    HVAE_DISPATCH_T(k_wrapped_logprob_bwd, true, D, B, s, mu, sigma, sigma0, z, glogp, gmu, gsigma, gz, S, B, (int)D, ball);
    return check_launch();
}

extern "C" int hvae_latent_head_fwd_f32(const float* mu, const float* sigma, const float* eps, float prior_scale, float* z,
                                        float* kl, int64_t B, int64_t D, float c, void* stream) {
    HVAE_CHECK_SBD(1, B, D)
    if (!mu || !sigma || !eps || !z || !kl) return HVAE_EARG;
    HVAE_ROW_DISPATCH(k_latent_head_fwd, D, B, (cudaStream_t)stream, mu, sigma, eps, prior_scale, z, kl, B, (int)D,
                      make_ball(c));
    return check_launch();
}

extern "C" int hvae_latent_head_bwd_f32(const float* mu, const float* sigma, const float* eps, float prior_scale,
                                        const float* gz, const float* gkl, float* gmu, float* gsigma, int64_t B, int64_t D,
                                        float c, void* stream) {
    HVAE_CHECK_SBD(1, B, D)
    if (!mu || !sigma || !eps || !gmu || !gsigma) return HVAE_EARG;
    HVAE_ROW_DISPATCH(k_latent_head_bwd, D, B, (cudaStream_t)stream, mu, sigma, eps, prior_scale, gz, gkl, gmu, gsigma, B,
                      (int)D, make_ball(c));
    return check_launch();
}
