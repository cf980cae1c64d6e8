// Fused Monte-Carlo KL of two Riemannian normals on the ball, posterior N(mu_b, sigma_b) against an origin-centred
// prior N(0, sigma_p):   kl[s,b] = log q(z) - log p(z)
//   = [-d(mu_b, z)^2/(2 sigma_b^2) - logZ(sigma_b)] - [-d(0, z)^2/(2 sigma_p^2) - logZ(sigma_p)]   (log|S^{D-1}| cancels)
// reference: pvae RiemannianNormal.log_prob (App. A.2) as used by the objective of
// hyperbolic_vae/training/old_pvae_train.py:53-58 — there two log_prob graphs of ~15 eager ops each.
// One row kernel forward, one backward (grads to mu, sigma_q, logZ_q, z); HBM-bound, 8D+16 bytes per row forward.
#include "hvae_common.cuh"

namespace hvae {

template <int G, int EPL>
__global__ void __launch_bounds__(kRowThreads)
k_rn_kl_fwd(const float* __restrict__ mu, const float* __restrict__ sigq, const float* __restrict__ logzq,
            const float* __restrict__ z, const float* __restrict__ sigp, const float* __restrict__ logzp,
            float* __restrict__ kl, int64_t S, int64_t B, int D, Ball ball) {
    HVAE_ROW_PROLOGUE(G)
    const int64_t rows = S * B;
    const float sp = __ldg(sigp), lzp = __ldg(logzp);
    for (int64_t r0 = warp_global * RPW; r0 < rows; r0 += warps_total * RPW) {
        const int64_t row = r0 + sub;
        const bool valid = row < rows;
        const int64_t b = valid ? row % B : 0;
        RowSlice<G, EPL> m, zr, s;
        m.load(mu, b, D, lg, valid);
        zr.load(z, row, D, lg, valid);
#pragma unroll
        for (int i = 0; i < EPL; ++i) m.v[i] = -m.v[i];
        mobius_add_raw<G, EPL>(m, zr, s, ball);
        const float r = sqrt_fast(sqnorm<G, EPL>(s));
        const float rz = sqrt_fast(sqnorm<G, EPL>(zr));
        const float dq = 2.0f * ball.rsc * artanh_c(ball.sc * r);
        const float dp = 2.0f * ball.rsc * artanh_c(ball.sc * rz);
        if (valid && lg == 0) {
            const float sq = __ldg(sigq + b);
            kl[row] = -dq * dq * (0.5f * rcpf(sq * sq)) - __ldg(logzq + b) + dp * dp * (0.5f * rcpf(sp * sp)) + lzp;
        }
    }
}

template <int G, int EPL>
__global__ void __launch_bounds__(kRowThreads)
k_rn_kl_bwd(const float* __restrict__ mu, const float* __restrict__ sigq, const float* __restrict__ z,
            const float* __restrict__ sigp, const float* __restrict__ gkl, float* __restrict__ gmu,
            float* __restrict__ gsigq, float* __restrict__ glogzq, float* __restrict__ gz, int64_t S, int64_t B, int D,
            Ball ball) {
    HVAE_ROW_PROLOGUE(G)
    const float sp = __ldg(sigp);
    for (int64_t r0 = warp_global * RPW; r0 < B; r0 += warps_total * RPW) {
        const int64_t b = r0 + sub;
        const bool valid = b < B;
        RowSlice<G, EPL> xn, gm;
        xn.load(mu, b, D, lg, valid);
#pragma unroll
        for (int i = 0; i < EPL; ++i) xn.v[i] = -xn.v[i];
        gm.zero();
        const float sq = valid ? __ldg(sigq + b) : 1.0f;
        const float rsq2 = rcpf(sq * sq);
        float gs_acc = 0.0f, glz_acc = 0.0f;
        for (int64_t sidx = 0; sidx < S; ++sidx) {
            const int64_t row = sidx * B + b;
            RowSlice<G, EPL> zr, s, gsv, gx, gy;
            zr.load(z, row, D, lg, valid);
            const float g = valid ? __ldg(gkl + row) : 0.0f;
            const MAddCtx ma = mobius_add_raw<G, EPL>(xn, zr, s, ball);
            const float r = sqrt_fast(sqnorm<G, EPL>(s));
            const float rz = sqrt_fast(sqnorm<G, EPL>(zr));
            const float dq = 2.0f * ball.rsc * artanh_c(ball.sc * r);
            const float dp = 2.0f * ball.rsc * artanh_c(ball.sc * rz);
            gs_acc += g * dq * dq * rsq2 * rcpf(sq);
            glz_acc -= g;
            // d kl / d s = g * (-dq/sq^2) * 2 artanh'(sc r) * s/r
            const float cq = (r > 0.0f) ? g * (-dq * rsq2) * 2.0f * artanh_grad(ball.sc * r) * rcpf(r) : 0.0f;
#pragma unroll
            for (int i = 0; i < EPL; ++i) gsv.v[i] = cq * s.v[i];
            mobius_add_raw_bwd<G, EPL>(xn, zr, ma, gsv, gx, gy, ball);
            const float cp = (rz > 0.0f) ? g * (dp * rcpf(sp * sp)) * 2.0f * artanh_grad(ball.sc * rz) * rcpf(rz) : 0.0f;
#pragma unroll
            for (int i = 0; i < EPL; ++i) {
                gm.v[i] -= gx.v[i];
                gy.v[i] = fmaf(cp, zr.v[i], gy.v[i]);
            }
            gy.store(gz, row, D, lg, valid);
        }
        gm.store(gmu, b, D, lg, valid);
        if (valid && lg == 0) {
            gsigq[b] = gs_acc;
            glogzq[b] = glz_acc;
        }
    }
}


// ---------------------------------------------------------------------------------------------------
// Fused RiemannianNormal head of the config-2 step (S = 1):  z = expmap_polar(mu, alpha, r),  kl = log q(z) - log p(z),
// and in the backward everything that hangs off (mu, sigma): the KL terms, the sample's gradient through expmap_polar,
// the implicit-reparameterisation term g_r dr/dsigma and the normaliser term -g dlogZ/dsigma.  The unfused graph runs
// k_expmap_polar_fwd + k_rn_kl_fwd forward and k_rn_kl_bwd, k_expmap_polar_bwd plus seven torch elementwise / reduction
// kernels backward (each a launch-latency-bound pass over B or B*D floats); the arithmetic per row is the SAME
// sequence of operations (the two kernels' bodies back to back), so results match the unfused path to rounding of the
// final sums.  reference: old_pvae_riemannian_normal.py:12-52 (rsample, log_prob), training/old_pvae_train.py:53-58.
// ---------------------------------------------------------------------------------------------------
template <int G, int EPL>
__global__ void __launch_bounds__(kRowThreads)
k_rn_head_fwd(const float* __restrict__ mu, const float* __restrict__ alpha, const float* __restrict__ r,
              const float* __restrict__ sigq, const float* __restrict__ logzq, const float* __restrict__ sigp,
              const float* __restrict__ logzp, float* __restrict__ z, float* __restrict__ kl, int64_t B, int D, Ball ball) {
    HVAE_ROW_PROLOGUE(G)
    const float sp = __ldg(sigp), lzp = __ldg(logzp);
    for (int64_t r0 = warp_global * RPW; r0 < B; r0 += warps_total * RPW) {
        const int64_t b = r0 + sub;
        const bool valid = b < B;
        RowSlice<G, EPL> m, a, w, zr, s;
        m.load(mu, b, D, lg, valid);
        a.load(alpha, b, D, lg, valid);
        const float rr = valid ? __ldg(r + b) : 0.0f;
        // expmap_polar (k_expmap_polar_fwd)
        const float an = fmaxf(sqrtf(sqnorm<G, EPL>(a)), kMinNorm);
        const float q = tanh_c(ball.sc * 0.5f * rr) / (ball.sc * an);
#pragma unroll
        for (int i = 0; i < EPL; ++i) w.v[i] = q * a.v[i];
        mobius_add_raw<G, EPL>(m, w, zr, ball);
        float pn;
        project_inplace<G, EPL>(zr, ball, pn);
        zr.store(z, b, D, lg, valid);
        // Monte-Carlo KL (k_rn_kl_fwd)
#pragma unroll
        for (int i = 0; i < EPL; ++i) m.v[i] = -m.v[i];
        mobius_add_raw<G, EPL>(m, zr, s, ball);
        const float rq = sqrt_fast(sqnorm<G, EPL>(s));
        const float rz = sqrt_fast(sqnorm<G, EPL>(zr));
        const float dq = 2.0f * ball.rsc * artanh_c(ball.sc * rq);
        const float dp = 2.0f * ball.rsc * artanh_c(ball.sc * rz);
        if (valid && lg == 0) {
            const float sq = __ldg(sigq + b);
            kl[b] = -dq * dq * (0.5f * rcpf(sq * sq)) - __ldg(logzq + b) + dp * dp * (0.5f * rcpf(sp * sp)) + lzp;
        }
    }
}

template <int G, int EPL>
__global__ void __launch_bounds__(kRowThreads)
k_rn_head_bwd(const float* __restrict__ mu, const float* __restrict__ alpha, const float* __restrict__ r,
              const float* __restrict__ sigq, const float* __restrict__ sigp, const float* __restrict__ z,
              const float* __restrict__ dr_dsigma, const float* __restrict__ dlogz_dsigma,
              const float* __restrict__ gz_in /* may be NULL */, const float* __restrict__ gkl /* may be NULL */,
              float* __restrict__ gmu, float* __restrict__ gsig, int64_t B, int D, Ball ball) {
    HVAE_ROW_PROLOGUE(G)
    const float sp = __ldg(sigp);
    for (int64_t r0 = warp_global * RPW; r0 < B; r0 += warps_total * RPW) {
        const int64_t b = r0 + sub;
        const bool valid = b < B;
        RowSlice<G, EPL> m, xn, zr, s, gsv, gx1, gzt;
        m.load(mu, b, D, lg, valid);
#pragma unroll
        for (int i = 0; i < EPL; ++i) xn.v[i] = -m.v[i];
        zr.load(z, b, D, lg, valid);
        const float sq = valid ? __ldg(sigq + b) : 1.0f;
        const float rsq2 = rcpf(sq * sq);
        const float g = (valid && gkl) ? __ldg(gkl + b) : 0.0f;
        // ---- KL terms (k_rn_kl_bwd, S = 1)
        const MAddCtx ma = mobius_add_raw<G, EPL>(xn, zr, s, ball);
        const float rq = sqrt_fast(sqnorm<G, EPL>(s));
        const float rz = sqrt_fast(sqnorm<G, EPL>(zr));
        const float dq = 2.0f * ball.rsc * artanh_c(ball.sc * rq);
        const float dp = 2.0f * ball.rsc * artanh_c(ball.sc * rz);
        const float gs_kl = g * dq * dq * rsq2 * rcpf(sq);
        const float glz = -g;
        const float cq = (rq > 0.0f) ? g * (-dq * rsq2) * 2.0f * artanh_grad(ball.sc * rq) * rcpf(rq) : 0.0f;
#pragma unroll
        for (int i = 0; i < EPL; ++i) gsv.v[i] = cq * s.v[i];
        mobius_add_raw_bwd<G, EPL>(xn, zr, ma, gsv, gx1, gzt, ball);
        const float cp = (rz > 0.0f) ? g * (dp * rcpf(sp * sp)) * 2.0f * artanh_grad(ball.sc * rz) * rcpf(rz) : 0.0f;
#pragma unroll
        for (int i = 0; i < EPL; ++i) gzt.v[i] = fmaf(cp, zr.v[i], gzt.v[i]);     // d kl / d z
        // ---- total gradient of the sample: decoder's + KL's
        if (gz_in) {
            RowSlice<G, EPL> gd;
            gd.load(gz_in, b, D, lg, valid);
#pragma unroll
            for (int i = 0; i < EPL; ++i) gzt.v[i] = gd.v[i] + gzt.v[i];
        }
        // ---- through z = project(mu (+) w), w = tanh(sc r/2) alpha/(sc |alpha|)  (k_expmap_polar_bwd, S = 1)
        RowSlice<G, EPL> a, w, o, gx2, gw;
        a.load(alpha, b, D, lg, valid);
        const float rr = valid ? __ldg(r + b) : 0.0f;
        const float an = fmaxf(sqrtf(sqnorm<G, EPL>(a)), kMinNorm);
        const float th = ball.sc * 0.5f * rr;
        const float t = tanh_c(th);
        const float q = t / (ball.sc * an);
#pragma unroll
        for (int i = 0; i < EPL; ++i) w.v[i] = q * a.v[i];
        const MAddCtx mb = mobius_add_raw<G, EPL>(m, w, o, ball);
        RowSlice<G, EPL> op = o;
        float pn;
        const bool hit = project_inplace<G, EPL>(op, ball, pn);
        project_bwd<G, EPL>(gzt, o, pn, hit, ball);
        mobius_add_raw_bwd<G, EPL>(m, w, mb, gzt, gx2, gw, ball);
        const float gwa = dot<G, EPL>(gw, a);
        const float g_r = gwa * (1.0f - t * t) * tanh_mask(th) * 0.5f / an;
        RowSlice<G, EPL> gm;
#pragma unroll
        for (int i = 0; i < EPL; ++i) gm.v[i] = -gx1.v[i] + gx2.v[i];
        gm.store(gmu, b, D, lg, valid);
        if (valid && lg == 0)
            gsig[b] = gs_kl + g_r * __ldg(dr_dsigma + b) + glz * __ldg(dlogz_dsigma + b);
    }
}

}  // namespace hvae

using namespace hvae;

extern "C" int hvae_rn_kl_fwd_f32(const float* mu, const float* sigma_q, const float* logz_q, const float* z,
                                  const float* sigma_p, const float* logz_p, float* kl, int64_t S, int64_t B, int64_t D,
                                  float c, void* stream) {
    if (S < 0 || B < 0 || D <= 0 || D > kMaxRowDim) return HVAE_ESHAPE;
    if (S == 0 || B == 0) return HVAE_OK;
    if (!mu || !sigma_q || !logz_q || !z || !sigma_p || !logz_p || !kl) return HVAE_EARG;
    HVAE_ROW_DISPATCH(k_rn_kl_fwd, D, S * B, (cudaStream_t)stream, mu, sigma_q, logz_q, z, sigma_p, logz_p, kl, S, B, (int)D,
                      make_ball(c));
    return check_launch();
}

extern "C" int hvae_rn_kl_bwd_f32(const float* mu, const float* sigma_q, const float* z, const float* sigma_p,
                                  const float* gkl, float* gmu, float* gsigma_q, float* glogz_q, float* gz, int64_t S,
                                  int64_t B, int64_t D, float c, void* stream) {
    if (S < 0 || B < 0 || D <= 0 || D > kMaxRowDim) return HVAE_ESHAPE;
    if (S == 0 || B == 0) return HVAE_OK;
    if (!mu || !sigma_q || !z || !sigma_p || !gkl || !gmu || !gsigma_q || !glogz_q || !gz) return HVAE_EARG;
    HVAE_ROW_DISPATCH(k_rn_kl_bwd, D, B, (cudaStream_t)stream, mu, sigma_q, z, sigma_p, gkl, gmu, gsigma_q, glogz_q, gz, S, B,
                      (int)D, make_ball(c));
    return check_launch();
}

extern "C" int hvae_rn_head_fwd_f32(const float* mu, const float* alpha, const float* r, const float* sigma_q,
                                    const float* logz_q, const float* sigma_p, const float* logz_p, float* z, float* kl,
                                    int64_t B, int64_t D, float c, void* stream) {
    if (B < 0 || D <= 0 || D > kMaxRowDim) return HVAE_ESHAPE;
    if (B == 0) return HVAE_OK;
    if (!mu || !alpha || !r || !sigma_q || !logz_q || !sigma_p || !logz_p || !z || !kl) return HVAE_EARG;
    HVAE_ROW_DISPATCH(k_rn_head_fwd, D, B, (cudaStream_t)stream, mu, alpha, r, sigma_q, logz_q, sigma_p, logz_p, z, kl, B, (int)D,
                      make_ball(c));
    return check_launch();
}

extern "C" int hvae_rn_head_bwd_f32(const float* mu, const float* alpha, const float* r, const float* sigma_q,
                                    const float* sigma_p, const float* z, const float* dr_dsigma, const float* dlogz_dsigma,
                                    const float* gz, const float* gkl, float* gmu, float* gsigma, int64_t B, int64_t D,
                                    float c, void* stream) {
    if (B < 0 || D <= 0 || D > kMaxRowDim) return HVAE_ESHAPE;
    if (B == 0) return HVAE_OK;
    if (!mu || !alpha || !r || !sigma_q || !sigma_p || !z || !dr_dsigma || !dlogz_dsigma || !gmu || !gsigma) return HVAE_EARG;
    HVAE_ROW_DISPATCH(k_rn_head_bwd, D, B, (cudaStream_t)stream, mu, alpha, r, sigma_q, sigma_p, z, dr_dsigma, dlogz_dsigma, gz, gkl,
                      gmu, gsigma, B, (int)D, make_ball(c));
    return check_launch();
}
