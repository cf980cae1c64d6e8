// Per-row scalar math of the Mobius matvec y = project(psi(|x|,|mx|) mx) and of its backward, shared by the SIMT
// kernels (mobius.cu) and the tensor-core path (tc_gemm.cu).  reference: geoopt mobius_matvec + project (App. A.1).
#pragma once
#include "hvae_common.cuh"

namespace hvae {

// per-row scalars of y = psi * mx (pre-projection), with derivative pieces
struct MobRow {
    float xn_raw, xn, mxn_raw, mxn, ax, at, kappa, theta, t, psi, ypn;
    bool zero_row, hit;
};

__device__ __forceinline__ void mob_row_scalars(float x2, float mx2, bool all_zero, const Ball& ball, MobRow& r) {
    r.xn_raw = sqrtf(x2);
    r.xn = fmaxf(r.xn_raw, kMinNorm);
    r.mxn_raw = sqrtf(mx2);
    r.mxn = fmaxf(r.mxn_raw, kMinNorm);
    r.ax = ball.sc * r.xn;
    r.at = artanh_c(r.ax);                // artanh(clamp(sc xn)); artan_k = at/sc
    r.kappa = r.at / r.xn;                // theta = sc * (mxn/xn * at/sc) = mxn * at / xn
    r.theta = r.mxn / r.xn * r.at;
    r.t = tanh_c(r.theta);
    r.psi = ball.rsc * r.t / r.mxn;       // y = psi * mx
    r.zero_row = all_zero;
    const float yn = fmaxf(ball.rsc * r.t * (r.mxn_raw / r.mxn), kMinNorm);  // |y_pre| (= t/sc unless mxn was clamped)
    r.ypn = yn;
    r.hit = (!all_zero) && (yn > ball.maxnorm);
}

// backward coefficients of one row:  gmx_j = alpha * gy_j + beta * mx_j ;  gx += gxc * x
//   gdm = <gy, mx>
__device__ __forceinline__ void mob_bwd_coefs(float x2, float mx2, float gdm, bool all_zero, const Ball& ball, float& alpha,
                                              float& beta, float& gxc) {
    MobRow rs;
    mob_row_scalars(x2, mx2, all_zero, ball, rs);
    if (rs.zero_row) { alpha = beta = gxc = 0.0f; return; }
    // projection backward on y_pre = psi*mx (norm ypn):  g' = s (g - (<g,ypre>/ypn^2) ypre)
    float a_ = 1.0f, sub = 0.0f;  // g'_j = a_ * gy_j - sub * mx_j
    float gdm_p = gdm;            // <g', mx>
    if (rs.hit) {
        const float s = ball.maxnorm / rs.ypn;
        const float q = rs.psi * gdm / (rs.ypn * rs.ypn);  // <g,ypre>/ypn^2, ypre = psi mx
        a_ = s;
        sub = s * q * rs.psi;
        gdm_p = s * gdm - sub * mx2;
    }
    // y_pre = psi(xn, mxn) mx
    float tt, sech2;
    tanh_sech2(rs.theta, tt, sech2);
    const float dpsi_dmxn = (sech2 * rs.kappa - rs.t / rs.mxn) * ball.rsc / rs.mxn;
    // kappa'(xn) = (artanh'(sc xn) sc xn - artanh(sc xn)) / xn^2
    const float dkappa = (artanh_grad(rs.ax) * rs.ax - rs.at) / (rs.xn * rs.xn);
    const float dpsi_dxn = sech2 * dkappa * ball.rsc;
    const float cm = (rs.mxn_raw >= kMinNorm) ? dpsi_dmxn * gdm_p / rs.mxn_raw : 0.0f;
    alpha = rs.psi * a_;
    beta = -rs.psi * sub + cm;
    gxc = (rs.xn_raw >= kMinNorm) ? dpsi_dxn * gdm_p / rs.xn_raw : 0.0f;
}

}  // namespace hvae
