// Row-level device functions shared by several translation units (expmap0 and its backward).
#pragma once
#include "hvae_common.cuh"

namespace hvae {

// =================================================================================================
// expmap0:  y = project( tanh(clamp(sc*n)) / sc * u / n ),  n = clamp_min(||u||, 1e-15)
// reference: geoopt expmap0 via hyperbolic_vae/layers.py:129-130
// =================================================================================================
template <int G, int EPL>
__device__ __forceinline__ void expmap0_row(const RowSlice<G, EPL>& u, RowSlice<G, EPL>& y, const Ball& ball,
                                            float& n_raw, float& n, float& t) {
    n_raw = sqrt_fast(sqnorm<G, EPL>(u));
    n = fmaxf(n_raw, kMinNorm);
    t = tanh_c(ball.sc * n);
    const float f = ball.rsc * t * rcpf(n);
#pragma unroll
    for (int i = 0; i < EPL; ++i) y.v[i] = f * u.v[i];
}

// backward through project then through f(n) u
template <int G, int EPL>
__device__ __forceinline__ void expmap0_row_bwd(const RowSlice<G, EPL>& u, RowSlice<G, EPL>& g /*in: gy, out: gu*/,
                                                const Ball& ball) {
    RowSlice<G, EPL> ypre;
    float n_raw, n, t, pn;
    expmap0_row<G, EPL>(u, ypre, ball, n_raw, n, t);
    RowSlice<G, EPL> yproj = ypre;
    const bool hit = project_inplace<G, EPL>(yproj, ball, pn);
    project_bwd<G, EPL>(g, ypre, pn, hit, ball);
    // ypre = f(n) u, f = tanh(sc n)/(sc n);  f'(n) = (sech^2(sc n) 1{|sc n|<=15} - tanh(sc n)/(sc n)) / n
    const float a = ball.sc * n;
    float t2, sech2;
    tanh_sech2(a, t2, sech2);
    const float f = t * rcpf(a);
    const float fp = (sech2 - f) * rcpf(n);
    const float gu_dot = dot<G, EPL>(g, u);
    // d n / d u = u/||u|| where the clamp_min is inactive and ||u|| > 0, else 0
    const float coef = (n_raw >= kMinNorm) ? fp * gu_dot * rcpf(n_raw) : 0.0f;
#pragma unroll
    for (int i = 0; i < EPL; ++i) g.v[i] = f * g.v[i] + coef * u.v[i];
}


}  // namespace hvae
