// K3 and friends: single-pass fused row kernels for the Poincare-ball point/vector maps.
// One row (length D) is owned by G lanes of a warp and lives in registers; norms are warp-shuffle
// reductions; every kernel reads each input once and writes each output once (HBM-bound).
//
// Algorithmic bytes per row (fp32): expmap0/logmap0 fwd 8D, bwd 12D; mobius_add fwd 12D, bwd 20D.
#include "hvae_common.cuh"
#include "row_maps.cuh"

namespace hvae {

template <int G, int EPL>
__global__ void __launch_bounds__(kRowThreads) k_expmap0_fwd(const float* __restrict__ u, float* __restrict__ y,
                                                              int64_t rows, int D, Ball ball) {
    HVAE_ROW_PROLOGUE(G)
    constexpr int U = HVAE_ROW_UNROLL(EPL);
    for (int64_t r0 = warp_global * (RPW * U); r0 < rows; r0 += warps_total * (RPW * U)) {
        RowSlice<G, EPL> ur[U];
#pragma unroll
        for (int j = 0; j < U; ++j) ur[j].load(u, r0 + j * RPW + sub, D, lg, r0 + j * RPW + sub < rows);
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const int64_t row = r0 + j * RPW + sub;
            RowSlice<G, EPL> yr;
            float n_raw, n, t, pn;
            expmap0_row<G, EPL>(ur[j], yr, ball, n_raw, n, t);
            project_inplace<G, EPL>(yr, ball, pn);
            yr.store(y, row, D, lg, row < rows);
        }
    }
}

template <int G, int EPL>
__global__ void __launch_bounds__(kRowThreads) k_expmap0_bwd(const float* __restrict__ u, const float* __restrict__ gy,
                                                              float* __restrict__ gu, int64_t rows, int D, Ball ball) {
    HVAE_ROW_PROLOGUE(G)
    constexpr int U = HVAE_ROW_UNROLL(EPL);
    for (int64_t r0 = warp_global * (RPW * U); r0 < rows; r0 += warps_total * (RPW * U)) {
        RowSlice<G, EPL> ur[U], g[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const int64_t row = r0 + j * RPW + sub;
            ur[j].load(u, row, D, lg, row < rows);
            g[j].load(gy, row, D, lg, row < rows);
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const int64_t row = r0 + j * RPW + sub;
            expmap0_row_bwd<G, EPL>(ur[j], g[j], ball);
            g[j].store(gu, row, D, lg, row < rows);
        }
    }
}

// =================================================================================================
// logmap0:  u = y/n * artanh(clamp(sc*n)) / sc   (no projection)
// reference: geoopt logmap0 via hyperbolic_vae/models/vae_one_b.py:218,226-227
// =================================================================================================
template <int G, int EPL>
__global__ void __launch_bounds__(kRowThreads) k_logmap0_fwd(const float* __restrict__ y, float* __restrict__ u,
                                                              int64_t rows, int D, Ball ball) {
    HVAE_ROW_PROLOGUE(G)
    for (int64_t r0 = warp_global * RPW; r0 < rows; r0 += warps_total * RPW) {
        const int64_t row = r0 + sub;
        const bool valid = row < rows;
        RowSlice<G, EPL> yr;
        yr.load(y, row, D, lg, valid);
        const float n = fmaxf(sqrt_fast(sqnorm<G, EPL>(yr)), kMinNorm);
        const float at = ball.rsc * artanh_c(ball.sc * n) * rcpf(n);
#pragma unroll
        for (int i = 0; i < EPL; ++i) yr.v[i] *= at;
        yr.store(u, row, D, lg, valid);
    }
}

template <int G, int EPL>
__device__ __forceinline__ void logmap0_row_bwd(const RowSlice<G, EPL>& y, RowSlice<G, EPL>& g, const Ball& ball) {
    const float n_raw = sqrt_fast(sqnorm<G, EPL>(y));
    const float n = fmaxf(n_raw, kMinNorm);
    const float a = ball.sc * n;
    const float h = artanh_c(a) * rcpf(a);              // artanh(sc n)/(sc n)
    const float hp = (artanh_grad(a) - h) * rcpf(n);     // h'(n)
    const float gdot = dot<G, EPL>(g, y);
    const float coef = (n_raw >= kMinNorm) ? hp * gdot * rcpf(n_raw) : 0.0f;
#pragma unroll
    for (int i = 0; i < EPL; ++i) g.v[i] = h * g.v[i] + coef * y.v[i];
}

template <int G, int EPL>
__global__ void __launch_bounds__(kRowThreads) k_logmap0_bwd(const float* __restrict__ y, const float* __restrict__ gu,
                                                              float* __restrict__ gy, int64_t rows, int D, Ball ball) {
    HVAE_ROW_PROLOGUE(G)
    for (int64_t r0 = warp_global * RPW; r0 < rows; r0 += warps_total * RPW) {
        const int64_t row = r0 + sub;
        const bool valid = row < rows;
        RowSlice<G, EPL> yr, g;
        yr.load(y, row, D, lg, valid);
        g.load(gu, row, D, lg, valid);
        logmap0_row_bwd<G, EPL>(yr, g, ball);
        g.store(gy, row, D, lg, valid);
    }
}

// =================================================================================================
// mobius_add(x, y) [+ project]
// =================================================================================================
template <int G, int EPL>
__global__ void __launch_bounds__(kRowThreads) k_mobius_add_fwd(const float* __restrict__ x, const float* __restrict__ y,
                                                                 float* __restrict__ out, int64_t rows, int D, Ball ball,
                                                                 int project) {
    HVAE_ROW_PROLOGUE(G)
    for (int64_t r0 = warp_global * RPW; r0 < rows; r0 += warps_total * RPW) {
        const int64_t row = r0 + sub;
        const bool valid = row < rows;
        RowSlice<G, EPL> xr, yr, o;
        xr.load(x, row, D, lg, valid);
        yr.load(y, row, D, lg, valid);
        mobius_add_raw<G, EPL>(xr, yr, o, ball);
        float pn;
        if (project) project_inplace<G, EPL>(o, ball, pn);
        o.store(out, row, D, lg, valid);
    }
}

template <int G, int EPL>
__global__ void __launch_bounds__(kRowThreads) k_mobius_add_bwd(const float* __restrict__ x, const float* __restrict__ y,
                                                                 const float* __restrict__ gout, float* __restrict__ gx,
                                                                 float* __restrict__ gy, int64_t rows, int D, Ball ball,
                                                                 int project) {
    HVAE_ROW_PROLOGUE(G)
    for (int64_t r0 = warp_global * RPW; r0 < rows; r0 += warps_total * RPW) {
        const int64_t row = r0 + sub;
        const bool valid = row < rows;
        RowSlice<G, EPL> xr, yr, g, o, gxr, gyr;
        xr.load(x, row, D, lg, valid);
        yr.load(y, row, D, lg, valid);
        g.load(gout, row, D, lg, valid);
        const MAddCtx m = mobius_add_raw<G, EPL>(xr, yr, o, ball);
        if (project) {
            RowSlice<G, EPL> op = o;
            float pn;
            const bool hit = project_inplace<G, EPL>(op, ball, pn);
            project_bwd<G, EPL>(g, o, pn, hit, ball);
        }
        mobius_add_raw_bwd<G, EPL>(xr, yr, m, g, gxr, gyr, ball);
        gxr.store(gx, row, D, lg, valid);
        gyr.store(gy, row, D, lg, valid);
    }
}

// =================================================================================================
// expmap(x, u) = project( x (+) tanh(sc * lambda_x/2 * ||u||)/sc * u/||u|| )
// logmap(x, y) = 2 artanh(sc ||s||)/sc * s / (lambda_x ||s||),  s = (-x) (+) y
// dist(x, y)   = 2 artanh(sc ||(-x) (+) y||)/sc
// reference: geoopt PoincareBall.expmap/logmap/dist via wrapped_normal.py:73,83; manifolds.py:31
// =================================================================================================
template <int G, int EPL>
struct ExpmapCtx {
    RowSlice<G, EPL> w;  // second term
    float x2, m, lam, un_raw, un, th, t, sech2;
    bool m_clamped;
    MAddCtx ma;
};

template <int G, int EPL>
__device__ __forceinline__ void expmap_row(const RowSlice<G, EPL>& x, const RowSlice<G, EPL>& u, RowSlice<G, EPL>& out,
                                           ExpmapCtx<G, EPL>& k, const Ball& ball) {
    k.x2 = sqnorm<G, EPL>(x);
    const float m = 1.0f - ball.c * k.x2;
    k.m_clamped = m < kMinNorm;
    k.m = fmaxf(m, kMinNorm);
    k.lam = 2.0f * rcpf(k.m);
    k.un_raw = sqrt_fast(sqnorm<G, EPL>(u));
    k.un = fmaxf(k.un_raw, kMinNorm);
    k.th = ball.sc * ((k.lam * 0.5f) * k.un);
    tanh_sech2(k.th, k.t, k.sech2);
    const float q = ball.rsc * k.t * rcpf(k.un);
#pragma unroll
    for (int i = 0; i < EPL; ++i) k.w.v[i] = q * u.v[i];
    k.ma = mobius_add_raw<G, EPL>(x, k.w, out, ball);
}

// g: in dL/d(pre-projection out) ; produces gx, gu
template <int G, int EPL>
__device__ __forceinline__ void expmap_row_bwd(const RowSlice<G, EPL>& x, const RowSlice<G, EPL>& u,
                                               const ExpmapCtx<G, EPL>& k, const RowSlice<G, EPL>& g,
                                               RowSlice<G, EPL>& gx, RowSlice<G, EPL>& gu, const Ball& ball) {
    RowSlice<G, EPL> gw;
    mobius_add_raw_bwd<G, EPL>(x, k.w, k.ma, g, gx, gw, ball);
    // w = q(un, lam) u,  q = tanh(th)/(sc un),  th = sc lam un / 2
    const float sech2 = k.sech2;
    const float run = rcpf(k.un);
    const float q = ball.rsc * k.t * run;
    const float dq_dun = (sech2 * (k.lam * 0.5f) - q) * run;   // sech^2*th' /(sc un) - tanh/(sc un^2)
    const float dq_dlam = sech2 * 0.5f;                          // sech^2 * (sc un/2) / (sc un)
    const float gwu = dot<G, EPL>(gw, u);
    const float coef = (k.un_raw >= kMinNorm) ? dq_dun * gwu * rcpf(k.un_raw) : 0.0f;
#pragma unroll
    for (int i = 0; i < EPL; ++i) gu.v[i] = q * gw.v[i] + coef * u.v[i];
    // lam = 2/m, m = clamp_min(1 - c x2): d lam/d x = (2/m^2) * 2 c x  when unclamped
    const float glam = dq_dlam * gwu;
    const float rm = rcpf(k.m);
    const float gx2 = k.m_clamped ? 0.0f : glam * (2.0f * rm * rm) * ball.c;
#pragma unroll
    for (int i = 0; i < EPL; ++i) gx.v[i] += 2.0f * gx2 * x.v[i];
}

template <int G, int EPL>
__global__ void __launch_bounds__(kRowThreads) k_expmap_fwd(const float* __restrict__ x, const float* __restrict__ u,
                                                             float* __restrict__ out, int64_t rows, int D, Ball ball) {
    HVAE_ROW_PROLOGUE(G)
    for (int64_t r0 = warp_global * RPW; r0 < rows; r0 += warps_total * RPW) {
        const int64_t row = r0 + sub;
        const bool valid = row < rows;
        RowSlice<G, EPL> xr, ur, o;
        xr.load(x, row, D, lg, valid);
        ur.load(u, row, D, lg, valid);
        ExpmapCtx<G, EPL> k;
        expmap_row<G, EPL>(xr, ur, o, k, ball);
        float pn;
        project_inplace<G, EPL>(o, ball, pn);
        o.store(out, row, D, lg, valid);
    }
}

template <int G, int EPL>
__global__ void __launch_bounds__(kRowThreads) k_expmap_bwd(const float* __restrict__ x, const float* __restrict__ u,
                                                             const float* __restrict__ gout, float* __restrict__ gx,
                                                             float* __restrict__ gu, int64_t rows, int D, Ball ball) {
    HVAE_ROW_PROLOGUE(G)
    for (int64_t r0 = warp_global * RPW; r0 < rows; r0 += warps_total * RPW) {
        const int64_t row = r0 + sub;
        const bool valid = row < rows;
        RowSlice<G, EPL> xr, ur, g, o, gxr, gur;
        xr.load(x, row, D, lg, valid);
        ur.load(u, row, D, lg, valid);
        g.load(gout, row, D, lg, valid);
        ExpmapCtx<G, EPL> k;
        expmap_row<G, EPL>(xr, ur, o, k, ball);
        RowSlice<G, EPL> op = o;
        float pn;
        const bool hit = project_inplace<G, EPL>(op, ball, pn);
        project_bwd<G, EPL>(g, o, pn, hit, ball);
        expmap_row_bwd<G, EPL>(xr, ur, k, g, gxr, gur, ball);
        gxr.store(gx, row, D, lg, valid);
        gur.store(gu, row, D, lg, valid);
    }
}

// logmap / dist share the (-x)(+)y subtraction
template <int G, int EPL, bool kDist>
__global__ void __launch_bounds__(kRowThreads) k_logmap_fwd(const float* __restrict__ x, const float* __restrict__ y,
                                                             float* __restrict__ out, int64_t rows, int D, Ball ball) {
    HVAE_ROW_PROLOGUE(G)
    for (int64_t r0 = warp_global * RPW; r0 < rows; r0 += warps_total * RPW) {
        const int64_t row = r0 + sub;
        const bool valid = row < rows;
        RowSlice<G, EPL> xr, yr, s;
        xr.load(x, row, D, lg, valid);
        yr.load(y, row, D, lg, valid);
#pragma unroll
        for (int i = 0; i < EPL; ++i) xr.v[i] = -xr.v[i];
        const MAddCtx ma = mobius_add_raw<G, EPL>(xr, yr, s, ball);
        const float r = sqrt_fast(sqnorm<G, EPL>(s));
        if (kDist) {
            const float d = 2.0f * (ball.rsc * artanh_c(ball.sc * r));
            if (valid && lg == 0) out[row] = d;
        } else {
            const float sn = fmaxf(r, kMinNorm);
            const float inv_lam = 0.5f * fmaxf(1.0f - ball.c * ma.x2, kMinNorm);
            const float f = 2.0f * (ball.rsc * artanh_c(ball.sc * sn)) * inv_lam * rcpf(sn);
#pragma unroll
            for (int i = 0; i < EPL; ++i) s.v[i] *= f;
            s.store(out, row, D, lg, valid);
        }
    }
}

template <int G, int EPL, bool kDist>
__global__ void __launch_bounds__(kRowThreads) k_logmap_bwd(const float* __restrict__ x, const float* __restrict__ y,
                                                             const float* __restrict__ gout, float* __restrict__ gx,
                                                             float* __restrict__ gy, int64_t rows, int D, Ball ball) {
    HVAE_ROW_PROLOGUE(G)
    for (int64_t r0 = warp_global * RPW; r0 < rows; r0 += warps_total * RPW) {
        const int64_t row = r0 + sub;
        const bool valid = row < rows;
        RowSlice<G, EPL> xr, yr, s, gs, gxr, gyr;
        xr.load(x, row, D, lg, valid);
        yr.load(y, row, D, lg, valid);
#pragma unroll
        for (int i = 0; i < EPL; ++i) xr.v[i] = -xr.v[i];
        const MAddCtx ma = mobius_add_raw<G, EPL>(xr, yr, s, ball);
        const float r = sqrt_fast(sqnorm<G, EPL>(s));
        float glam = 0.0f, m = 1.0f;
        bool m_clamped = false;
        if (kDist) {
            const float gd = valid ? __ldg(gout + row) : 0.0f;
            // d = 2 artanh(sc r)/sc ; dd/dr = 2 * artanh'(sc r); d r/d s = s/r (0 at r = 0)
            const float coef = (r > 0.0f) ? gd * 2.0f * artanh_grad(ball.sc * r) * rcpf(r) : 0.0f;
#pragma unroll
            for (int i = 0; i < EPL; ++i) gs.v[i] = coef * s.v[i];
        } else {
            RowSlice<G, EPL> g;
            g.load(gout, row, D, lg, valid);
            const float sn = fmaxf(r, kMinNorm);
            const float mm = 1.0f - ball.c * ma.x2;
            m_clamped = mm < kMinNorm;
            m = fmaxf(mm, kMinNorm);
            const float inv_lam = 0.5f * m;
            // out = phi(sn)/lam * s,  phi = 2 artanh(sc sn)/(sc sn)
            const float a = ball.sc * sn;
            const float phi = 2.0f * artanh_c(a) * rcpf(a);
            const float dphi = (2.0f * artanh_grad(a) - phi) * rcpf(sn);
            const float gds = dot<G, EPL>(g, s);
            const float coef = (r >= kMinNorm) ? dphi * gds * inv_lam * rcpf(r) : 0.0f;
#pragma unroll
            for (int i = 0; i < EPL; ++i) gs.v[i] = (phi * inv_lam) * g.v[i] + coef * s.v[i];
            glam = -phi * gds * inv_lam * inv_lam;
        }
        mobius_add_raw_bwd<G, EPL>(xr, yr, ma, gs, gxr, gyr, ball);
        // x entered negated; lambda depends on ||x||^2
        const float gx2 = (kDist || m_clamped) ? 0.0f : glam * (2.0f * rcpf(m * m)) * ball.c;
#pragma unroll
        for (int i = 0; i < EPL; ++i) gxr.v[i] = -gxr.v[i] + 2.0f * gx2 * (-xr.v[i]);
        gxr.store(gx, row, D, lg, valid);
        gyr.store(gy, row, D, lg, valid);
    }
}

}  // namespace hvae

// =================================================================================================
// C ABI
// =================================================================================================
using namespace hvae;

#define HVAE_CHECK_ROW_ARGS(rows, D)                                   \
    if ((rows) < 0 || (D) <= 0 || (D) > kMaxRowDim) return HVAE_ESHAPE; \
    if ((rows) == 0) return HVAE_OK;

extern "C" int hvae_expmap0_fwd_f32(const float* u, float* y, int64_t rows, int64_t D, float c, void* stream) {
    HVAE_CHECK_ROW_ARGS(rows, D)
    if (!u || !y) return HVAE_EARG;
    HVAE_ROW_DISPATCH(k_expmap0_fwd, D, rows, (cudaStream_t)stream, u, y, rows, (int)D, make_ball(c));
    return check_launch();
}

extern "C" int hvae_expmap0_bwd_f32(const float* u, const float* gy, float* gu, int64_t rows, int64_t D, float c,
                                    void* stream) {
    HVAE_CHECK_ROW_ARGS(rows, D)
    if (!u || !gy || !gu) return HVAE_EARG;
    HVAE_ROW_DISPATCH(k_expmap0_bwd, D, rows, (cudaStream_t)stream, u, gy, gu, rows, (int)D, make_ball(c));
    return check_launch();
}

extern "C" int hvae_logmap0_fwd_f32(const float* y, float* u, int64_t rows, int64_t D, float c, void* stream) {
    HVAE_CHECK_ROW_ARGS(rows, D)
    if (!u || !y) return HVAE_EARG;
    HVAE_ROW_DISPATCH(k_logmap0_fwd, D, rows, (cudaStream_t)stream, y, u, rows, (int)D, make_ball(c));
    return check_launch();
}

extern "C" int hvae_logmap0_bwd_f32(const float* y, const float* gu, float* gy, int64_t rows, int64_t D, float c,
                                    void* stream) {
    HVAE_CHECK_ROW_ARGS(rows, D)
    if (!y || !gy || !gu) return HVAE_EARG;
    HVAE_ROW_DISPATCH(k_logmap0_bwd, D, rows, (cudaStream_t)stream, y, gu, gy, rows, (int)D, make_ball(c));
    return check_launch();
}

extern "C" int hvae_mobius_add_fwd_f32(const float* x, const float* y, float* out, int64_t rows, int64_t D, float c,
                                       int project, void* stream) {
    HVAE_CHECK_ROW_ARGS(rows, D)
    if (!x || !y || !out) return HVAE_EARG;
    HVAE_ROW_DISPATCH(k_mobius_add_fwd, D, rows, (cudaStream_t)stream, x, y, out, rows, (int)D, make_ball(c), project);
    return check_launch();
}

extern "C" int hvae_mobius_add_bwd_f32(const float* x, const float* y, const float* gout, float* gx, float* gy,
                                       int64_t rows, int64_t D, float c, int project, void* stream) {
    HVAE_CHECK_ROW_ARGS(rows, D)
    if (!x || !y || !gout || !gx || !gy) return HVAE_EARG;
    HVAE_ROW_DISPATCH(k_mobius_add_bwd, D, rows, (cudaStream_t)stream, x, y, gout, gx, gy, rows, (int)D, make_ball(c),
                      project);
    return check_launch();
}

extern "C" int hvae_expmap_fwd_f32(const float* x, const float* u, float* out, int64_t rows, int64_t D, float c,
                                   void* stream) {
    HVAE_CHECK_ROW_ARGS(rows, D)
    if (!x || !u || !out) return HVAE_EARG;
    HVAE_ROW_DISPATCH(k_expmap_fwd, D, rows, (cudaStream_t)stream, x, u, out, rows, (int)D, make_ball(c));
    return check_launch();
}

extern "C" int hvae_expmap_bwd_f32(const float* x, const float* u, const float* gout, float* gx, float* gu,
                                   int64_t rows, int64_t D, float c, void* stream) {
    HVAE_CHECK_ROW_ARGS(rows, D)
    if (!x || !u || !gout || !gx || !gu) return HVAE_EARG;
    HVAE_ROW_DISPATCH(k_expmap_bwd, D, rows, (cudaStream_t)stream, x, u, gout, gx, gu, rows, (int)D, make_ball(c));
    return check_launch();
}

extern "C" int hvae_logmap_fwd_f32(const float* x, const float* y, float* out, int64_t rows, int64_t D, float c,
                                   void* stream) {
    HVAE_CHECK_ROW_ARGS(rows, D)
    if (!x || !y || !out) return HVAE_EARG;
    cudaStream_t s = (cudaStream_t)stream;
    const Ball ball = make_ball(c);
    const int d = (int)D;
#define L_(G_, E_) k_logmap_fwd<G_, E_, false><<<row_grid(rows, G_), kRowThreads, 0, s>>>(x, y, out, rows, d, ball)
    if (D <= 2) L_(1, 2); else if (D <= 4) L_(1, 4); else if (D <= 8) L_(1, 8); else if (D <= 16) L_(2, 8);
    else if (D <= 32) L_(4, 8); else if (D <= 64) L_(8, 8); else if (D <= 128) L_(16, 8); else if (D <= 256) L_(32, 8);
    else if (D <= 512) L_(32, 16); else L_(32, 32);
#undef L_
    return check_launch();
}

extern "C" int hvae_dist_fwd_f32(const float* x, const float* y, float* dd, int64_t rows, int64_t D, float c,
                                 void* stream) {
    HVAE_CHECK_ROW_ARGS(rows, D)
    if (!x || !y || !dd) return HVAE_EARG;
    cudaStream_t s = (cudaStream_t)stream;
    const Ball ball = make_ball(c);
    const int d = (int)D;
#define L_(G_, E_) k_logmap_fwd<G_, E_, true><<<row_grid(rows, G_), kRowThreads, 0, s>>>(x, y, dd, rows, d, ball)
    if (D <= 2) L_(1, 2); else if (D <= 4) L_(1, 4); else if (D <= 8) L_(1, 8); else if (D <= 16) L_(2, 8);
    else if (D <= 32) L_(4, 8); else if (D <= 64) L_(8, 8); else if (D <= 128) L_(16, 8); else if (D <= 256) L_(32, 8);
    else if (D <= 512) L_(32, 16); else L_(32, 32);
#undef L_
    return check_launch();
}

extern "C" int hvae_logmap_bwd_f32(const float* x, const float* y, const float* gout, float* gx, float* gy,
                                   int64_t rows, int64_t D, float c, void* stream) {
    HVAE_CHECK_ROW_ARGS(rows, D)
    if (!x || !y || !gout || !gx || !gy) return HVAE_EARG;
    cudaStream_t s = (cudaStream_t)stream;
    const Ball ball = make_ball(c);
    const int d = (int)D;
#define L_(G_, E_) \
    k_logmap_bwd<G_, E_, false><<<row_grid(rows, G_), kRowThreads, 0, s>>>(x, y, gout, gx, gy, rows, d, ball)
    if (D <= 2) L_(1, 2); else if (D <= 4) L_(1, 4); else if (D <= 8) L_(1, 8); else if (D <= 16) L_(2, 8);
    else if (D <= 32) L_(4, 8); else if (D <= 64) L_(8, 8); else if (D <= 128) L_(16, 8); else if (D <= 256) L_(32, 8);
    else if (D <= 512) L_(32, 16); else L_(32, 32);
#undef L_
    return check_launch();
}

extern "C" int hvae_dist_bwd_f32(const float* x, const float* y, const float* gd, float* gx, float* gy, int64_t rows,
                                 int64_t D, float c, void* stream) {
    HVAE_CHECK_ROW_ARGS(rows, D)
    if (!x || !y || !gd || !gx || !gy) return HVAE_EARG;
    cudaStream_t s = (cudaStream_t)stream;
    const Ball ball = make_ball(c);
    const int d = (int)D;
#define L_(G_, E_) \
    k_logmap_bwd<G_, E_, true><<<row_grid(rows, G_), kRowThreads, 0, s>>>(x, y, gd, gx, gy, rows, d, ball)
    if (D <= 2) L_(1, 2); else if (D <= 4) L_(1, 4); else if (D <= 8) L_(1, 8); else if (D <= 16) L_(2, 8);
    else if (D <= 32) L_(4, 8); else if (D <= 64) L_(8, 8); else if (D <= 128) L_(16, 8); else if (D <= 256) L_(32, 8);
    else if (D <= 512) L_(32, 16); else L_(32, 32);
#undef L_
    return check_launch();
}
