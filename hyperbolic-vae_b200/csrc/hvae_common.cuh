// Shared device helpers for the Poincare-ball row kernels (sm_100a).
//
// Clamp/eps constants follow geoopt's stereographic math exactly:
//   MIN_NORM 1e-15, fp32 ball eps 4e-3, artanh clamp 1-1e-7 (0.99999988 in fp32), tanh clamp +-15.
// Reference call sites: hyperbolic_vae/layers.py:60,67,130,146; distributions/wrapped_normal.py:66-89.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/hvae_b200.h"

namespace hvae {

constexpr float kMinNorm = 1e-15f;
constexpr float kArtanhClamp = 0.99999988079071044921875f;  // float(1 - 1e-7)
constexpr float kTanhClamp = 15.0f;
constexpr int kNumSMs = 148;

struct Ball {
    float c;        // curvature magnitude (k = -c)
    float sc;       // sqrt(c)
    float rsc;      // 1/sqrt(c)
    float maxnorm;  // (1 - 4e-3)/sqrt(c): fp32 projection radius
};

__host__ __device__ inline Ball make_ball(float c) {
    Ball b;
    b.c = c;
    b.sc = sqrtf(c + 1e-15f);
    b.rsc = 1.0f / b.sc;
    b.maxnorm = 0.996f / b.sc;
    return b;
}

// ---- lane-group reductions: a row is owned by G consecutive lanes of a warp -----------------
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int G>
__device__ __forceinline__ bool group_all(bool p) {
    if (G == 1) return p;
    unsigned m = __ballot_sync(0xffffffffu, p);
    const unsigned lane = threadIdx.x & 31u;
    const unsigned base = lane & ~(unsigned)(G - 1);
    const unsigned gm = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << base;
    return (m & gm) == gm;
}

// ---- a row slice held in registers: element idx = lg + i*G for i < EPL ------------------------------
template <int G, int EPL>
struct RowSlice {
    float v[EPL];

    __device__ __forceinline__ void load(const float* __restrict__ base, int64_t row, int D, int lg, bool valid) {
        const float* p = base + row * (int64_t)D;
#pragma unroll
        for (int i = 0; i < EPL; ++i) {
            const int idx = lg + i * G;
            v[i] = (valid && idx < D) ? __ldg(p + idx) : 0.0f;
        }
    }
    __device__ __forceinline__ void store(float* __restrict__ base, int64_t row, int D, int lg, bool valid) const {
        float* p = base + row * (int64_t)D;
#pragma unroll
        for (int i = 0; i < EPL; ++i) {
            const int idx = lg + i * G;
            if (valid && idx < D) p[idx] = v[i];
        }
    }
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int i = 0; i < EPL; ++i) v[i] = 0.0f;
    }
};

template <int G, int EPL>
__device__ __forceinline__ float dot(const RowSlice<G, EPL>& a, const RowSlice<G, EPL>& b) {
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < EPL; ++i) s = fmaf(a.v[i], b.v[i], s);
    return group_sum<G>(s);
}

template <int G, int EPL>
__device__ __forceinline__ float sqnorm(const RowSlice<G, EPL>& a) {
    return dot<G, EPL>(a, a);
}

// ---- scalar functions with the reference's clamps, plus the derivative masks ----------------------
__device__ __forceinline__ float tanh_c(float x) { return tanhf(fminf(fmaxf(x, -kTanhClamp), kTanhClamp)); }
__device__ __forceinline__ float tanh_mask(float x) { return (x >= -kTanhClamp && x <= kTanhClamp) ? 1.0f : 0.0f; }

__device__ __forceinline__ float artanh_c(float x) { return atanhf(fminf(fmaxf(x, -kArtanhClamp), kArtanhClamp)); }
// d/dx artanh(clamp(x)) = 1/(1-xc^2) inside the clamp, 0 outside
__device__ __forceinline__ float artanh_grad(float x) {
    if (!(x >= -kArtanhClamp && x <= kArtanhClamp)) return 0.0f;
    return 1.0f / ((1.0f - x) * (1.0f + x));
}

// log(sinh(x)/x) for x >= 0, accurate at small x (the reference's log sinh - log x cancels there)
__device__ __forceinline__ float log_sinhc(float x) {
    if (x < 0.5f) {
        const float x2 = x * x;
        // log(sinh x / x) = x^2/6 - x^4/180 + x^6/2835 - x^8/37800
        return x2 * (1.0f / 6.0f + x2 * (-1.0f / 180.0f + x2 * (1.0f / 2835.0f + x2 * (-1.0f / 37800.0f))));
    }
    // log sinh x = x + log(1 - e^{-2x}) - log 2
    return x + log1pf(-expf(-2.0f * x)) - 0.69314718055994530942f - logf(x);
}
// d/dx log(sinh(x)/x) = coth x - 1/x
__device__ __forceinline__ float dlog_sinhc(float x) {
    if (x < 0.5f) {
        const float x2 = x * x;
        // x/3 - x^3/45 + 2x^5/945 - x^7/4725
        return x * (1.0f / 3.0f + x2 * (-1.0f / 45.0f + x2 * (2.0f / 945.0f + x2 * (-1.0f / 4725.0f))));
    }
    const float e = expf(-2.0f * x);
    return (1.0f + e) / (1.0f - e) - 1.0f / x;
}

// ---- project(x): where(||x|| > maxnorm, x/||x||*maxnorm, x) ----------------------------------------
// Returns scale s so that y = s*x; *norm_out gets clamp_min(||x||, 1e-15).
template <int G, int EPL>
__device__ __forceinline__ bool project_inplace(RowSlice<G, EPL>& y, const Ball& ball, float& norm_out) {
    const float n = fmaxf(sqrtf(sqnorm<G, EPL>(y)), kMinNorm);
    norm_out = n;
    const bool hit = n > ball.maxnorm;
    if (hit) {
#pragma unroll
        for (int i = 0; i < EPL; ++i) y.v[i] = y.v[i] / n * ball.maxnorm;
    }
    return hit;
}

// backward of project given the PRE-projection row ypre (norm n, hit flag): g <- J^T g
template <int G, int EPL>
__device__ __forceinline__ void project_bwd(RowSlice<G, EPL>& g, const RowSlice<G, EPL>& ypre, float n, bool hit,
                                            const Ball& ball) {
    const float gy = dot<G, EPL>(g, ypre);  // all lanes take part in the shuffle
    if (hit) {
        const float s = ball.maxnorm / n;
        const float r = gy / (n * n);
#pragma unroll
        for (int i = 0; i < EPL; ++i) g.v[i] = s * (g.v[i] - r * ypre.v[i]);
    }
}

// ---- mobius_add (raw, no projection) ---------------------------------------------------------------
struct MAddCtx {
    float A, B, den, x2, y2, xy;
    bool den_clamped;
};

template <int G, int EPL>
__device__ __forceinline__ MAddCtx mobius_add_raw(const RowSlice<G, EPL>& x, const RowSlice<G, EPL>& y,
                                                  RowSlice<G, EPL>& out, const Ball& ball) {
    MAddCtx m;
    const float c = ball.c;
    m.x2 = sqnorm<G, EPL>(x);
    m.y2 = sqnorm<G, EPL>(y);
    m.xy = dot<G, EPL>(x, y);
    m.A = 1.0f + 2.0f * c * m.xy + c * m.y2;
    m.B = 1.0f - c * m.x2;
    const float den = 1.0f + 2.0f * c * m.xy + c * c * m.x2 * m.y2;
    m.den_clamped = den < kMinNorm;
    m.den = fmaxf(den, kMinNorm);
#pragma unroll
    for (int i = 0; i < EPL; ++i) out.v[i] = (m.A * x.v[i] + m.B * y.v[i]) / m.den;
    return m;
}

// g = dL/d(out) ; produces gx, gy
template <int G, int EPL>
__device__ __forceinline__ void mobius_add_raw_bwd(const RowSlice<G, EPL>& x, const RowSlice<G, EPL>& y,
                                                   const MAddCtx& m, const RowSlice<G, EPL>& g,
                                                   RowSlice<G, EPL>& gx, RowSlice<G, EPL>& gy, const Ball& ball) {
    const float c = ball.c;
    const float s = 1.0f / m.den;
    const float gdx = dot<G, EPL>(g, x);
    const float gdy = dot<G, EPL>(g, y);
    const float dA = s * gdx;
    const float dB = s * gdy;
    const float dnum = m.A * gdx + m.B * gdy;  // g . (A x + B y)
    const float dden = m.den_clamped ? 0.0f : -s * s * dnum;
    const float dxy = 2.0f * c * dA + 2.0f * c * dden;
    const float dx2 = -c * dB + c * c * m.y2 * dden;
    const float dy2 = c * dA + c * c * m.x2 * dden;
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
        gx.v[i] = s * m.A * g.v[i] + dxy * y.v[i] + 2.0f * dx2 * x.v[i];
        gy.v[i] = s * m.B * g.v[i] + dxy * x.v[i] + 2.0f * dy2 * y.v[i];
    }
}

// ---- launch geometry for row kernels -------------------------------------------------------------------
constexpr int kRowThreads = 256;

inline int row_grid(int64_t rows, int G) {
    const int64_t rows_per_block = kRowThreads / G;
    int64_t blocks = (rows + rows_per_block - 1) / rows_per_block;
    const int64_t cap = (int64_t)kNumSMs * 16;  // 16 resident CTAs of 256 threads is more than an SM holds; grid-stride
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

inline int check_launch() {
    return cudaPeekAtLastError() == cudaSuccess ? HVAE_OK : HVAE_ELAUNCH;
}

// Row-kernel dispatch over (G, EPL) by the row length D.  KERN is a template<int G,int EPL> __global__.
#define HVAE_ROW_LAUNCH(KERN, G_, EPL_, rows, stream, ...)                                        \
    KERN<G_, EPL_><<<hvae::row_grid((rows), G_), hvae::kRowThreads, 0, (stream)>>>(__VA_ARGS__)

#define HVAE_ROW_DISPATCH(KERN, D, rows, stream, ...)                                             \
    do {                                                                                          \
        if ((D) <= 2)        HVAE_ROW_LAUNCH(KERN, 1, 2, rows, stream, __VA_ARGS__);              \
        else if ((D) <= 4)   HVAE_ROW_LAUNCH(KERN, 1, 4, rows, stream, __VA_ARGS__);              \
        else if ((D) <= 8)   HVAE_ROW_LAUNCH(KERN, 2, 4, rows, stream, __VA_ARGS__);              \
        else if ((D) <= 16)  HVAE_ROW_LAUNCH(KERN, 4, 4, rows, stream, __VA_ARGS__);              \
        else if ((D) <= 32)  HVAE_ROW_LAUNCH(KERN, 8, 4, rows, stream, __VA_ARGS__);              \
        else if ((D) <= 64)  HVAE_ROW_LAUNCH(KERN, 16, 4, rows, stream, __VA_ARGS__);             \
        else if ((D) <= 128) HVAE_ROW_LAUNCH(KERN, 32, 4, rows, stream, __VA_ARGS__);             \
        else if ((D) <= 256) HVAE_ROW_LAUNCH(KERN, 32, 8, rows, stream, __VA_ARGS__);             \
        else if ((D) <= 512) HVAE_ROW_LAUNCH(KERN, 32, 16, rows, stream, __VA_ARGS__);            \
        else                 HVAE_ROW_LAUNCH(KERN, 32, 32, rows, stream, __VA_ARGS__);            \
    } while (0)

constexpr int kMaxRowDim = 1024;

// rows owned per warp-iteration and the lane's place in its group
#define HVAE_ROW_PROLOGUE(G)                                                  \
    const int lane = threadIdx.x & 31;                                        \
    const int lg = lane & ((G)-1);                                            \
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; \
    const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;       \
    constexpr int RPW = 32 / (G);                                             \
    const int sub = lane / (G);

}  // namespace hvae
