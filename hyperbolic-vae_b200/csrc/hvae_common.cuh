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

// ---- a row slice held in registers ----------------------------------------------------------------------
// Lane lg of the row's G-lane group holds EPL elements.  When D % 4 == 0 (and EPL % 4 == 0) a lane owns
// contiguous float4 chunks (index 4*(lg + G*j) + e): 128-bit coalesced loads/stores.  Otherwise elements are
// interleaved (index lg + i*G): 32-bit accesses, consecutive lanes on consecutive addresses.  D == 2 with one
// lane per row uses one 64-bit access.
template <int G, int EPL>
struct RowSlice {
    float v[EPL];

    __device__ __forceinline__ static bool vec4(int D) { return (EPL % 4 == 0) && ((D & 3) == 0); }
    __device__ __forceinline__ static int index(int lg, int i, int D) {
        return vec4(D) ? (((lg + G * (i >> 2)) << 2) + (i & 3)) : (lg + i * G);
    }

    __device__ __forceinline__ void load(const float* __restrict__ base, int64_t row, int D, int lg, bool valid) {
        const float* p = base + row * (int64_t)D;
        if (EPL % 4 == 0 && vec4(D)) {
#pragma unroll
            for (int j = 0; j < EPL / 4; ++j) {
                const int idx = (lg + G * j) << 2;
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid && idx < D) t = __ldg(reinterpret_cast<const float4*>(p + idx));
                v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
            }
        } else if (G == 1 && EPL == 2 && D == 2) {
            float2 t = make_float2(0.f, 0.f);
            if (valid) t = __ldg(reinterpret_cast<const float2*>(p));
            v[0] = t.x; v[1] = t.y;
        } else {
#pragma unroll
            for (int i = 0; i < EPL; ++i) {
                const int idx = lg + i * G;
                v[i] = (valid && idx < D) ? __ldg(p + idx) : 0.0f;
            }
        }
    }
    __device__ __forceinline__ void store(float* __restrict__ base, int64_t row, int D, int lg, bool valid) const {
        float* p = base + row * (int64_t)D;
        if (EPL % 4 == 0 && vec4(D)) {
#pragma unroll
            for (int j = 0; j < EPL / 4; ++j) {
                const int idx = (lg + G * j) << 2;
                if (valid && idx < D)
                    *reinterpret_cast<float4*>(p + idx) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
        } else if (G == 1 && EPL == 2 && D == 2) {
            if (valid) *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
        } else {
#pragma unroll
            for (int i = 0; i < EPL; ++i) {
                const int idx = lg + i * G;
                if (valid && idx < D) p[idx] = v[i];
            }
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
// Branch-free fast-intrinsic evaluations, each accurate to ~1e-6 relative (parity budget 1e-5):
//   tanh   : odd Taylor polynomial to x^13 below 0.4 (3.9e-9), (E-1)/(E+1) with E = ex2.approx above
//   artanh : odd series to x^11 below 0.2 (3.2e-10), 0.5*lg2.approx((1+x)/(1-x)) above (1-x exact for x >= 0.5)
__device__ __forceinline__ float rcpf(float v) { return __fdividef(1.0f, v); }

// t = tanh(clamp(x, +-15)); s2 = sech^2 of the clamped argument (0 outside the clamp: clamp has zero gradient there)
__device__ __forceinline__ void tanh_sech2(float x, float& t, float& s2) {
    const float ax = fminf(fabsf(x), kTanhClamp);
    const float x2 = ax * ax;
    const float poly = ax * fmaf(x2, fmaf(x2, fmaf(x2, fmaf(x2, fmaf(x2, fmaf(x2, 21844.0f / 6081075.0f, -1382.0f / 155925.0f),
                       62.0f / 2835.0f), -17.0f / 315.0f), 2.0f / 15.0f), -1.0f / 3.0f), 1.0f);
    const float E = __expf(2.0f * ax);
    const float r = rcpf(E + 1.0f);
    const float big = (E - 1.0f) * r;
    const float tt = ax < 0.4f ? poly : big;
    t = copysignf(tt, x);
    // sech^2 = 4E/(E+1)^2 has no cancellation when tanh saturates (1 - t^2 does)
    s2 = (fabsf(x) <= kTanhClamp) ? 4.0f * E * r * r : 0.0f;
}
__device__ __forceinline__ float tanh_c(float x) {
    float t, s2;
    tanh_sech2(x, t, s2);
    return t;
}
__device__ __forceinline__ float tanh_mask(float x) { return (x >= -kTanhClamp && x <= kTanhClamp) ? 1.0f : 0.0f; }

__device__ __forceinline__ float artanh_c(float x) {
    const float ax = fminf(fabsf(x), kArtanhClamp);
    const float x2 = ax * ax;
    const float poly = ax * fmaf(x2, fmaf(x2, fmaf(x2, fmaf(x2, fmaf(x2, 1.0f / 11.0f, 1.0f / 9.0f), 1.0f / 7.0f), 1.0f / 5.0f),
                       1.0f / 3.0f), 1.0f);
    const float big = 0.5f * __logf(__fdividef(1.0f + ax, 1.0f - ax));
    return copysignf(ax < 0.2f ? poly : big, x);
}
// d/dx artanh(clamp(x)) = 1/(1-xc^2) inside the clamp, 0 outside
__device__ __forceinline__ float artanh_grad(float x) {
    if (!(x >= -kArtanhClamp && x <= kArtanhClamp)) return 0.0f;
    return rcpf((1.0f - x) * (1.0f + x));
}

// log(sinh(x)/x) for x >= 0, accurate at small x (the reference's log sinh - log x cancels there)
__device__ __forceinline__ float log_sinhc(float x) {
    const float x2 = x * x;
    // x^2/6 - x^4/180 + x^6/2835 - x^8/37800 + x^10/467775   (|x| < 0.75: next term 3e-9)
    const float small = x2 * fmaf(x2, fmaf(x2, fmaf(x2, fmaf(x2, 1.0f / 467775.0f, -1.0f / 37800.0f), 1.0f / 2835.0f), -1.0f / 180.0f), 1.0f / 6.0f);
    // log sinh x - log x = x + log((1 - e^{-2x})/(2x))
    const float big = x + __logf(__fdividef(1.0f - __expf(-2.0f * x), 2.0f * x));
    return x < 0.75f ? small : big;
}
// d/dx log(sinh(x)/x) = coth x - 1/x
__device__ __forceinline__ float dlog_sinhc(float x) {
    const float x2 = x * x;
    // x/3 - x^3/45 + 2x^5/945 - x^7/4725 + 2x^9/93555
    const float small = x * fmaf(x2, fmaf(x2, fmaf(x2, fmaf(x2, 2.0f / 93555.0f, -1.0f / 4725.0f), 2.0f / 945.0f), -1.0f / 45.0f), 1.0f / 3.0f);
    const float e = __expf(-2.0f * x);
    const float big = __fdividef(1.0f + e, 1.0f - e) - rcpf(x);
    return x < 0.75f ? small : big;
}

// sqrt via rsqrt (1 ulp-ish), exact 0 at 0
__device__ __forceinline__ float sqrt_fast(float v) { return v > 0.0f ? v * rsqrtf(v) : 0.0f; }

// ---- project(x): where(||x|| > maxnorm, x/||x||*maxnorm, x) ----------------------------------------
// Returns scale s so that y = s*x; *norm_out gets clamp_min(||x||, 1e-15).
template <int G, int EPL>
__device__ __forceinline__ bool project_inplace(RowSlice<G, EPL>& y, const Ball& ball, float& norm_out) {
    const float n = fmaxf(sqrt_fast(sqnorm<G, EPL>(y)), kMinNorm);
    norm_out = n;
    const bool hit = n > ball.maxnorm;
    if (hit) {
        const float s = ball.maxnorm * rcpf(n);
#pragma unroll
        for (int i = 0; i < EPL; ++i) y.v[i] *= s;
    }
    return hit;
}

// backward of project given the PRE-projection row ypre (norm n, hit flag): g <- J^T g
template <int G, int EPL>
__device__ __forceinline__ void project_bwd(RowSlice<G, EPL>& g, const RowSlice<G, EPL>& ypre, float n, bool hit,
                                            const Ball& ball) {
    const float gy = dot<G, EPL>(g, ypre);  // all lanes take part in the shuffle
    if (hit) {
        const float rn = rcpf(n);
        const float s = ball.maxnorm * rn;
        const float r = gy * rn * rn;
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
    const float rden = rcpf(m.den);
    const float ar = m.A * rden, br = m.B * rden;
#pragma unroll
    for (int i = 0; i < EPL; ++i) out.v[i] = fmaf(ar, x.v[i], br * y.v[i]);
    return m;
}

// g = dL/d(out) ; produces gx, gy
template <int G, int EPL>
__device__ __forceinline__ void mobius_add_raw_bwd(const RowSlice<G, EPL>& x, const RowSlice<G, EPL>& y,
                                                   const MAddCtx& m, const RowSlice<G, EPL>& g,
                                                   RowSlice<G, EPL>& gx, RowSlice<G, EPL>& gy, const Ball& ball) {
    const float c = ball.c;
    const float s = rcpf(m.den);
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
        else if ((D) <= 8)   HVAE_ROW_LAUNCH(KERN, 1, 8, rows, stream, __VA_ARGS__);              \
        else if ((D) <= 16)  HVAE_ROW_LAUNCH(KERN, 2, 8, rows, stream, __VA_ARGS__);              \
        else if ((D) <= 32)  HVAE_ROW_LAUNCH(KERN, 4, 8, rows, stream, __VA_ARGS__);              \
        else if ((D) <= 64)  HVAE_ROW_LAUNCH(KERN, 8, 8, rows, stream, __VA_ARGS__);              \
        else if ((D) <= 128) HVAE_ROW_LAUNCH(KERN, 16, 8, rows, stream, __VA_ARGS__);             \
        else if ((D) <= 256) HVAE_ROW_LAUNCH(KERN, 32, 8, rows, stream, __VA_ARGS__);             \
        else if ((D) <= 512) HVAE_ROW_LAUNCH(KERN, 32, 16, rows, stream, __VA_ARGS__);            \
        else                 HVAE_ROW_LAUNCH(KERN, 32, 32, rows, stream, __VA_ARGS__);            \
    } while (0)

constexpr int kMaxRowDim = 1024;

// Rows in flight per lane group.  At D <= 4 a row is 8-16 bytes: with one row per thread the kernels are bound by
// memory-level parallelism (bytes in flight per SM), not bandwidth; loading U rows before the math fixes that.
#define HVAE_ROW_UNROLL(EPL) ((EPL) <= 2 ? 4 : ((EPL) <= 4 ? 2 : 1))

// rows owned per warp-iteration and the lane's place in its group
#define HVAE_ROW_PROLOGUE(G)                                                  \
    const int lane = threadIdx.x & 31;                                        \
    const int lg = lane & ((G)-1);                                            \
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; \
    const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;       \
    constexpr int RPW = 32 / (G);                                             \
    const int sub = lane / (G);


// ---------------------------------------------------------------------------------------------------
// Deterministic slab reduction  out[i] = sum_k w[k][i]  for up to four (w, out, n, slabs) segments in ONE launch
// (blockIdx.y = segment): the partial-sum buffers of the skinny gradient GEMMs (gyroplane gx / gp / ga / gbias, Mobius
// gM).  A block is 32 float4 columns x 8 slab lanes: lane group s walks slabs s, s + 8, ... with up to 8 independent
// 16-byte loads in flight per thread (coalesced 512-byte rows), then the 8 partials are summed in a fixed order.
// (A thread per element walking all slabs serially is a chain of up to 64 dependent L2 round trips: 7-14 us for 3 MB.)
// ---------------------------------------------------------------------------------------------------
struct SlabSeg {
    const float* w;
    float* out;
    int64_t n;
    int slabs;
    int vec;   // n % 4 == 0 and both pointers 16-byte aligned: float4 columns
};
struct SlabReduceArgs {
    SlabSeg seg[4];
};

#pragma nv_diag_suppress 177   // a translation unit that includes this header without reducing anything: no "never referenced" noise
static __global__ void __launch_bounds__(256) k_reduce_slabs(SlabReduceArgs a) {
    __shared__ float4 part[8][32];
    const SlabSeg sg = a.seg[blockIdx.y];
    const int e = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int64_t units = sg.vec ? (sg.n >> 2) : sg.n;
    for (int64_t u0 = (int64_t)blockIdx.x * 32; u0 < units; u0 += (int64_t)gridDim.x * 32) {
        const int64_t u = u0 + e;
        float4 s = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (u < units) {
            if (sg.vec) {
                const float4* w4 = reinterpret_cast<const float4*>(sg.w);
#pragma unroll 8
                for (int k = sl; k < sg.slabs; k += 8) {
                    const float4 v = __ldg(w4 + (int64_t)k * units + u);
                    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
                }
            } else {
#pragma unroll 8
                for (int k = sl; k < sg.slabs; k += 8) s.x += __ldg(sg.w + (int64_t)k * sg.n + u);
            }
        }
        part[sl][e] = s;
        __syncthreads();
        if (sl == 0 && u < units) {
            float4 t = part[0][e];
#pragma unroll
            for (int q = 1; q < 8; ++q) {
                const float4 v = part[q][e];
                t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
            }
            if (sg.vec) reinterpret_cast<float4*>(sg.out)[u] = t;
            else sg.out[u] = t.x;
        }
        __syncthreads();
    }
}

#pragma nv_diag_default 177

struct SlabReducer {
    SlabReduceArgs args;
    int nseg = 0;
    int64_t max_units = 0;
    void add(const float* w, float* out, int64_t n, int slabs) {
        if (n <= 0 || nseg >= 4) return;
        SlabSeg& g = args.seg[nseg++];
        g.w = w; g.out = out; g.n = n; g.slabs = slabs;
        g.vec = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(out)) % 16 == 0);
        const int64_t units = g.vec ? n / 4 : n;
        if (units > max_units) max_units = units;
    }
    void launch(cudaStream_t s) const {
        if (nseg == 0) return;
        const int64_t bx = (max_units + 31) / 32;
        dim3 grid((unsigned)(bx < 4 * kNumSMs ? bx : 4 * kNumSMs), (unsigned)nseg);
        k_reduce_slabs<<<grid, 256, 0, s>>>(args);
    }
};

}  // namespace hvae
