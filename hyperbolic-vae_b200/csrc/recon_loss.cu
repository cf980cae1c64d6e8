// Reconstruction-loss head of the decoders: Bernoulli negative log-likelihood with logits, summed over the feature
// axis of every (sample, row):   nll[s,b] = sum_n  max(l,0) - l x + log1p(exp(-|l|)),   l = logits[s,b,n], x = x[b,n]
// reference: torch.distributions.Bernoulli(logits).log_prob(x).sum(-1) in the pvae objective
// (hyperbolic_vae/training/old_pvae_train.py:53-58) and F.binary_cross_entropy_with_logits in
// hyperbolic_vae/models/vae_hyperbolic_gyroplane_decoder.py loss_recon (SURVEY 8f "recon-loss heads").
// torch runs this as ~6 elementwise + reduce kernels forward and ~6 backward over (B, 784) tensors; here one warp
// walks one row once per direction.  HBM-bound: forward 8N + 4 bytes per row, backward 12N + 4.
#include "hvae_common.cuh"

namespace hvae {

__device__ __forceinline__ float bce_term(float l, float x) {
    return fmaxf(l, 0.0f) - l * x + log1pf(expf(-fabsf(l)));
}
__device__ __forceinline__ float sigmoid_acc(float l) {
    const float e = expf(-fabsf(l));
    const float s = 1.0f / (1.0f + e);     // sigmoid(|l|)
    return l >= 0.0f ? s : e * s;           // sigmoid(-|l|) = e / (1 + e)
}

__global__ void __launch_bounds__(256)
k_bce_logits_rows_fwd(const float* __restrict__ logits, const float* __restrict__ x, float* __restrict__ nll, int64_t S,
                      int64_t B, int64_t N) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const bool v4 = (N & 3) == 0;
    for (int64_t row = warp; row < S * B; row += nw) {
        const float* lr = logits + row * N;
        const float* xr = x + (row % B) * N;
        float acc = 0.0f;
        if (v4) {
            for (int64_t i = lane * 4; i < N; i += 128) {
                const float4 l = __ldg(reinterpret_cast<const float4*>(lr + i));
                const float4 t = __ldg(reinterpret_cast<const float4*>(xr + i));
                acc += (bce_term(l.x, t.x) + bce_term(l.y, t.y)) + (bce_term(l.z, t.z) + bce_term(l.w, t.w));
            }
        } else {
            for (int64_t i = lane; i < N; i += 32) acc += bce_term(__ldg(lr + i), __ldg(xr + i));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) nll[row] = acc;
    }
}

// glogits[s,b,n] = gnll[s,b] * (sigmoid(l) - x)
__global__ void __launch_bounds__(256)
k_bce_logits_rows_bwd(const float* __restrict__ logits, const float* __restrict__ x, const float* __restrict__ gnll,
                      float* __restrict__ glogits, int64_t S, int64_t B, int64_t N) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const bool v4 = (N & 3) == 0;
    for (int64_t row = warp; row < S * B; row += nw) {
        const float* lr = logits + row * N;
        const float* xr = x + (row % B) * N;
        float* gr = glogits + row * N;
        const float g = __ldg(gnll + row);
        if (v4) {
            for (int64_t i = lane * 4; i < N; i += 128) {
                const float4 l = __ldg(reinterpret_cast<const float4*>(lr + i));
                const float4 t = __ldg(reinterpret_cast<const float4*>(xr + i));
                float4 o;
                o.x = g * (sigmoid_acc(l.x) - t.x);
                o.y = g * (sigmoid_acc(l.y) - t.y);
                o.z = g * (sigmoid_acc(l.z) - t.z);
                o.w = g * (sigmoid_acc(l.w) - t.w);
                *reinterpret_cast<float4*>(gr + i) = o;
            }
        } else {
            for (int64_t i = lane; i < N; i += 32) gr[i] = g * (sigmoid_acc(__ldg(lr + i)) - __ldg(xr + i));
        }
    }
}

// ---- generic reconstruction heads: per-(sample, row) sums over the feature axis ---------------------
//   HVAE_RECON_MSE          sum_n (in - x)^2                      F.mse_loss(x_hat, x, "sum")       models/vae_hyperbolic.py:219
//   HVAE_RECON_SIGMOID_MSE  sum_n (sigmoid(in) - x)^2             decoder's final nn.Sigmoid fused  ...rnaseq.py:57,107
//   HVAE_RECON_RB_LOGITS    -RelaxedBernoulli(T, logits=in).log_prob(x)   models/vae_hyperbolic.py:224-225
//   HVAE_RECON_RB_PROBS     -RelaxedBernoulli(T, probs=in).log_prob(x)    ...gyroplane_decoder.py:121-122
//   HVAE_RECON_RB_SIGMOID   the same with probs = sigmoid(in) (the decoder's final nn.Sigmoid fused)
// RelaxedBernoulli = torch.distributions semantics (LogitRelaxedBernoulli + SigmoidTransform): probs clamped to
// [eps, 1-eps] (eps = 2^-23), value to [tiny, 1-eps];  with y = logit(v), d = logits - T y:
//   -log_prob = -log T - d + 2 softplus(d) - softplus(-y) - softplus(y),      d(-log_prob)/dlogits = 2 sigmoid(d) - 1.
__device__ __forceinline__ float softplus_acc(float v) { return fmaxf(v, 0.0f) + log1pf(expf(-fabsf(v))); }

template <int KIND>
__device__ __forceinline__ float recon_term(float in, float x, float T, float logT) {
    if (KIND == HVAE_RECON_MSE) { const float d = in - x; return d * d; }
    if (KIND == HVAE_RECON_SIGMOID_MSE) { const float d = sigmoid_acc(in) - x; return d * d; }
    constexpr float eps = 1.1920928955078125e-07f, tiny = 1.17549435e-38f;
    float lg = in;
    if (KIND != HVAE_RECON_RB_LOGITS) {
        const float p = (KIND == HVAE_RECON_RB_SIGMOID) ? sigmoid_acc(in) : in;
        const float ps = fminf(fmaxf(p, eps), 1.0f - eps);
        lg = logf(ps) - log1pf(-ps);
    }
    const float v = fminf(fmaxf(x, tiny), 1.0f - eps);
    const float y = logf(v) - log1pf(-v);
    const float d = lg - y * T;
    return -logT - d + 2.0f * softplus_acc(d) - softplus_acc(-y) - softplus_acc(y);
}
// d term / d in
template <int KIND>
__device__ __forceinline__ float recon_grad(float in, float x, float T) {
    if (KIND == HVAE_RECON_MSE) return 2.0f * (in - x);
    if (KIND == HVAE_RECON_SIGMOID_MSE) { const float s = sigmoid_acc(in); return 2.0f * (s - x) * s * (1.0f - s); }
    constexpr float eps = 1.1920928955078125e-07f, tiny = 1.17549435e-38f;
    float lg = in, dlg = 1.0f;
    if (KIND != HVAE_RECON_RB_LOGITS) {
        const float p = (KIND == HVAE_RECON_RB_SIGMOID) ? sigmoid_acc(in) : in;
        const bool inside = p >= eps && p <= 1.0f - eps;       // clamp: zero gradient outside
        const float ps = fminf(fmaxf(p, eps), 1.0f - eps);
        lg = logf(ps) - log1pf(-ps);
        // d logit(ps)/d in: 1/(ps (1-ps)) [probs]   or   1 [sigmoid: logit(sigmoid(in)) = in]
        dlg = inside ? ((KIND == HVAE_RECON_RB_SIGMOID) ? 1.0f : 1.0f / (ps * (1.0f - ps))) : 0.0f;
    }
    const float v = fminf(fmaxf(x, tiny), 1.0f - eps);
    const float y = logf(v) - log1pf(-v);
    return (2.0f * sigmoid_acc(lg - y * T) - 1.0f) * dlg;
}

template <int KIND>
__global__ void __launch_bounds__(256)
k_recon_rows_fwd(const float* __restrict__ in, const float* __restrict__ x, float* __restrict__ out, int64_t S, int64_t B,
                 int64_t N, float T, float logT) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const bool v4 = (N & 3) == 0;
    for (int64_t row = warp; row < S * B; row += nw) {
        const float* lr = in + row * N;
        const float* xr = x + (row % B) * N;
        float acc = 0.0f;
        if (v4) {
            for (int64_t i = lane * 4; i < N; i += 128) {
                const float4 l = __ldg(reinterpret_cast<const float4*>(lr + i));
                const float4 t = __ldg(reinterpret_cast<const float4*>(xr + i));
                acc += (recon_term<KIND>(l.x, t.x, T, logT) + recon_term<KIND>(l.y, t.y, T, logT)) +
                       (recon_term<KIND>(l.z, t.z, T, logT) + recon_term<KIND>(l.w, t.w, T, logT));
            }
        } else {
            for (int64_t i = lane; i < N; i += 32) acc += recon_term<KIND>(__ldg(lr + i), __ldg(xr + i), T, logT);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) out[row] = acc;
    }
}

template <int KIND>
__global__ void __launch_bounds__(256)
k_recon_rows_bwd(const float* __restrict__ in, const float* __restrict__ x, const float* __restrict__ gout,
                 float* __restrict__ gin, int64_t S, int64_t B, int64_t N, float T) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const bool v4 = (N & 3) == 0;
    for (int64_t row = warp; row < S * B; row += nw) {
        const float* lr = in + row * N;
        const float* xr = x + (row % B) * N;
        float* gr = gin + row * N;
        const float g = __ldg(gout + row);
        if (v4) {
            for (int64_t i = lane * 4; i < N; i += 128) {
                const float4 l = __ldg(reinterpret_cast<const float4*>(lr + i));
                const float4 t = __ldg(reinterpret_cast<const float4*>(xr + i));
                float4 o;
                o.x = g * recon_grad<KIND>(l.x, t.x, T);
                o.y = g * recon_grad<KIND>(l.y, t.y, T);
                o.z = g * recon_grad<KIND>(l.z, t.z, T);
                o.w = g * recon_grad<KIND>(l.w, t.w, T);
                *reinterpret_cast<float4*>(gr + i) = o;
            }
        } else {
            for (int64_t i = lane; i < N; i += 32) gr[i] = g * recon_grad<KIND>(__ldg(lr + i), __ldg(xr + i), T);
        }
    }
}

// Column sums of a row-major (R, C) matrix (bias gradient of a dense layer: gb = sum_rows gy), deterministic two-pass:
// pass 1: block (col tile of 32, row chunk) -> partial[chunk][C]; pass 2: sum the chunks.  A warp reads 128 contiguous
// bytes of a row.  torch's generic reduction spends ~12 us on (4096, 784); this is bandwidth-bound (~3 us from L2).
constexpr int kColsumChunks = 32;
__global__ void __launch_bounds__(256)
k_colsum_partial(const float* __restrict__ x, float* __restrict__ part, int64_t R, int64_t C) {
    __shared__ float sm[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t c = (int64_t)blockIdx.x * 32 + tx;
    const int64_t per = (R + kColsumChunks - 1) / kColsumChunks;
    const int64_t r0 = (int64_t)blockIdx.y * per, r1 = (r0 + per < R) ? r0 + per : R;
    float a0 = 0.0f, a1 = 0.0f;
    if (c < C) {
        int64_t r = r0 + ty;
        for (; r + 8 < r1; r += 16) {
            a0 += __ldg(x + r * C + c);
            a1 += __ldg(x + (r + 8) * C + c);
        }
        if (r < r1) a0 += __ldg(x + r * C + c);
    }
    sm[ty][tx] = a0 + a1;
    __syncthreads();
    if (ty == 0 && c < C) {
        float a = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) a += sm[i][tx];
        part[(int64_t)blockIdx.y * C + c] = a;
    }
}
__global__ void k_colsum_final(const float* __restrict__ part, float* __restrict__ out, int64_t C) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float a = 0.0f;
    for (int i = 0; i < kColsumChunks; ++i) a += part[(int64_t)i * C + c];
    out[c] = a;
}

}  // namespace hvae

using namespace hvae;

extern "C" size_t hvae_colsum_workspace_bytes(int64_t C) { return C > 0 ? (size_t)kColsumChunks * C * 4 : 0; }

extern "C" int hvae_colsum_f32(const float* x, float* out, int64_t R, int64_t C, void* workspace, size_t workspace_bytes,
                               void* stream) {
    if (R <= 0 || C <= 0) return HVAE_ESHAPE;
    if (!x || !out || !workspace || workspace_bytes < (size_t)kColsumChunks * C * 4) return HVAE_EARG;
    cudaStream_t s = (cudaStream_t)stream;
    dim3 grid((unsigned)((C + 31) / 32), kColsumChunks);
    k_colsum_partial<<<grid, 256, 0, s>>>(x, (float*)workspace, R, C);
    k_colsum_final<<<(unsigned)((C + 255) / 256), 256, 0, s>>>((const float*)workspace, out, C);
    return check_launch();
}

static unsigned bce_grid(int64_t rows) {
    const int64_t want = (rows + 7) / 8;  // 8 warps per CTA
    const int64_t cap = (int64_t)kNumSMs * 16;
    return (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
}

extern "C" int hvae_bce_logits_rows_fwd_f32(const float* logits, const float* x, float* nll, int64_t S, int64_t B, int64_t N,
                                            void* stream) {
    if (S <= 0 || B <= 0 || N <= 0) return HVAE_ESHAPE;
    if (!logits || !x || !nll) return HVAE_EARG;
    k_bce_logits_rows_fwd<<<bce_grid(S * B), 256, 0, (cudaStream_t)stream>>>(logits, x, nll, S, B, N);
    return check_launch();
}

extern "C" int hvae_bce_logits_rows_bwd_f32(const float* logits, const float* x, const float* gnll, float* glogits, int64_t S,
                                            int64_t B, int64_t N, void* stream) {
    if (S <= 0 || B <= 0 || N <= 0) return HVAE_ESHAPE;
    if (!logits || !x || !gnll || !glogits) return HVAE_EARG;
    k_bce_logits_rows_bwd<<<bce_grid(S * B), 256, 0, (cudaStream_t)stream>>>(logits, x, gnll, glogits, S, B, N);
    return check_launch();
}

// in (S,B,N), x (B,N) broadcast over S -> out (S,B); kind: HVAE_RECON_*; temperature: RelaxedBernoulli kinds only
extern "C" int hvae_recon_rows_fwd_f32(const float* in, const float* x, float* out, int64_t S, int64_t B, int64_t N, int kind,
                                       float temperature, void* stream) {
    if (S <= 0 || B <= 0 || N <= 0) return HVAE_ESHAPE;
    if (!in || !x || !out) return HVAE_EARG;
    cudaStream_t s = (cudaStream_t)stream;
    const float T = temperature, lT = (kind >= HVAE_RECON_RB_LOGITS) ? logf(temperature) : 0.0f;
    const unsigned grid = bce_grid(S * B);
    switch (kind) {
        case HVAE_RECON_MSE: k_recon_rows_fwd<HVAE_RECON_MSE><<<grid, 256, 0, s>>>(in, x, out, S, B, N, T, lT); break;
        case HVAE_RECON_SIGMOID_MSE: k_recon_rows_fwd<HVAE_RECON_SIGMOID_MSE><<<grid, 256, 0, s>>>(in, x, out, S, B, N, T, lT); break;
        case HVAE_RECON_RB_LOGITS: k_recon_rows_fwd<HVAE_RECON_RB_LOGITS><<<grid, 256, 0, s>>>(in, x, out, S, B, N, T, lT); break;
        case HVAE_RECON_RB_PROBS: k_recon_rows_fwd<HVAE_RECON_RB_PROBS><<<grid, 256, 0, s>>>(in, x, out, S, B, N, T, lT); break;
        case HVAE_RECON_RB_SIGMOID: k_recon_rows_fwd<HVAE_RECON_RB_SIGMOID><<<grid, 256, 0, s>>>(in, x, out, S, B, N, T, lT); break;
        default: return HVAE_EARG;
    }
    return check_launch();
}

extern "C" int hvae_recon_rows_bwd_f32(const float* in, const float* x, const float* gout, float* gin, int64_t S, int64_t B,
                                       int64_t N, int kind, float temperature, void* stream) {
    if (S <= 0 || B <= 0 || N <= 0) return HVAE_ESHAPE;
    if (!in || !x || !gout || !gin) return HVAE_EARG;
    cudaStream_t s = (cudaStream_t)stream;
    const float T = temperature;
    const unsigned grid = bce_grid(S * B);
    switch (kind) {
        case HVAE_RECON_MSE: k_recon_rows_bwd<HVAE_RECON_MSE><<<grid, 256, 0, s>>>(in, x, gout, gin, S, B, N, T); break;
        case HVAE_RECON_SIGMOID_MSE: k_recon_rows_bwd<HVAE_RECON_SIGMOID_MSE><<<grid, 256, 0, s>>>(in, x, gout, gin, S, B, N, T); break;
        case HVAE_RECON_RB_LOGITS: k_recon_rows_bwd<HVAE_RECON_RB_LOGITS><<<grid, 256, 0, s>>>(in, x, gout, gin, S, B, N, T); break;
        case HVAE_RECON_RB_PROBS: k_recon_rows_bwd<HVAE_RECON_RB_PROBS><<<grid, 256, 0, s>>>(in, x, gout, gin, S, B, N, T); break;
        case HVAE_RECON_RB_SIGMOID: k_recon_rows_bwd<HVAE_RECON_RB_SIGMOID><<<grid, 256, 0, s>>>(in, x, gout, gin, S, B, N, T); break;
        default: return HVAE_EARG;
    }
    return check_launch();
}

// ---- loss tail of the pvae objective (training/old_pvae_train.py:53-58 with K samples): from the per-(sample, row)
// negative log-likelihood and KL terms to the three scalars the step reports, in ONE launch per direction
//   recon = sum_b mean_s nll[s,b],  kl = sum_b mean_s kld[s,b],  total = recon + beta kl
// (torch: neg, mean, sum, neg, mean, sum, mul, add forward and as many small kernels backward - ~24 launches on 16 KB)
namespace hvae {
__global__ void __launch_bounds__(1024)
k_pvae_loss_fwd(const float* __restrict__ nll, const float* __restrict__ kld, float* __restrict__ out, int64_t n, float inv_s,
                float beta) {
    __shared__ double sa[32], sb[32];
    double a = 0.0, b = 0.0;   // 4096+ terms of O(1e3): keep the batch sums exact to fp32 rounding
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) { a += (double)nll[i]; b += (double)kld[i]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = a; sb[threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double A = 0.0, B = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { A += sa[w]; B += sb[w]; }
        const float recon = (float)(A * inv_s), kl = (float)(B * inv_s);
        out[0] = recon + beta * kl;
        out[1] = recon;
        out[2] = kl;
    }
}
// gout: upstream gradients of (total, recon, kl) (NULL entries = 0 are passed as zeros by the caller)
__global__ void k_pvae_loss_bwd(const float* __restrict__ gout, float* __restrict__ gnll, float* __restrict__ gkld, int64_t n,
                                float inv_s, float beta) {
    const float ga = (gout[0] + gout[1]) * inv_s, gb = (beta * gout[0] + gout[2]) * inv_s;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        gnll[i] = ga;
        gkld[i] = gb;
    }
}
}  // namespace hvae

// nll, kld: (S, B) -> out[3] = {total, recon, kl}
extern "C" int hvae_pvae_loss_fwd_f32(const float* nll, const float* kld, float* out, int64_t S, int64_t B, float beta, void* stream) {
    if (S <= 0 || B <= 0) return HVAE_ESHAPE;
    if (!nll || !kld || !out) return HVAE_EARG;
    hvae::k_pvae_loss_fwd<<<1, 1024, 0, (cudaStream_t)stream>>>(nll, kld, out, S * B, 1.0f / (float)S, beta);
    return check_launch();
}
extern "C" int hvae_pvae_loss_bwd_f32(const float* gout, float* gnll, float* gkld, int64_t S, int64_t B, float beta, void* stream) {
    if (S <= 0 || B <= 0) return HVAE_ESHAPE;
    if (!gout || !gnll || !gkld) return HVAE_EARG;
    const int64_t n = S * B;
    hvae::k_pvae_loss_bwd<<<(unsigned)((n + 255) / 256 < 64 ? (n + 255) / 256 : 64), 256, 0, (cudaStream_t)stream>>>(gout, gnll, gkld, n,
                                                                                                               1.0f / (float)S, beta);
    return check_launch();
}

// ---- posterior scale head of the pvae encoder (scripts/_9_pvae_replicate.py: softplus(fc22(e)) + 1e-5) together with the
// clamp RiemannianNormal applies to it (distributions/old_pvae_riemannian_normal.py:30: scale.clamp(0.1, 7)):
//   sigma = clamp(softplus(h) + eps, lo, hi);  d sigma / d h = sigmoid(h) inside the clamp, 0 where it binds.
namespace hvae {
__global__ void k_sigma_head_fwd(const float* __restrict__ h, float* __restrict__ out, int64_t n, float eps, float lo, float hi) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = h[i];
    const float sp = (v > 20.0f) ? v : log1pf(expf(v));   // torch.nn.functional.softplus (threshold 20)
    out[i] = fminf(fmaxf(sp + eps, lo), hi);
}
__global__ void k_sigma_head_bwd(const float* __restrict__ h, const float* __restrict__ g, float* __restrict__ gh, int64_t n, float eps,
                                 float lo, float hi) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = h[i];
    const float sp = (v > 20.0f) ? v : log1pf(expf(v));
    const float s = sp + eps;
    const float d = (v > 20.0f) ? 1.0f : sigmoid_acc(v);
    gh[i] = (s >= lo && s <= hi) ? g[i] * d : 0.0f;      // clamp passes the gradient on [lo, hi] (torch semantics)
}
}  // namespace hvae

extern "C" int hvae_sigma_head_fwd_f32(const float* h, float* out, int64_t n, float eps, float lo, float hi, void* stream) {
    if (n <= 0) return HVAE_ESHAPE;
    if (!h || !out) return HVAE_EARG;
    hvae::k_sigma_head_fwd<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(h, out, n, eps, lo, hi);
    return check_launch();
}
extern "C" int hvae_sigma_head_bwd_f32(const float* h, const float* g, float* gh, int64_t n, float eps, float lo, float hi, void* stream) {
    if (n <= 0) return HVAE_ESHAPE;
    if (!h || !g || !gh) return HVAE_EARG;
    hvae::k_sigma_head_bwd<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(h, g, gh, n, eps, lo, hi);
    return check_launch();
}
