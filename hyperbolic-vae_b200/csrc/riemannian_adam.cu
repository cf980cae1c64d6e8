// f-1: the optimizer step right after the gradient exchange - geoopt.optim.RiemannianAdam as ONE
// multi-tensor kernel over every parameter of the model, reading the gradients where the data-parallel exchange left
// them (views of the flat bucket).
//
// reference call sites: hyperbolic_vae/models/vae_hyperbolic.py:235-243 (configure_optimizers),
// ...gyroplane_decoder.py:173, ...rnaseq.py:139, vae_one_b.py:270; arithmetic: geoopt RiemannianAdam (App. A.1 end):
//   Euclidean tensor (elementwise):  g += wd p;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
//                                    p -= lr (m / bc1) / (sqrt(v / bc2) + eps)
//   Poincare-ball rows x (ManifoldParameter: gyroplane `points`, over-parameterised `_bias`), lambda = 2 / (1 - c|x|^2):
//                                    rg = (g + wd x) / lambda^2                       (egrad2rgrad)
//                                    m = b1 m + (1-b1) rg;  v = b2 v + (1-b2) lambda^2 |rg|^2   (one value per row)
//                                    y = project(x - lr (m / bc1) / (sqrt(v / bc2) + eps))      (retraction)
//                                    m <- gyr[y, -x] m * lambda_x / lambda_y                    (parallel transport)
// torch runs this as ~12 launches per parameter (x 10-20 parameters); here the whole model is one launch.
// Hyper-parameters and the step count live in DEVICE memory so a captured CUDA graph sees the scheduler's new learning
// rate and the advancing bias corrections on every replay.
#include "hvae_common.cuh"

namespace hvae {

struct AdamTensor {
    float* p;
    const float* g;
    float* m;
    float* v;
    int64_t numel;
    int64_t cols;      // manifold tensors: row length D (numel = rows * D); Euclidean: unused
    float c;           // > 0: Poincare-ball rows with this curvature; 0: Euclidean
    int first_block;   // first block of the grid that works on this tensor
};
struct AdamHyper {
    float lr, b1, b2, eps, wd;
    float step;        // 1-based step count of THIS update (the host or a graph node advances it before the launch)
};

constexpr int kAdamThreads = 256;
constexpr int kAdamChunk = kAdamThreads * 8;   // Euclidean elements per block
constexpr int kAdamRowsPerBlock = kAdamThreads / 32;

__global__ void __launch_bounds__(kAdamThreads)
k_riemannian_adam(const AdamTensor* __restrict__ tensors, int n_tensors, const AdamHyper* __restrict__ hyper) {
    // which tensor does this block belong to (first_block is ascending; a handful of tensors: linear scan)
    int t = 0;
    while (t + 1 < n_tensors && (int)blockIdx.x >= tensors[t + 1].first_block) ++t;
    const AdamTensor T = tensors[t];
    const AdamHyper h = *hyper;
    const float bc1 = 1.0f - powf(h.b1, h.step), bc2 = 1.0f - powf(h.b2, h.step);
    const int blk = blockIdx.x - T.first_block;
    if (T.c == 0.0f) {
        const int64_t i0 = (int64_t)blk * kAdamChunk;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int64_t i = i0 + j * kAdamThreads + threadIdx.x;
            if (i >= T.numel) break;
            const float p = T.p[i];
            const float g = T.g[i] + h.wd * p;
            const float m = h.b1 * T.m[i] + (1.0f - h.b1) * g;
            const float v = h.b2 * T.v[i] + (1.0f - h.b2) * g * g;
            T.m[i] = m;
            T.v[i] = v;
            T.p[i] = p - h.lr * (m / bc1) / (sqrtf(v / bc2) + h.eps);
        }
        return;
    }
    // ---- Poincare-ball rows: one warp per row ----
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t D = T.cols, rows = T.numel / D;
    const int64_t r = (int64_t)blk * kAdamRowsPerBlock + warp;
    if (r >= rows) return;
    const float c = T.c;
    const Ball ball = make_ball(c);
    float* x = T.p + r * D;
    const float* g = T.g + r * D;
    float* m = T.m + r * D;
    float* v = T.v + r * D;
    auto wsum = [](float a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        return a;
    };
    float x2 = 0.0f;
    for (int64_t i = lane; i < D; i += 32) x2 = fmaf(x[i], x[i], x2);
    x2 = wsum(x2);
    const float lam = 2.0f / fmaxf(1.0f - c * x2, kMinNorm);
    const float il2 = 1.0f / (lam * lam);
    float rg2 = 0.0f;
    for (int64_t i = lane; i < D; i += 32) {
        const float rg = (g[i] + h.wd * x[i]) * il2;
        rg2 = fmaf(rg, rg, rg2);
    }
    rg2 = wsum(rg2);
    const float vn = h.b2 * v[lane < D ? lane : 0] + (1.0f - h.b2) * lam * lam * rg2;   // (all entries of a row are equal)
    const float vrow = __shfl_sync(0xffffffffu, vn, 0);
    const float scale = -h.lr / bc1 / (sqrtf(vrow / bc2) + h.eps);
    // y_pre = x + u,  u = scale * m_new;  accumulate the inner products the projection and the gyration need
    float y2 = 0.0f, xy = 0.0f, xm = 0.0f, ym = 0.0f;
    for (int64_t i = lane; i < D; i += 32) {
        const float rg = (g[i] + h.wd * x[i]) * il2;
        const float mn = h.b1 * m[i] + (1.0f - h.b1) * rg;
        const float y = fmaf(scale, mn, x[i]);
        y2 = fmaf(y, y, y2);
        xy = fmaf(x[i], y, xy);
        xm = fmaf(x[i], mn, xm);
        ym = fmaf(y, mn, ym);
    }
    y2 = wsum(y2); xy = wsum(xy); xm = wsum(xm); ym = wsum(ym);
    // project(y): y <- y * s with s = maxnorm / |y| when |y| > maxnorm
    const float yn = fmaxf(sqrtf(y2), kMinNorm);
    const float s = yn > ball.maxnorm ? ball.maxnorm / yn : 1.0f;
    const float Y2 = s * s * y2, XY = s * xy, YM = s * ym;
    // transport: gyr[u = y, v = -x] w  with  k = -c  (geoopt math.gyration), then * lambda_x / lambda_y
    //   a = -k^2 <u,w> |v|^2 - k <v,w> + 2 k^2 <u,v> <v,w>,  b = -k^2 <v,w> |u|^2 + k <u,w>,  d = 1 - 2k <u,v> + k^2 |u|^2 |v|^2
    const float k = -c, k2 = c * c;
    const float uv = -XY, uw = YM, vw = -xm, u2 = Y2, v2 = x2;
    const float a = -k2 * uw * v2 - k * vw + 2.0f * k2 * uv * vw;
    const float b = -k2 * vw * u2 + k * uw;
    const float d = fmaxf(1.0f - 2.0f * k * uv + k2 * u2 * v2, kMinNorm);
    const float lam_y = 2.0f / fmaxf(1.0f - c * Y2, kMinNorm);
    const float tr = lam / lam_y;
    const float ca = 2.0f * a / d, cb = 2.0f * b / d;
    for (int64_t i = lane; i < D; i += 32) {
        const float xi = x[i];
        const float rg = (g[i] + h.wd * xi) * il2;
        const float mn = h.b1 * m[i] + (1.0f - h.b1) * rg;
        const float yi = s * fmaf(scale, mn, xi);
        // w + 2 (a u + b v) / d  with u = y, v = -x
        m[i] = (mn + ca * yi - cb * xi) * tr;
        v[i] = vrow;
        x[i] = yi;
    }
}

}  // namespace hvae

using namespace hvae;

extern "C" size_t hvae_riemannian_adam_desc_bytes(void) { return sizeof(AdamTensor); }
extern "C" size_t hvae_riemannian_adam_hyper_bytes(void) { return sizeof(AdamHyper); }

// Fill descriptor `index` of a HOST table of n descriptors (hvae_riemannian_adam_desc_bytes each) and return the number
// of grid blocks the tensor needs.  c > 0: Poincare-ball rows of length cols; c == 0: Euclidean.  first_block = the sum
// of the block counts of the tensors before it.
extern "C" int64_t hvae_riemannian_adam_describe(void* host_table, int index, float* p, const float* g, float* m, float* v,
                                                 int64_t numel, int64_t cols, float c, int first_block) {
    if (!host_table || index < 0 || numel <= 0 || (c > 0.0f && (cols <= 0 || numel % cols))) return -1;
    AdamTensor& T = reinterpret_cast<AdamTensor*>(host_table)[index];
    T.p = p; T.g = g; T.m = m; T.v = v; T.numel = numel; T.cols = cols; T.c = c; T.first_block = first_block;
    if (c > 0.0f) return (numel / cols + kAdamRowsPerBlock - 1) / kAdamRowsPerBlock;
    return (numel + kAdamChunk - 1) / kAdamChunk;
}

// One optimizer step for every described tensor.  table_dev: the descriptor table copied to the device; hyper_dev:
// {lr, beta1, beta2, eps, weight_decay, step} as 6 floats in device memory (step = 1-based count of this update).
extern "C" int hvae_riemannian_adam_step_f32(const void* table_dev, int n_tensors, int total_blocks, const void* hyper_dev,
                                             void* stream) {
    if (!table_dev || !hyper_dev) return HVAE_EARG;
    if (n_tensors <= 0 || total_blocks <= 0) return HVAE_ESHAPE;
    k_riemannian_adam<<<total_blocks, kAdamThreads, 0, (cudaStream_t)stream>>>((const AdamTensor*)table_dev, n_tensors,
                                                                                  (const AdamHyper*)hyper_dev);
    return check_launch();
}
