// f-4: on-device input normalisation of the RNA-seq pipeline, so a batch goes host -> HBM once, raw,
// and is normalised where the step reads it.
// reference: hyperbolic_vae/datasets/jerby_arnon.py:97-106 (normalize_rnaseq):
//   "sum_to_one" / "sum_to_million": every row (cell) divided by its sum (x 1e6)
//   "z_score": scipy.stats.zscore per COLUMN (gene) over the cells: (x - mean) / std, population std (ddof = 0);
//              a constant column gives 0/0 = NaN, as scipy does
#include "hvae_common.cuh"

namespace hvae {

__global__ void __launch_bounds__(256)
k_rows_sum_normalize(const float* __restrict__ x, float* __restrict__ out, int64_t R, int64_t C, float target) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < R; r += nw) {
        double s = 0.0;   // counts sum to ~1e6 over 2e4 genes: keep the row sum exact to fp32 rounding
        for (int64_t c = lane; c < C; c += 32) s += (double)__ldg(x + r * C + c);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float k = (float)((double)target / s);
        for (int64_t c = lane; c < C; c += 32) out[r * C + c] = __ldg(x + r * C + c) * k;
    }
}

// pass 1: per (row chunk, column) partial sum and sum of squares about a per-column pivot (the column's first value:
// removes the mean^2 cancellation), double accumulators; block = 32 columns x 8 row lanes
constexpr int kZChunks = 32;
__global__ void __launch_bounds__(256)
k_cols_moments(const float* __restrict__ x, double* __restrict__ part, int64_t R, int64_t C) {
    __shared__ double s1[8][33], s2[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t c = (int64_t)blockIdx.x * 32 + tx;
    const int64_t per = (R + kZChunks - 1) / kZChunks;
    const int64_t r0 = (int64_t)blockIdx.y * per, r1 = (r0 + per < R) ? r0 + per : R;
    double a = 0.0, b = 0.0;
    if (c < C) {
        const double piv = (double)__ldg(x + c);
        for (int64_t r = r0 + ty; r < r1; r += 8) {
            const double d = (double)__ldg(x + r * C + c) - piv;
            a += d;
            b += d * d;
        }
    }
    s1[ty][tx] = a;
    s2[ty][tx] = b;
    __syncthreads();
    if (ty == 0 && c < C) {
        double A = 0.0, B = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { A += s1[i][tx]; B += s2[i][tx]; }
        part[((int64_t)blockIdx.y * C + c) * 2] = A;
        part[((int64_t)blockIdx.y * C + c) * 2 + 1] = B;
    }
}
// pass 2: out = (x - mean) / std
__global__ void __launch_bounds__(256)
k_cols_zscore_apply(const float* __restrict__ x, const double* __restrict__ part, float* __restrict__ out, float* __restrict__ mean_out,
                    float* __restrict__ std_out, int64_t R, int64_t C) {
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t c = (int64_t)blockIdx.x * 32 + tx;
    if (c >= C) return;
    double A = 0.0, B = 0.0;
    for (int i = 0; i < kZChunks; ++i) { A += part[((int64_t)i * C + c) * 2]; B += part[((int64_t)i * C + c) * 2 + 1]; }
    const double piv = (double)__ldg(x + c);
    const double md = A / (double)R;                      // mean - pivot
    const double var = fmax(B / (double)R - md * md, 0.0);
    const float mean = (float)(piv + md), sd = (float)sqrt(var);
    if (blockIdx.y == 0 && ty == 0) {
        if (mean_out) mean_out[c] = mean;
        if (std_out) std_out[c] = sd;
    }
    const int64_t per = (R + gridDim.y - 1) / gridDim.y;
    const int64_t r0 = (int64_t)blockIdx.y * per, r1 = (r0 + per < R) ? r0 + per : R;
    for (int64_t r = r0 + ty; r < r1; r += 8) out[r * C + c] = (__ldg(x + r * C + c) - mean) / sd;
}

}  // namespace hvae

using namespace hvae;

// out[r, :] = x[r, :] / sum(x[r, :]) * target   (target 1 = "sum_to_one", 1e6 = "sum_to_million"); out may alias x
extern "C" int hvae_rows_sum_normalize_f32(const float* x, float* out, int64_t R, int64_t C, float target, void* stream) {
    if (R <= 0 || C <= 0) return HVAE_ESHAPE;
    if (!x || !out) return HVAE_EARG;
    const int64_t want = (R + 7) / 8, cap = (int64_t)kNumSMs * 16;
    k_rows_sum_normalize<<<(unsigned)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(x, out, R, C, target);
    return check_launch();
}

extern "C" size_t hvae_cols_zscore_workspace_bytes(int64_t C) { return C > 0 ? (size_t)kZChunks * C * 2 * sizeof(double) : 0; }

// out[:, c] = (x[:, c] - mean_c) / std_c, population std over the R rows; mean_out / std_out (C,) optional; out may alias x
extern "C" int hvae_cols_zscore_f32(const float* x, float* out, float* mean_out, float* std_out, int64_t R, int64_t C,
                                    void* workspace, size_t workspace_bytes, void* stream) {
    if (R <= 0 || C <= 0) return HVAE_ESHAPE;
    if (!x || !out || !workspace || workspace_bytes < hvae_cols_zscore_workspace_bytes(C)) return HVAE_EARG;
    cudaStream_t s = (cudaStream_t)stream;
    dim3 grid((unsigned)((C + 31) / 32), kZChunks);
    k_cols_moments<<<grid, 256, 0, s>>>(x, (double*)workspace, R, C);
    k_cols_zscore_apply<<<grid, 256, 0, s>>>(x, (const double*)workspace, out, mean_out, std_out, R, C);
    return check_launch();
}
