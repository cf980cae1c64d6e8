// Interface of the 2-CTA (cta_group::2) tcgen05 GEMM in tc_gemm2.cu: the big Mobius / gyroplane contractions
// (config 5: B = 2^20, F/D = 512, P = 4096) on 256x256 output tiles owned by a CTA pair.
#pragma once
#include <cuda_bf16.h>

#include "gyro_pair.cuh"
#include "hvae_common.cuh"

namespace hvae {
namespace tc2 {

// EPI_GYRO: any flag combination of the a == p gyroplane (launch_gemm2 picks the lean instantiation EPI_GYRO_LEAN for the
// plain signed distance, the general pair function otherwise)
// EPI_MOBIUS_F: the whole Mobius forward of an m-block in one unit of the A-resident schedule: first the tiles of x G
// (G = M^T M, the B2 operand) whose epilogue dots them with the resident bf16 rows of x - |M x_b|^2 = x_b^T G x_b - then,
// with the row's rescale + projection factor known, the tiles of x M^T scaled on the way out.  launch_gemm2 returns
// HVAE_ESHAPE when the problem is not eligible for the A-resident schedule (the caller keeps the three-kernel path).
enum { EPI_PLAIN = 0, EPI_GYRO = 1, EPI_ROWDOT = 2, EPI_MOBIUS = 3, EPI_GEO = 4, EPI_GYRO_BWD = 5, EPI_GYRO_LEAN = 6, EPI_MOBIUS_F = 7 };

constexpr int kPairM = 256;   // output rows per CTA pair (128 per CTA)
constexpr int kTileN = 256;   // accumulator columns per tile
constexpr int kCG = 4;        // epilogue column groups per tile (row-partial layouts: [(n_tile * kCG + cg)][M])

struct Params2 {
    float* D;               // (M, N) row-major output (split-K: S partial planes of M x N)
    int64_t M, N, K;        // N = output columns (GEO: planes; the B operand then has 2N rows in two maps)
    int splits;             // > 1: split-K, unit (tile, s) writes its partial tile to D + s * M * N
    int a_mn, b_mn;         // operand stored contraction-major-OUTER: buffer rows = contraction index, columns = M / N index
    int64_t a_pitch, b_pitch;  // row pitch of the operand buffers in elements (0 = dense)
    const float* g;         // GYRO_BWD: (M, N) upstream gradient (read by TMA);  D16: (M, N) bf16 output CP = dL/d<x,p>
    __nv_bfloat16* D16;
    float* srow;            // GYRO_BWD: [n_tiles * kCG][M] partial row sums of CP * (v_j / u_j)
    float* vcol;            // GYRO_BWD: [ceil(M / 32)][N] partial column sums of CP * w_b  (one row per 32-row block)
    const float* rowscale;  // PLAIN: optional (M,); MOBIUS: required (M,)
    const float* axpy_x;    // PLAIN: optional (M, N) fp32;  D = acc * rowscale + axpy_coef[m] * axpy_x[m][n]
    const float* axpy_coef; //        (M,) or NULL (= 1)
    float* rowsq;           // PLAIN: optional [n_tiles * kCG][M] partial sums of acc^2
    const float* x2;        // GYRO / GEO / MOBIUS_F: (M,) |x|^2
    float* mxsq_out;        // MOBIUS_F: optional (M,) |M x_b|^2 (the backward's saved row statistic);  gp = the ball
    const float* p2;        // GYRO / GEO: (N,) |p|^2
    const float* pa;        // GEO: (N,) <p_j, a_j>
    const float* an;        // GEO: (N,) |a_j|
    const float* bias;      // GYRO / GEO: optional (N,)
    GyroParams gp;
    const float* xrow;      // ROWDOT: (M, N) fp32 rows dotted with the accumulator rows
    float* rowdot;          // ROWDOT: [n_tiles * kCG][M] partial sums of acc * xrow
#ifdef HVAE_EXPERIMENT
    int dbg;                // experiment build only: 1 = skip the global stores, 2 = skip the whole drain
#endif
};

// number of row-partial planes a (.., N) problem produces (rowsq / rowdot)
inline int row_partials(int64_t N) { return (int)((N + kTileN - 1) / kTileN) * kCG; }

// A (M, K) bf16 K-major, B (N, K) bf16 K-major (GEO: B = p rows, B2 = a rows, both (N, K)).  Returns HVAE_OK / error code.
int launch_gemm2(int epi, const __nv_bfloat16* A, const __nv_bfloat16* B, const __nv_bfloat16* B2, const Params2& prm,
                 cudaStream_t s);

}  // namespace tc2
}  // namespace hvae
