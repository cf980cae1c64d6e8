/* hvae_b200 — C ABI of the B200-native Poincare-ball VAE hot path.
 *
 * The reference (grisaitis/hyperbolic-vae) is pure Python: it has no FFI/plugin registry, so the
 * drop-in boundary it exposes is its Python class API.  This header is the C-ABI
 * layer UNDER that API: every entry point replaces the eager PyTorch/geoopt graph of one reference
 * call site (cited per function as file:line into /root/reference/hyperbolic_vae/).  The Python
 * package `hvae` binds these symbols with ctypes and wraps them as torch custom ops with analytic
 * backward (INTEGRATION.md shows the stub a reference maintainer would add).
 *
 * Conventions
 *  - All tensors: row-major contiguous float32 device pointers owned by the caller.  The library
 *    allocates nothing and keeps no state besides cached function attributes.
 *  - `stream` is a cudaStream_t passed as void*; calls are asynchronous, stream-ordered, never
 *    synchronize the device or the host, and are re-entrant.
 *  - Return 0 on success, a negative HVAE_E* code otherwise; nothing throws across the ABI.
 *  - sm_100a only.  There is no CPU fallback and no other-arch fallback by design.
 *  - c is the curvature MAGNITUDE as read from the manifold: float(manifold.c)
 *    (= softplus(isp_c) in fp32, not the constructor argument; SURVEY.md §7 "hard parts").
 */
#ifndef HVAE_B200_H
#define HVAE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HVAE_VERSION 100

#define HVAE_OK 0
#define HVAE_ESHAPE (-1)  /* unsupported or inconsistent shape */
#define HVAE_EALIGN (-2)  /* pointer alignment */
#define HVAE_EARCH (-3)   /* device is not sm_100 */
#define HVAE_ELAUNCH (-4) /* CUDA launch failure */
#define HVAE_EARG (-5)    /* null pointer / bad flag / workspace too small */

/* flags for the gyroplane (hyperplane-distance) op */
#define HVAE_GYRO_SIGNED 1u   /* keep the sign of <diff,a> (else abs)                         */
#define HVAE_GYRO_SQUARED 2u  /* d^2 (times sign(d) if signed)       layers.py:203-207        */
#define HVAE_GYRO_SCALED 4u   /* multiply by ||a||  (geoopt `scaled`, pvae `norm`)            */
#define HVAE_GYRO_PVAE 8u     /* normdist2plane clamps + projected (-p)(+)x  manifolds.py:41-65 */
#define HVAE_GYRO_RELU 16u    /* SIMT path only: out = max(out + bias, 0) - the ReLU that follows the decoder's gyroplane
                                 layer (scripts/_9_pvae_replicate.py Dec: relu(GeodesicLayer(z))) fused into the kernel */

int hvae_version(void);
const char* hvae_strerror(int code);
/* 0 if the current device is compute capability 10.x, else HVAE_EARCH */
int hvae_device_check(void);

/* ---- K3: expmap0 / logmap0  (layers.py:124-130; models/vae_hyperbolic.py:120,187; vae_one_b.py:218) -- */
int hvae_expmap0_fwd_f32(const float* u, float* y, int64_t rows, int64_t D, float c, void* stream);
int hvae_expmap0_bwd_f32(const float* u, const float* gy, float* gu, int64_t rows, int64_t D, float c, void* stream);
int hvae_logmap0_fwd_f32(const float* y, float* u, int64_t rows, int64_t D, float c, void* stream);
int hvae_logmap0_bwd_f32(const float* y, const float* gu, float* gy, int64_t rows, int64_t D, float c, void* stream);

/* ---- mobius_add (geoopt PoincareBall.mobius_add; manifolds.py:54) ---------------------------------- */
int hvae_mobius_add_fwd_f32(const float* x, const float* y, float* out, int64_t rows, int64_t D, float c,
                            int project, void* stream);
int hvae_mobius_add_bwd_f32(const float* x, const float* y, const float* gout, float* gx, float* gy,
                            int64_t rows, int64_t D, float c, int project, void* stream);

/* ---- generic two-point row ops: expmap(x,u), logmap(x,y), dist(x,y), transp(x,y,v) ------------------ */
int hvae_expmap_fwd_f32(const float* x, const float* u, float* out, int64_t rows, int64_t D, float c, void* stream);
int hvae_expmap_bwd_f32(const float* x, const float* u, const float* gout, float* gx, float* gu,
                        int64_t rows, int64_t D, float c, void* stream);
int hvae_logmap_fwd_f32(const float* x, const float* y, float* out, int64_t rows, int64_t D, float c, void* stream);
int hvae_logmap_bwd_f32(const float* x, const float* y, const float* gout, float* gx, float* gy,
                        int64_t rows, int64_t D, float c, void* stream);
int hvae_dist_fwd_f32(const float* x, const float* y, float* d, int64_t rows, int64_t D, float c, void* stream);
int hvae_dist_bwd_f32(const float* x, const float* y, const float* gd, float* gx, float* gy,
                      int64_t rows, int64_t D, float c, void* stream);

/* ---- K4: WrappedNormal.rsample  (distributions/wrapped_normal.py:66-74) --------------------------------
 * mu, sigma: (B,D); eps, z: (S,B,D).  z = project(mu (+) tanh(sqrt(c) lambda_mu ||u||/2) u/(sqrt(c)||u||)),
 * u = sigma*eps / lambda_0 * lambda_0/lambda_mu.  bwd reduces over S into gmu/gsigma (B,D). */
int hvae_wrapped_sample_fwd_f32(const float* mu, const float* sigma, const float* eps, float* z,
                                int64_t S, int64_t B, int64_t D, float c, void* stream);
int hvae_wrapped_sample_bwd_f32(const float* mu, const float* sigma, const float* eps, const float* gz,
                                float* gmu, float* gsigma, int64_t S, int64_t B, int64_t D, float c, void* stream);

/* ---- K5: WrappedNormal.log_prob  (distributions/wrapped_normal.py:76-89; manifolds.py:25-35) -------------
 * mu, sigma: (B,D); z: (S,B,D); logp: (S,B).  If mu == NULL the location is the origin and sigma is the
 * scalar `sigma0` (the prior WrappedNormal(0, prior_scale*1): vae_hyperbolic.py:194-199). */
int hvae_wrapped_logprob_fwd_f32(const float* mu, const float* sigma, float sigma0, const float* z, float* logp,
                                 int64_t S, int64_t B, int64_t D, float c, void* stream);
int hvae_wrapped_logprob_bwd_f32(const float* mu, const float* sigma, float sigma0, const float* z,
                                 const float* glogp, float* gmu, float* gsigma, float* gz,
                                 int64_t S, int64_t B, int64_t D, float c, void* stream);

/* ---- K4+K5 fused latent head: z = rsample(mu,sigma;eps), kl = log q(z|x) - log p(z) --------------------
 * (models/vae_hyperbolic.py:126-127,191-216; vae_hyperbolic_gyroplane_decoder.py:95-144).  S = 1.
 * kl: (B,).  bwd takes the upstream gradients of z (B,D; may be NULL) and kl (B,; may be NULL). */
int hvae_latent_head_fwd_f32(const float* mu, const float* sigma, const float* eps, float prior_scale,
                             float* z, float* kl, int64_t B, int64_t D, float c, void* stream);
int hvae_latent_head_bwd_f32(const float* mu, const float* sigma, const float* eps, float prior_scale,
                             const float* gz, const float* gkl, float* gmu, float* gsigma,
                             int64_t B, int64_t D, float c, void* stream);

/* ---- K2: gyroplane / hyperplane distance  (layers.py:193-210 & geoopt Distance2StereographicHyperplanes;
 *      layers.py:96-121 -> manifolds.py:41-65 with HVAE_GYRO_PVAE) ---------------------------------------
 * x: (B,D); p, a: (P,D) (a may alias p); bias: (P,) or NULL; out: (B,P).
 * out[b,j] = asinh(2 sqrt(c) <diff,a_j> / ((1 - c||diff||^2) ||a_j||)) / sqrt(c), diff = (-p_j) (+) x_b. */
int hvae_gyroplane_fwd_f32(const float* x, const float* p, const float* a, const float* bias, float* out,
                           int64_t B, int64_t D, int64_t P, float c, uint32_t flags, void* stream);
size_t hvae_gyroplane_bwd_workspace_bytes(int64_t B, int64_t D, int64_t P);
/* gp/ga: (P,D) (ga may be NULL when a aliases p: its gradient is added into gp); gbias: (P,) or NULL.
 * With HVAE_GYRO_RELU in flags, gout is the gradient of the ACTIVATED output: the backward recomputes the pre-activation
 * sign and masks it; use hvae_gyroplane_relu_bwd_f32 to pass the forward's bias (the plain entry assumes bias = NULL). */
int hvae_gyroplane_bwd_f32(const float* x, const float* p, const float* a, const float* gout,
                           float* gx, float* gp, float* ga, float* gbias,
                           int64_t B, int64_t D, int64_t P, float c, uint32_t flags,
                           void* workspace, size_t workspace_bytes, void* stream);
/* host-only: the grid of the backward's pair kernel for a problem - plane chunks sized so that the CTAs fill the resident
 * slots of the 148 SMs in whole waves (config 2: 23 chunks x 27 planes, 736 CTAs on 740 slots) */
int hvae_gyroplane_bwd_plan(int64_t B, int64_t D, int64_t P, int* planes_per_chunk, int* chunks, int* ctas);
int hvae_gyroplane_relu_bwd_f32(const float* x, const float* p, const float* a, const float* bias, const float* gout,
                                float* gx, float* gp, float* ga, float* gbias,
                                int64_t B, int64_t D, int64_t P, float c, uint32_t flags,
                                void* workspace, size_t workspace_bytes, void* stream);

/* ---- K1b: Riemannian layer weight prep  (layers.py:58-67) ------------------------------------------------
 * W: (P,F) `_weight`; beta: (P,) `_bias` (over_param=False) -> bpt = expmap0(W*beta) (P,F),
 * M = W * clamp_min(1 - c||bpt||^2, 1e-15) (P,F).  If bias_pt_in != NULL (over_param=True) it is used as bpt. */
int hvae_weight_prep_fwd_f32(const float* W, const float* beta, const float* bias_pt_in, float* bpt, float* M,
                             int64_t P, int64_t F, float c, void* stream);
int hvae_weight_prep_bwd_f32(const float* W, const float* beta, const float* bias_pt_in,
                             const float* gM, const float* gbpt, float* gW, float* gbeta, float* gbias_pt,
                             int64_t P, int64_t F, float c, void* stream);

/* ---- K1: Mobius matvec  (layers.py:145-147 -> geoopt mobius_matvec + project) ---------------------------
 * x: (B,F); M: (P,F); y: (B,P).  mx (B,P) is written when mx_out != NULL (saved for backward). */
int hvae_mobius_matvec_fwd_f32(const float* x, const float* M, float* y, float* mx_out,
                               int64_t B, int64_t F, int64_t P, float c, void* stream);
size_t hvae_mobius_matvec_bwd_workspace_bytes(int64_t B, int64_t F, int64_t P);
int hvae_mobius_matvec_bwd_f32(const float* x, const float* M, const float* mx, const float* gy,
                               float* gx, float* gM, int64_t B, int64_t F, int64_t P, float c,
                               void* workspace, size_t workspace_bytes, void* stream);

/* ---- K6/K7: HyperbolicRadius  (distributions/old_pvae_riemannian_normal.py:31,51 -> pvae, App. A.2) ------
 * sigma: (B,) per-row scale (already clamped to [0.1,7] by the caller as RiemannianNormal does). */
int hvae_hradius_lognorm_fwd_f32(const float* sigma, float* logZ, float* dlogZ_dsigma, int64_t B, int64_t dim,
                                 float c, void* stream);
/* r: (S,B) samples by rejection from a tangent hull, Philox4x32-10 stream (seed, counter).  Sample i uses
 * counter offset + (offset_dev ? *offset_dev : 0) + i; the device-side word lets a captured CUDA graph draw
 * fresh noise on every replay (the caller bumps it in-graph). */
int hvae_hradius_sample_f32(const float* sigma, float* r, int64_t S, int64_t B, int64_t dim, float c,
                            uint64_t seed, uint64_t offset, const int64_t* offset_dev, void* stream);
/* alpha (rows, D) ~ U(S^{D-1}) (pvae HypersphericalUniform.sample: a normalised standard normal vector), Philox in the
 * kernel; row i draws from counters offset + *offset_dev + i, a stream disjoint from the radius sampler's. */
int hvae_sphere_sample_f32(float* out, int64_t rows, int64_t D, uint64_t seed, uint64_t offset, const int64_t* offset_dev,
                           void* stream);
/* implicit reparameterisation: dr/dsigma = -(dF/dsigma)/(dF/dr) at the given (r, sigma); cdf optional */
int hvae_hradius_rgrad_f32(const float* sigma, const float* r, float* dr_dsigma, float* cdf,
                           int64_t S, int64_t B, int64_t dim, float c, void* stream);
/* expmap_polar: z = mu (+) tanh(sqrt(c) r/2) alpha/(sqrt(c)||alpha||)   (pvae manifolds; App. A.2) */
int hvae_expmap_polar_fwd_f32(const float* mu, const float* alpha, const float* r, float* z,
                              int64_t S, int64_t B, int64_t D, float c, void* stream);
int hvae_expmap_polar_bwd_f32(const float* mu, const float* alpha, const float* r, const float* gz,
                              float* gmu, float* gr, int64_t S, int64_t B, int64_t D, float c, void* stream);

/* ---- fused Monte-Carlo KL of RiemannianNormal(mu_b, sigma_b) against the origin prior RiemannianNormal(0, sigma_p)
 * (pvae RiemannianNormal.log_prob x2, objective of training/old_pvae_train.py:53-58).  sigma_q, logz_q: (B,);
 * sigma_p, logz_p: device scalars (no host sync); z: (S,B,D); kl: (S,B). */
int hvae_rn_kl_fwd_f32(const float* mu, const float* sigma_q, const float* logz_q, const float* z,
                       const float* sigma_p, const float* logz_p, float* kl, int64_t S, int64_t B, int64_t D,
                       float c, void* stream);
int hvae_rn_kl_bwd_f32(const float* mu, const float* sigma_q, const float* z, const float* sigma_p, const float* gkl,
                       float* gmu, float* gsigma_q, float* glogz_q, float* gz, int64_t S, int64_t B, int64_t D,
                       float c, void* stream);

/* ---- the same head fused with the sample (S = 1): z = expmap_polar(mu, alpha, r) and kl in ONE kernel; the backward
 * takes the decoder's gradient of z (B,D; may be NULL) and of kl (B,; may be NULL) and returns the TOTAL gradients of mu
 * and sigma_q: KL terms + the sample's path through expmap_polar + g_r * dr/dsigma (implicit reparameterisation,
 * dr_dsigma from hvae_hradius_rgrad_f32) - gkl * dlogZ/dsigma (dlogz_dsigma from hvae_hradius_lognorm_fwd_f32).
 * Replaces old_pvae_riemannian_normal.py:12-52 rsample + 2x log_prob and their ~10 autograd kernels. */
int hvae_rn_head_fwd_f32(const float* mu, const float* alpha, const float* r, const float* sigma_q,
                         const float* logz_q, const float* sigma_p, const float* logz_p, float* z, float* kl,
                         int64_t B, int64_t D, float c, void* stream);
int hvae_rn_head_bwd_f32(const float* mu, const float* alpha, const float* r, const float* sigma_q,
                         const float* sigma_p, const float* z, const float* dr_dsigma, const float* dlogz_dsigma,
                         const float* gz, const float* gkl, float* gmu, float* gsigma, int64_t B, int64_t D,
                         float c, void* stream);

/* ---- K1-TC / K2-TC: tcgen05 (bf16 operands, fp32 accumulate) forward paths for GEMM-sized shapes --------------
 * Same math as hvae_mobius_matvec_fwd_f32 / hvae_gyroplane_fwd_f32 (a == p), operands rounded to bf16: the
 * "bf16 GEMM mode" of BASELINE.json (1e-2 tolerance).  K (= F or D) must be a multiple of 8. */
size_t hvae_tc_workspace_bytes(int64_t B, int64_t K, int64_t P);
/* mx_out (B,P) optional: when given, the pre-activation is materialised (two-pass); when NULL, |mx_b|^2 comes from the
 * Gram matrix M^T M and the rescale + projection is fused into the GEMM epilogue (single pass, needs P % 8 == 0).
 * mxsq_out (B,) optional: |M x_b|^2 as the forward used it, the only extra state the tensor-core backward needs. */
int hvae_mobius_matvec_tc_fwd_f32(const float* x, const float* M, float* y, float* mx_out, float* mxsq_out, int64_t B,
                                  int64_t F, int64_t P, float c, void* workspace, size_t workspace_bytes, void* stream);
/* backward of the above on the tensor cores (autograd of geoopt mobius_matvec + project; reference call site
 * geoopt/layers/stereographic.py MobiusLinear / pvae MobiusLayer.forward): mx is recovered from y and mxsq,
 *   gmx = alpha gy + beta mx (bf16),  gx = gmx M + gxc x  (contraction over P),  gM = gmx^T x  (contraction over B).
 * B, F, P multiples of 8.  gx or gM may be NULL. */
size_t hvae_mobius_tc_bwd_workspace_bytes(int64_t B, int64_t F, int64_t P);
int hvae_mobius_matvec_tc_bwd_f32(const float* x, const float* M, const float* y, const float* mxsq, const float* gy,
                                  float* gx, float* gM, int64_t B, int64_t F, int64_t P, float c, void* workspace,
                                  size_t workspace_bytes, void* stream);
int hvae_gyroplane_tc_fwd_f32(const float* x, const float* p, const float* bias, float* out, int64_t B, int64_t D,
                              int64_t P, float c, uint32_t flags, void* workspace, size_t workspace_bytes, void* stream);
/* GeodesicLayer / normdist2plane with a != p (hyperbolic_vae/layers.py:96-121 -> manifolds.py:41-65) on the tensor
 * cores: <x,p_j> and <x,a_j> come from ONE N-concatenated cta_group::2 GEMM (the B tile of a CTA pair is 128 rows of p
 * and the same planes' 128 rows of a), the pair function with the reference's clamps and projection (HVAE_GYRO_PVAE) is
 * the epilogue.  p: (P,D) Mobius-subtracted point, a: (P,D) normal; bf16 operands, 1e-2 tolerance.  D % 8 == 0. */
size_t hvae_geodesic_tc_workspace_bytes(int64_t B, int64_t D, int64_t P);
int hvae_geodesic_tc_fwd_f32(const float* x, const float* p, const float* a, const float* bias, float* out, int64_t B,
                             int64_t D, int64_t P, float c, uint32_t flags, void* workspace, size_t workspace_bytes,
                             void* stream);
/* backward of the above (a == p): px = x p^T is recomputed by a GEMM, a tile kernel forms the pair gradients
 * (bf16 coefficient matrix + row / column scalar sums), then gx = CP p + rowcoef x and gp = CP^T x + colcoef p are two
 * more GEMMs with the axpy fused into their epilogues.  B, D, P multiples of 8; gx or gp may be NULL.  The bias
 * gradient is the column sum of gout (hvae_colsum_f32). */
size_t hvae_gyroplane_tc_bwd_workspace_bytes(int64_t B, int64_t D, int64_t P);
int hvae_gyroplane_tc_bwd_f32(const float* x, const float* p, const float* gout, float* gx, float* gp, int64_t B,
                              int64_t D, int64_t P, float c, uint32_t flags, void* workspace, size_t workspace_bytes,
                              void* stream);

/* ---- reconstruction-loss head: Bernoulli NLL with logits, summed over the feature axis (reference:
 * Bernoulli(logits).log_prob(x).sum(-1) in training/old_pvae_train.py:53-58; F.binary_cross_entropy_with_logits in
 * models/vae_hyperbolic_gyroplane_decoder.py).  logits (S,B,N), x (B,N) broadcast over S, nll / gnll (S,B). */
int hvae_bce_logits_rows_fwd_f32(const float* logits, const float* x, float* nll, int64_t S, int64_t B, int64_t N,
                                 void* stream);
int hvae_bce_logits_rows_bwd_f32(const float* logits, const float* x, const float* gnll, float* glogits, int64_t S,
                                 int64_t B, int64_t N, void* stream);

/* loss tail of the pvae objective (hyperbolic_vae/training/old_pvae_train.py:53-58): nll, kld (S,B) ->
 * out[3] = {recon + beta kl, recon = sum_b mean_s nll, kl = sum_b mean_s kld}; backward from gout[3]. */
int hvae_pvae_loss_fwd_f32(const float* nll, const float* kld, float* out, int64_t S, int64_t B, float beta, void* stream);
int hvae_pvae_loss_bwd_f32(const float* gout, float* gnll, float* gkld, int64_t S, int64_t B, float beta, void* stream);

/* posterior scale head: sigma = clamp(softplus(h) + eps, lo, hi) - the encoder's softplus(fc22(e)) + 1e-5
 * (scripts/_9_pvae_replicate.py) fused with RiemannianNormal's scale.clamp(0.1, 7) (old_pvae_riemannian_normal.py:30). */
int hvae_sigma_head_fwd_f32(const float* h, float* out, int64_t n, float eps, float lo, float hi, void* stream);
int hvae_sigma_head_bwd_f32(const float* h, const float* g, float* gh, int64_t n, float eps, float lo, float hi, void* stream);

/* ---- generic reconstruction heads, per-(sample,row) sums over the feature axis.  `in` (S,B,N) is the
 * decoder output (or its pre-sigmoid activation for the *_SIGMOID kinds: the decoder's final nn.Sigmoid is fused),
 * x (B,N) the target, broadcast over S; out (S,B).  reference: F.mse_loss(x_hat, x, "sum") models/vae_hyperbolic.py:219;
 * (x_hat - x)^2.sum(-1) models/vae_hyperbolic_rnaseq.py:107; RelaxedBernoulli(T, probs|logits).log_prob
 * models/vae_hyperbolic_gyroplane_decoder.py:121-122, models/vae_hyperbolic.py:224-225 (torch.distributions clamps). */
#define HVAE_RECON_MSE 1
#define HVAE_RECON_SIGMOID_MSE 2
#define HVAE_RECON_RB_LOGITS 3
#define HVAE_RECON_RB_PROBS 4
#define HVAE_RECON_RB_SIGMOID 5
int hvae_recon_rows_fwd_f32(const float* in, const float* x, float* out, int64_t S, int64_t B, int64_t N, int kind,
                            float temperature, void* stream);
int hvae_recon_rows_bwd_f32(const float* in, const float* x, const float* gout, float* gin, int64_t S, int64_t B, int64_t N,
                            int kind, float temperature, void* stream);

/* ---- data-parallel gradient exchange over NVLink peer memory (the reference is single-device:
 * training/trainer_mnist.py:19).  buf_ptrs_dev / pad_ptrs_dev: DEVICE arrays of `world` pointers - every rank's copy of
 * the flat gradient bucket and of a zero-initialised uint32 signal pad, as mapped into THIS rank's address space
 * (symmetric memory).  Sums elements [offset, offset+n) across ranks in place (fixed rank order: identical bits on all
 * ranks), scaled by `scale`; one kernel per rank, barriers on the pad slots [pad_slot_base, + hvae_allreduce_p2p_slots),
 * pad_slot_base >= 1: word 0 of a rank's pad is its error word, set to 1 when a barrier timed out (~4 s; a peer never
 * arrived) - the kernel then finishes instead of spinning forever and the host reads the word.  blocks: grid size
 * (1..128, 0 = default 64); it MUST be the same on every rank. */
int hvae_allreduce_p2p_slots(int world);
int hvae_allreduce_p2p_f32(const void* buf_ptrs_dev, const void* pad_ptrs_dev, int rank, int world, int64_t offset,
                           int64_t n, int pad_slot_base, float scale, int blocks, void* stream);
/* NVLS flavour of the same exchange: mc_ptr is the bucket's NVSwitch MULTICAST address on this rank; rank r sums slice r
 * with multimem.ld_reduce (the switch adds the W copies) and writes it to all copies with multimem.st: 1/W of the loads
 * and one NVLink round trip instead of W dependent ones.  blocks: 0 = sized from n (the same on every rank). */
int hvae_allreduce_nvls_f32(void* mc_ptr, const void* pad_ptrs_dev, int rank, int world, int64_t offset, int64_t n,
                            int pad_slot_base, float scale, int blocks, void* stream);

/* ---- f-1: geoopt.optim.RiemannianAdam as ONE multi-tensor kernel (reference call sites models/vae_hyperbolic.py:235-243,
 * ...gyroplane_decoder.py:173, ...rnaseq.py:139, vae_one_b.py:270).  The caller builds a HOST table of n descriptors
 * (hvae_riemannian_adam_desc_bytes each) with hvae_riemannian_adam_describe - parameter, gradient (e.g. its view in the
 * all-reduced flat bucket), exp_avg, exp_avg_sq, all fp32 contiguous; c > 0 marks Poincare-ball rows of length `cols`
 * (ManifoldParameter), c == 0 a Euclidean tensor - copies it to the device and launches one step.  describe returns the
 * tensor's block count (the next tensor's first_block is the running sum; the total is total_blocks) or -1.
 * hyper_dev: 6 floats in device memory {lr, beta1, beta2, eps, weight_decay, step}, step = 1-based count of this update. */
size_t hvae_riemannian_adam_desc_bytes(void);
size_t hvae_riemannian_adam_hyper_bytes(void);
int64_t hvae_riemannian_adam_describe(void* host_table, int index, float* p, const float* g, float* m, float* v, int64_t numel,
                                      int64_t cols, float c, int first_block);
int hvae_riemannian_adam_step_f32(const void* table_dev, int n_tensors, int total_blocks, const void* hyper_dev, void* stream);

/* ---- f-4: on-device input normalisation of the RNA-seq pipeline (hyperbolic_vae/datasets/jerby_arnon.py:97-106,
 * normalize_rnaseq): rows divided by their sum (x target: 1 = "sum_to_one", 1e6 = "sum_to_million"), or the per-column
 * z-score scipy.stats.zscore computes (population std; a constant column gives NaN as scipy does).  out may alias x. */
int hvae_rows_sum_normalize_f32(const float* x, float* out, int64_t R, int64_t C, float target, void* stream);
size_t hvae_cols_zscore_workspace_bytes(int64_t C);
int hvae_cols_zscore_f32(const float* x, float* out, float* mean_out, float* std_out, int64_t R, int64_t C, void* workspace,
                         size_t workspace_bytes, void* stream);

/* column sums of a row-major (R, C) matrix: the bias gradient of a dense layer (autograd of nn.Linear's bias). */
size_t hvae_colsum_workspace_bytes(int64_t C);
int hvae_colsum_f32(const float* x, float* out, int64_t R, int64_t C, void* workspace, size_t workspace_bytes,
                    void* stream);

/* ---- fp32-accurate dense GEMM on the tensor cores: the Euclidean trunk layers either side of the hyperbolic path
 * (reference: nn.Linear in hyperbolic_vae/models/vae_hyperbolic_*.py encoders/decoders and pvae Enc/Dec).
 *   C (M,N) = opA (M,K) . opB (N,K)^T  (+ bias[n]) (ReLU)
 * a_trans / b_trans != 0: the operand is stored (K,M) / (K,N).  Every fp32 operand is split into three bf16 pieces
 * (24 mantissa bits) and the six piece products down to 2^-16 are accumulated in fp32 (cf. cuBLAS BF16x9): the error
 * against a float64 product is that of an fp32 FMA GEMM (tests/test_gpu_trunk.py), so the 1e-5 budget holds. */
/* Split once, multiply several times (a dense layer uses each split in forward, dgrad and wgrad):
 *   hvae_split3_f32: src (rows, cols) fp32 -> dst (rows, 3*Cp) bf16, Cp = cols rounded up to 64 (hvae_split3_bytes).
 *   hvae_gemm_x3s_f32: the GEMM above on split operands.  a_mn / b_mn != 0: that operand is the split of a (K, M) /
 *   (K, N) matrix - the contraction runs over its rows and the tensor core reads it MN-major, no transpose is made. */
size_t hvae_split3_bytes(int64_t rows, int64_t cols);
int hvae_split3_f32(const float* src, void* dst, int64_t rows, int64_t cols, void* stream);
/* rows split (rows, 3*Cp) and transposed split (cols, 3*Rp) of the same matrix from one read; either may be NULL */
int hvae_split3_both_f32(const float* src, void* dst_rows, void* dst_t, int64_t rows, int64_t cols, void* stream);
size_t hvae_gemm_x3s_workspace_bytes(int64_t M, int64_t N);
int hvae_gemm_x3s_num_launches(int64_t M, int64_t N, int64_t K);
int hvae_gemm_x3s_f32(const void* As, int a_mn, const void* Bs, int b_mn, const float* bias, int relu, float* C,
                      int64_t M, int64_t N, int64_t K, void* workspace, size_t workspace_bytes, void* stream);
size_t hvae_gemm_x3_workspace_bytes(int64_t M, int64_t N, int64_t K);
int hvae_gemm_x3_num_launches(int64_t M, int64_t N, int64_t K);
int hvae_gemm_x3_f32(const float* A, int a_trans, const float* B, int b_trans, const float* bias, int relu, float* C,
                     int64_t M, int64_t N, int64_t K, void* workspace, size_t workspace_bytes, void* stream);

/* ---- the same dense layers on the fp16 tensor-core path: two-piece split with power-of-two row scales, three piece
 * products (hyperbolic-vae_b200/csrc/tc_x2.cu; reference: the nn.Linear layers above).  fp32 in, fp32 out; per-term
 * error <= 3 * 2^-22 (elements more than 2^28 below their row's largest magnitude keep an absolute error of 2^-39 of it).
 *   hvae_split2h_rows_f32: src (rows, cols) fp32 -> dst (rows, 2*Cp) fp16 [hi | lo] of row r times 2^e_r, Cp = cols
 *     rounded up to 64 (hvae_split2h_bytes); inv_scale (rows,) = 2^-e_r.
 *   hvae_split2h_both_f32: that, and/or the split of src^T (cols, 2*Rp) scaled per COLUMN of src (inv_cols (cols,)) -
 *     the layout a contraction over src's rows needs (weight gradients).  workspace: hvae_split2h_workspace_bytes.
 *   hvae_gemm_x2s_f32: C (M,N) = A (M,K) . B (N,K)^T (+ bias[n]) (ReLU) on split operands and their inverse scales.
 *   hvae_gemm_x2s_plan: the tile width (128..256), split-K factor and TMA ring depth chosen for a problem. */
size_t hvae_split2h_bytes(int64_t rows, int64_t cols);
size_t hvae_split2h_workspace_bytes(int64_t rows, int64_t cols);
/* general form: optional mask (same shape; elements where mask <= 0 count as zero - the ReLU backward of a fused
 * Linear + ReLU layer folded into the gradient's operand split), either layout optional, optional colsum (cols,) = column
 * sums of the (masked) src - the bias gradient, taken from the maximum pass's read. */
size_t hvae_split2h_ex_workspace_bytes(int64_t rows, int64_t cols);
int hvae_split2h_both_ex_f32(const float* src, const float* mask, void* dst_rows, float* inv_rows, void* dst_t,
                             float* inv_cols, float* colsum, int64_t rows, int64_t cols, void* workspace,
                             size_t workspace_bytes, void* stream);
int hvae_split2h_rows_f32(const float* src, void* dst, float* inv_scale, int64_t rows, int64_t cols, void* stream);
int hvae_split2h_both_f32(const float* src, void* dst_rows, float* inv_rows, void* dst_t, float* inv_cols, int64_t rows,
                          int64_t cols, void* workspace, size_t workspace_bytes, void* stream);
size_t hvae_gemm_x2s_workspace_bytes(int64_t M, int64_t N);
int hvae_gemm_x2s_num_launches(int64_t M, int64_t N, int64_t K);
int hvae_gemm_x2s_plan(int64_t M, int64_t N, int64_t K, int* bn, int* splits, int* stages);
int hvae_gemm_x2s_f32(const void* As, const float* inv_a, const void* Bs, const float* inv_b, const float* bias, int relu,
                      float* C, int64_t M, int64_t N, int64_t K, void* workspace, size_t workspace_bytes, void* stream);

/* ---- K2 at fp32 accuracy for GEMM-sized latent dims (D > 64: beyond the SIMT kernels) and ANY (a, p): a == p
 * (layers.py:193-210, geoopt Distance2StereographicHyperplanes) and GeodesicLayer / normdist2plane (layers.py:96-121 ->
 * manifolds.py:41-65, HVAE_GYRO_PVAE), forward and backward.  <x,p> and <x,a> come from the three-way split GEMM above
 * (2^-24 per term), the pair function and its gradient are applied elementwise, the parameter / input gradients are four more GEMMs
 * (hyperbolic-vae_b200/csrc/gyro_tc32.cu).  Same arguments as hvae_gyroplane_{fwd,bwd}_f32; a == NULL or a == p: a aliases p
 * (`two` = 0 in the workspace queries) and ga must be NULL. */
size_t hvae_gyroplane_tc32_fwd_workspace_bytes(int64_t B, int64_t D, int64_t P, int two);
size_t hvae_gyroplane_tc32_bwd_workspace_bytes(int64_t B, int64_t D, int64_t P, int two);
int hvae_gyroplane_tc32_fwd_f32(const float* x, const float* p, const float* a, const float* bias, float* out, int64_t B,
                              int64_t D, int64_t P, float c, uint32_t flags, void* workspace, size_t workspace_bytes,
                              void* stream);
int hvae_gyroplane_tc32_bwd_f32(const float* x, const float* p, const float* a, const float* gout, float* gx, float* gp,
                              float* ga, int64_t B, int64_t D, int64_t P, float c, uint32_t flags, void* workspace,
                              size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HVAE_B200_H */
