/*
 * dcgp.h -- C ABI of libdcgp.so, the B200 (sm_100a) implementation of the conv-GP layer's
 * doubly-stochastic variational forward pass of kekeblom/DeepCGP.
 *
 * The reference has no native code and no FFI: its "operator interface" for this path
 * is the set of Python methods listed below.  Each entry point names the reference symbol
 * (file:line under the reference checkout; DS/ = submodules/Doubly-Stochastic-DGP/doubly_stochastic_dgp/)
 * whose arithmetic it replaces.  INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller owns all memory
 *     (inputs, outputs, workspaces); the library never allocates or frees device memory;
 *   - all calls are asynchronous and ordered on `stream` (a cudaStream_t passed as void*);
 *   - activations are float32, row-major, contiguous; parameters (Z, q_mu, q_sqrt, patch weights,
 *     kernel hyper-parameters) are float64 like the reference (gpflowrc:7); "M-only" linear algebra
 *     (Kuu, Cholesky, triangular inverses, KL) is float64, "T-sized" work (T = patches x images) is
 *     float32 / split-fp16 on the tensor cores with float32 accumulation;
 *   - images are NHWC flattened to [rows, H*W*C]; a patch vector is ordered (dy, dx, c), c fastest, and
 *     patch p = oy*OW + ox (views.py:32-44 + tf.extract_image_patches);
 *   - return value: DCGP_OK, or an error code below; text via dcgp_last_error().
 *     A non-positive-definite Kuu is reported LAPACK-style through the device int `*info`
 *     (0 = ok, k>0 = leading minor k not PD), mirroring tf.errors.InvalidArgumentError at
 *     conditionals.py:29 (caught at experiment.py:45-49); the call itself still returns DCGP_OK
 *     because the pivot is only known on the device -- the Python host checks `info`.
 */
#ifndef DCGP_H_
#define DCGP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCGP_OK 0
#define DCGP_ERR_ARG 1        /* bad argument / unsupported shape */
#define DCGP_ERR_WORKSPACE 2  /* workspace too small */
#define DCGP_ERR_NOT_PD 3     /* reserved for host-checked info */
#define DCGP_ERR_CUDA 4       /* CUDA runtime error, see dcgp_last_error() */

#define DCGP_LAYER_CONV 0      /* conv_gp/layers.py:ConvLayer (MultiOutputConvKernel)            */
#define DCGP_LAYER_SVGP_CONV 1 /* DS/layers.py:SVGP_Layer with conv_gp/kernels.py:ConvKernel     */

#define DCGP_ALGO_SIMT 0   /* fp32 CUDA-core path (validation path)                               */
#define DCGP_ALGO_TC 1     /* tcgen05 split-fp16 tensor-core path (product path)                  */

/* Geometry + hyper-parameters of one layer (constrained values, as GPflow hands them to the graph). */
typedef struct {
  int32_t kind;        /* DCGP_LAYER_*                                                            */
  int32_t H, W, C;     /* input image (views.py:20-30)                                            */
  int32_t f, s;        /* filter size, stride (dilation 1, VALID)                                 */
  int32_t M;           /* inducing patches                                                        */
  int32_t R;           /* gp_count (ConvLayer) / num_outputs (SVGP_Layer)                         */
  int32_t white;       /* arguments.py:33                                                         */
  double variance;     /* RBF variance  (models.py:115-117)                                       */
  double lengthscale;  /* RBF lengthscale                                                         */
  double jitter;       /* settings.jitter = 1e-3 (gpflowrc:11)                                    */
} dcgp_layer_desc;

const char* dcgp_last_error(void);
int dcgp_version(void);
/* number of CUDA kernels this library has launched so far in this process (bench.py reports it as gpu_launches) */
long long dcgp_launch_count(void);
/* Live kernel timing for roofline reports: when enabled, CUDA events are recorded on the launching stream around the
 * most recent conditional-GEMM (which = 0), Kuf (1), dK+dd GEMM (2) and dQ GEMM (3) kernel; dcgp_kernel_ms() waits for the end event and
 * returns the duration in milliseconds (-1 if none was recorded). */
void dcgp_set_kernel_timing(int on);
/* SMs the persistent tensor-core kernels launched from now on leave free (0 = use every SM): lets small latency-critical
 * kernels on another stream (a layer's chain rule / optimiser update / next-step prepare) advance underneath a long GEMM. */
void dcgp_set_reserved_sms(int n);
double dcgp_kernel_ms(int which);
/* Tensor-pipe flops that launch really issued (split products x the k-blocks not skipped as structurally zero), counted by the
 * launcher: the "executed" figure next to the algorithmic one in bench.py's roofline. */
double dcgp_kernel_tensor_flops(int which);
/* Precision of the three big T-sized GEMM families on the tensor-core path.  Every operand is held as two fp16 planes
 * x = hi + lo (22 bits); a GEMM issues per k-step: 3 = Al*Bh + Ah*Bl + Ah*Bh (fp32-class), 2 = Al*Bh + Ah*Bh (B rounded to
 * fp16), 4 = Ah*Bl + Ah*Bh (A rounded to fp16), 1 = Ah*Bh (fp16 operands); fp32 accumulation throughout.  cond = second stage of
 * the forward conditional (G_r = C_r^T a, A = a; the first stage a = Lm^-1 k, where cancellation happens, always uses 3),
 * dk = da = sum_r s_r a SP_r (A = a), dq = dS_r = a^T diag(s_r) a -- the large GEMMs of conditionals.py:31-65 and of its
 * derivative.  0 leaves a value unchanged.  Defaults: env DCGP_PROD_COND / DCGP_PROD_DK / DCGP_PROD_DQ, else the library's
 * built-in choice (3, 3, 3); dcgp_get_products reports the values in force. */
void dcgp_set_products(int cond, int dk, int dq);
void dcgp_get_products(int* cond_host, int* dk_host, int* dq_host);

/* Accumulation of the first stage a = Lm^-1 k (conditionals.py:29-32) on the tensor-core path.  The TMEM fp32 accumulator
 * rounds toward zero at every MMA; with 3 M / 16 MMAs per output that bias, amplified by the cancellation in Lm^-1 k, reaches
 * the 1e-4 gate for ill-conditioned Kuu at M >= 1024.  mode 1 (default): the low-order split products and three k-chunks of the
 * dominant product accumulate separately (four 128-column TMEM accumulators) and are added in fp32 round-to-nearest -- 5x less
 * bias for 1-2 % of the conditional's time; used whenever M is padded to a multiple of 128.  mode 0: one accumulator.
 * mode < 0: mode 1 from 1024 inducing points on.  env DCGP_PRECISE_STAGE1; dcgp_get_precise_stage1 returns -1, 0 or 1. */
void dcgp_set_precise_stage1(int mode);
int dcgp_get_precise_stage1(void);

/* views.py:56-68 FullView._patch_count/_patch_length/_out_image_size */
int dcgp_view_geometry(int H, int W, int C, int f, int s, int* OH_host, int* OW_host, int* P_host, int* L_host);

/* views.py:40-54 FullView.extract_patches_PNL (layout 0: [P,N,L]) / extract_patches (layout 1: [N,P,L]).
 * Only the API mirror and tests use it; the layer path never materialises patches. */
int dcgp_extract_patches(const float* X, int N, int H, int W, int C, int f, int s, int layout, float* out,
                         void* stream);

/* layers.py:18-21 MultiOutputConvKernel.Kuu; kernels.py:135-136,172-174 ConvKernel.Kzz + Kuu dispatch.
 * Kuu[M,M] = variance*exp(-0.5*|zi-zj|^2/l^2) + jitter*I, float64. */
int dcgp_kuu(const double* Z, int M, int L, double variance, double lengthscale, double jitter, double* Kuu,
             void* stream);

/* layers.py:23-32 MultiOutputConvKernel.Kuf: fused im2col + squared distance + RBF.
 * layout 0: out[P,M,N] float32 (the reference's layout); layout 1: out[N*P, ldo] with row t = n*P+p
 * (the layout the conditional GEMM consumes; ldo >= M, columns >= M are zero-filled).
 * ws: dcgp_kuf_workspace_bytes(M, L) (holds Z/lengthscale in float32). */
size_t dcgp_kuf_workspace_bytes(int M, int L);
int dcgp_kuf(const float* X, int N, int H, int W, int C, int f, int s, const double* Z, int M,
             double variance, double lengthscale, int layout, int ldo, float* out, void* ws, size_t ws_bytes,
             void* stream);

/* conditionals.py:29 tf.cholesky(Kmm): in-place lower Cholesky, float64, blocked left-looking.
 * ws: dcgp_cholesky_workspace_bytes(M). *info (device int) = 0 or failing minor. */
size_t dcgp_cholesky_workspace_bytes(int M);
int dcgp_cholesky(double* A, int M, void* ws, size_t ws_bytes, int* info, void* stream);

/* conditionals.py:6-67 conditional(Kmn, Kmm, Knn, f, full_cov=False, q_sqrt, white).
 * Kmn[P,M,N] f32, Kmm[M,M] f64, Knn[P,N] f32, f[M,R] f64, q_sqrt[R,M,M] f64 (lower triangle used).
 * Outputs fmean[N,P,R], fvar[R,P,N] float32 -- the reference's return layouts. */
size_t dcgp_conditional_workspace_bytes(int P, int M, int N, int R);
int dcgp_conditional(const float* Kmn, const double* Kmm, const float* Knn, const double* f,
                     const double* q_sqrt, int white, int P, int M, int N, int R, int algo,
                     float* fmean, float* fvar, void* ws, size_t ws_bytes, int* info, void* stream);

/* Per-step "M-only" work of one layer (everything that does not depend on the minibatch):
 * Kuu, Cholesky, triangular inverse, the stacked conditional operand W and the KL term.
 *   ConvLayer : layers.py:111 (Kuu), conditionals.py:29 (Cholesky), layers.py:137-152 (KL vs Kuu(Z_prior))
 *   SVGP_Layer: DS/layers.py:181-188 (build_cholesky_if_needed), :231-256 (KL)
 * Z[M,L], Z_prior[M,L] or NULL (= Z), q_mu[M,R], q_sqrt[R,M,M], float64.
 * `prep` is an opaque device buffer of dcgp_prepare_bytes() that dcgp_layer_apply consumes;
 * kl (device double) receives the layer's KL. */
size_t dcgp_prepare_bytes(const dcgp_layer_desc* d);
size_t dcgp_prepare_workspace_bytes(const dcgp_layer_desc* d);
/* Byte offsets, inside the workspace of the last dcgp_layer_prepare (algo = DCGP_ALGO_TC), of float64 results the host-side
 * chain rule re-uses: Kuu^-1 [M,M] (ld M), Lm^-1 and the prior's Lp^-1 (lower triangular, ld = *ld_inv >= M). */
int dcgp_prepare_workspace_layout(const dcgp_layer_desc* d, size_t* off_kinv, size_t* off_linv, size_t* off_lpinv, int* ld_inv);
/* Offset (bytes) inside the `prep` buffer of C_r = Lm^-1 L_r (L_r when whitened) as float32 [R, ld_b, ld_b] (tensor-core path),
 * left there by dcgp_layer_prepare for the host-side M-only chain rule. */
int dcgp_prepare_layout(const dcgp_layer_desc* d, size_t* off_b32, int* ld_b);
/* The same plus S_r = C_r C_r^T (float32 [R, ld, ld]); and the workspace layout plus the Cholesky factor Lm (lower triangle of
 * the [M, M] block at *off_lm, leading dimension M; the strict upper triangle is scratch). */
int dcgp_prepare_layout2(const dcgp_layer_desc* d, size_t* off_c32, size_t* off_s32, int* ld);
int dcgp_prepare_workspace_layout2(const dcgp_layer_desc* d, size_t* off_kinv, size_t* off_linv, size_t* off_lpinv, size_t* off_lm,
                                   int* ld_inv);
int dcgp_layer_prepare(const dcgp_layer_desc* d, const double* Z, const double* Z_prior, const double* q_mu,
                       const double* q_sqrt, int algo, void* prep, double* kl, void* ws, size_t ws_bytes,
                       int* info, void* stream);
/* Same, and records the CUDA event `fwd_ready_event` (a cudaEvent_t, may be NULL) on `stream` as soon as the operands
 * dcgp_layer_apply needs are complete; the KL and the backward-pass operands follow on the same stream.  Lets a host
 * that runs the prepare on a side stream start the layer's forward before the whole call has drained. */
int dcgp_layer_prepare_ev(const dcgp_layer_desc* d, const double* Z, const double* Z_prior, const double* q_mu,
                          const double* q_sqrt, int algo, void* prep, double* kl, void* ws, size_t ws_bytes, int* info,
                          void* fwd_ready_event, void* stream);
/* Same, with the kernel hyper-parameters read from device memory: hyp = {variance, lengthscale} (float64, may be NULL = use
 * the values in `d`).  For a host that queues the next step's prepare right behind the optimiser update
 * (experiment.py:97-108 runs them back to back inside one session.run) before it has read the updated values back. */
int dcgp_layer_prepare_hyp(const dcgp_layer_desc* d, const double* Z, const double* Z_prior, const double* q_mu,
                           const double* q_sqrt, int algo, void* prep, double* kl, void* ws, size_t ws_bytes, int* info,
                           void* fwd_ready_event, const double* hyp, void* stream);

/* The minibatch-sized work of one layer: layers.py:96-135 ConvLayer.conditional_ND or
 * DS/layers.py:191-229 SVGP_Layer.conditional_ND (with kernels.py:106-133 Kzx/Kdiag), followed by the
 * reparameterised sample of DS/layers.py:90-105 + DS/utils.py:41 when z is given.
 *   X[n_rows, H*W*C]; logical input is `n_rep` stacked copies of X (DS/dgp.py:63 tiles the first
 *   layer's input S times), so outputs have n_rows*n_rep rows, row = rep*n_rows + n.
 *   z, sample: [n_rows*n_rep, D] or NULL; mean, var: [n_rows*n_rep, D];  D = P*R (conv) or R (svgp).
 *   patch_weights[P] float64 (SVGP_CONV only; NULL = ones, kernels.py:26-28). */
size_t dcgp_apply_workspace_bytes(const dcgp_layer_desc* d, int n_rows, int n_rep);
int dcgp_layer_apply(const dcgp_layer_desc* d, const void* prep, const double* patch_weights, const float* X,
                     int n_rows, int n_rep, const float* z, int algo, float* mean, float* var, float* sample,
                     void* ws, size_t ws_bytes, void* stream);

/* ---- backward pass (the reference gets these from TensorFlow autodiff inside GPflow's AdamOptimizer,
 * experiment.py:84-108).  Tensor-core path only.
 *
 * dcgp_layer_backward: gradient of the objective through the minibatch-sized part of one layer.
 *   prep, apply_ws : the buffers the matching dcgp_layer_prepare / dcgp_layer_apply (algo = DCGP_ALGO_TC) calls used
 *                    (the forward leaves the kernel-matrix planes in apply_ws).
 *   g_mean, g_var  : [n_rows*n_rep, D] float32 gradients w.r.t. the layer's mean / var outputs.
 *   gX   [n_rows, H*W*C] float32, or NULL when the input needs no gradient (first layer)
 *   gQB  [(R+1)*Mp + 64, Mp] float64, Mp = M rounded up to 64: rows r*Mp + i (r = 1..R) hold dS_r[i, :] = sum_t s_r(t) a_t a_t^T
 *        (a_t = Lm^-1 k_t; var_r = knn - |a|^2 + a^T S_r a, S_r = C_r C_r^T), block 0 holds -sum_r dS_r; rows (R+1)*Mp + r
 *        hold dalpha[:, r] = sum_t g_mean_r(t) a_t   (mean_r = alpha_r^T a)
 *   gZ   [M, L] float64: direct path through Kuf;   gscal[4]: {d/dvariance, d/dlengthscale, -, -} direct paths
 *   gw   [P] float64 (SVGP_CONV only).  The M-only chain rule (Q, beta, KL -> Z, hyper-parameters, q_mu, q_sqrt) is
 *        small dense float64 algebra done by the host (deepcgp_b200/grad.py). */
size_t dcgp_backward_workspace_bytes(const dcgp_layer_desc* d, int n_rows, int n_rep);
int dcgp_layer_backward(const dcgp_layer_desc* d, const void* prep, const void* apply_ws, const double* Z,
                        const double* patch_weights, const float* X, int n_rows, int n_rep, const float* g_mean,
                        const float* g_var, float* gX, double* gQB, double* gZ, double* gscal, double* gw, void* ws,
                        size_t ws_bytes, void* stream);
/* The same in two parts (same arguments and workspace for both calls):  phases & 1 = everything the input gradient gX
 * needs (the critical path towards the layer below: dK / dd, dd . Z, col2im, Kdiag path); phases & 2 = the parameter-only
 * remainder (dQ, dbeta, dZ), which the host may queue after the backward of the layers below so that this layer's M-only
 * chain rule, optimiser update and next-step dcgp_layer_prepare overlap with it.  phases = 3 is dcgp_layer_backward. */
int dcgp_layer_backward_phases(const dcgp_layer_desc* d, const void* prep, const void* apply_ws, const double* Z,
                               const double* patch_weights, const float* X, int n_rows, int n_rep, const float* g_mean,
                               const float* g_var, float* gX, double* gQB, double* gZ, double* gscal, double* gw,
                               void* ws, size_t ws_bytes, int phases, void* stream);
/* The M-only chain rule of one layer: from dS_r, dalpha (rows of gQB as written by dcgp_layer_backward) and the direct-path
 * gradients gZ_direct / gscal to the gradients of the ELBO w.r.t. the layer's (constrained) parameters -- through
 * C_r = Lm^-1 L_r, S_r = C_r C_r^T, alpha = Lm^-1 q_mu, the Cholesky factor (Cholesky backward rule) and the RBF kernel --
 * plus the gradient of -kl_weight * KL (layers.py:137-147 / DS/layers.py:231-256).  What tf.gradients derives for the
 * minibatch-independent part of conditionals.py:29-58 when experiment.py:105-108 builds the optimiser.
 *   prep, prepare_ws : the buffers of this step's dcgp_layer_prepare (algo = DCGP_ALGO_TC): factors and tensor-core products
 *   hyp              : device {variance, lengthscale} to use instead of the descriptor's values (NULL = the descriptor's);
 *                      lets the call sit in a CUDA graph that is replayed while the host's copy is one step behind
 *   parts            : 1 = the parameter-only part (KL gradient; may be queued before the backward pass), 2 = the part that
 *                      needs gQB / gZ_direct / gscal, 3 = both.  `ws` carries state from part 1 to part 2.
 *   gZ [M,L], ghyp[2] = {d/dvariance, d/dlengthscale}, g_qmu [M,R], g_qsqrt [R,M,M] (lower triangular): float64 outputs. */
size_t dcgp_chain_rule_workspace_bytes(const dcgp_layer_desc* d);
int dcgp_layer_chain_rule(const dcgp_layer_desc* d, const void* prep, const void* prepare_ws, const double* Z,
                          const double* Z_prior, const double* q_mu, const double* q_sqrt, const double* hyp, const double* gQB,
                          const double* gZ_direct, const double* gscal, double kl_weight, int parts, double* gZ, double* ghyp,
                          double* g_qmu, double* g_qsqrt, void* ws, size_t ws_bytes, void* stream);
/* Batched C[b] = A[b] B[b]^T, float32 in / out, on the split-fp16 tcgen05 GEMM (22-bit products, fp32 accumulation): the
 * R-batched M^3 products of the M-only chain rule (what tf.gradients emits as batched MatMul ops).  A [batch or 1, m, k],
 * B [batch or 1, n, k], C [batch, m, n], all row-major contiguous; a batch stride (in elements) of 0 broadcasts the operand;
 * n % 4 == 0. */
size_t dcgp_bgemm_workspace_bytes(int batch, int m, int n, int k);
int dcgp_bgemm_nt(const float* A, const float* B, float* C, int batch, int m, int n, int k, long long a_bstride,
                  long long b_bstride, void* ws, size_t ws_bytes, void* stream);
/* gradient of coef * sum(varexp) w.r.t. Fmu, Fvar ([S*N, K] float32) */
int dcgp_multiclass_varexp_grad(const float* Fmu, const float* Fvar, const int32_t* Y, int S, int N, int K,
                                double epsilon, double coef, float* gmu, float* gvar, void* stream);
/* DS/utils.py:41 backward: g_mean = gF, g_var = gF * z / (2 sqrt(var + jitter)) */
int dcgp_sample_backward(const float* gF, const float* z, const float* var, size_t n, double jitter, float* g_mean,
                         float* g_var, void* stream);
/* Adam on a flat float64 vector (experiment.py:97-99; tf.train.AdamOptimizer update rule); maximize != 0 ascends */
int dcgp_adam(double* param, const double* grad, double* m, double* v, size_t n, double lr, double beta1, double beta2,
              double eps, int step, int maximize, void* stream);

/* kernels.py:117-133 ConvKernel.Kzx -> out[M,N] f32; kernels.py:106-115 ConvKernel.Kdiag -> out[N] f32. */
size_t dcgp_convkernel_kzx_workspace_bytes(const dcgp_layer_desc* d, int N);
int dcgp_convkernel_kzx(const dcgp_layer_desc* d, const double* Z, const double* patch_weights, const float* X,
                        int N, float* out, void* ws, size_t ws_bytes, void* stream);
int dcgp_convkernel_kdiag(const dcgp_layer_desc* d, const double* patch_weights, const float* X, int N,
                          float* out, void* stream);

/* tf.random_normal of DS/layers.py:104 as a counter-based generator (Philox-4x32-10 + Box-Muller): z[S, n_local, D] float32,
 * the draw for (sample s, GLOBAL image n0 + n, output d) depends only on (seed, step, layer, s, n0 + n, d, n_global), not
 * on the rank that holds the image -- image-sharded runs on 1/2/4/8 GPUs see identical noise. */
int dcgp_randn(float* z, int S, int n_local, int D, long long n_global, long long n0, unsigned long long seed,
               unsigned long long step, int layer, void* stream);

/* DS/utils.py:40-41 reparameterize (diag): out = mean + z*sqrt(var + jitter). */
int dcgp_reparameterize(const float* mean, const float* var, const float* z, size_t n, double jitter,
                        float* out, void* stream);

/* DS/utils.py:88-93 -> GPflow MultiClass(RobustMax(eps=1e-3)).variational_expectations, 20-point
 * Gauss-Hermite; Fmu,Fvar[S*N,K] f32, Y[N] int32 -> varexp[S*N] f64 and *sum (device double) = sum of
 * varexp (DS/dgp.py:90,94 take mean over S then sum over N: divide by S on the host side). */
int dcgp_multiclass_varexp(const float* Fmu, const float* Fvar, const int32_t* Y, int S, int N, int K,
                           double epsilon, double* varexp, double* sum, void* stream);
/* Prediction path (DS/dgp.py:116-126 predict_y / predict_density -> DS/utils.py:107-121 -> GPflow MultiClass(RobustMax)
 * predict_mean_and_var / predict_density): pmean[SN, K] = p_c (1-eps) + (1-p_c) eps/(K-1) with p_c = P(f_c largest) by the
 * same 20-point quadrature, pvar = pmean - pmean^2, logdens[SN] = log pmean[., Y] (any of the three may be NULL; Y [N]
 * is only read for logdens).  float64 outputs. */
int dcgp_multiclass_predict(const float* Fmu, const float* Fvar, const int32_t* Y, int S, int N, int K, double epsilon,
                            double* pmean, double* pvar, double* logdens, void* stream);

/* DS/dgp.py:92-98 _build_likelihood: elbo = sum_varexp/S * (num_data/N_global) - sum_l KL_l   (device doubles) */
int dcgp_elbo(const double* sum_varexp, int S, double num_data, double n_global, const double* kls,
              int n_layers, double* elbo, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DCGP_H_ */
