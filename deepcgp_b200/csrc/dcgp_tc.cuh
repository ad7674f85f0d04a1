// Tensor-core (tcgen05 / TMEM / TMA) path: split-fp16 operands and the kernels that consume them (dcgp_tc.cu).
#pragma once
#include "dcgp_kernels.cuh"

namespace dcgp {

// Split-fp16 planes of one layer's minibatch-independent operands.
//   x ~= (hi + lo) / 2^e  with hi = fp16(x * 2^e), lo = fp16(x * 2^e - hi): ~22 mantissa bits per operand, so that
//   A*B ~= hi*hi + hi*lo + lo*hi on the fp16 tensor pipe with fp32 accumulation reproduces an fp32 GEMM.
struct TcPrep {
  int M, Mp, R, L, Lp, LpT;   // Lp = ceil(L/64)*64; LpT = ceil((L+1)/64)*64 (room for the ones row of ZT)
  void *Wh, *Wl;     // [(R+1)*Mp, Mp] fp16
  void *Wmh, *Wml;   // [64, Mp] fp16 (mean rows, zero padded)
  void *Zh, *Zl;     // [Mp, Lp] fp16 (Z / lengthscale)
  float* zz;         // [Mp] |z/ls|^2 (fp32, from fp64)
  float* scal;       // [32] device {scale, 1/scale} pairs: W blocks, mean rows, q_sqrt^T, G, Lp^-1, QP, kXScale, B_r, beta
  float* mx;         // [8] running max |x| per slot
  void *QTh, *QTl;   // [R*Mp, Mp] fp16: transposed q_sqrt planes (A operand of W_r = L_r^T G, B operand of Lp^-1 L_r)
  void *Gh, *Gl;     // [Mp, Mp]
  void *Lph, *Lpl;   // [Mp, Mp]
  float* Wr32;       // [R*Mp, Mp] fp32 W_r (tensor-core product)
  // ---- operands of the backward pass
  void *BRh, *BRl;   // [R*Mp, Mp]   B_r = W_r^T = G L_r
  float* Br32;       // [R*Mp, Mp]   B_r in fp32 (read back by the host's M-only chain rule, dcgp_prepare_layout)
  float* Qr32;       // [R*Mp, Mp]   Q_r = B_r B_r^T
  void *QBh, *QBl;   // [R*Mp + 256, Mp]  QP_r = 2 (Q_r - Q_0) stacked over r = 1..R (B operand of the dK GEMM), zero padded
  float* beta32;     // [Mp, 64]     beta[m, r] (fp32), zero padded
  void *ZTh, *ZTl;   // [LpT, Mp]    bf16 planes of (Z / lengthscale)^T * kXScale; row L = kXScale (ones row: DDZ[:, L] = row sums of dd)
  void *BTh, *BTl;   // [Mp, 64]     fp16 planes of alpha[m, r] (B operand of the mean tile of the da GEMM)
  void *LTh, *LTl;   // [Mp + 256, Mp] fp16 planes of Lm^-T (row m, column j: Lm^-1[j, m]; B operand of dK = da Lm^-1), zero padded
  size_t bytes;
};
void tc_carve_prep(TcPrep& t, int M, int Mp, int R, int L, void* buf);
int tc_pack_operands(const TcPrep& t, const double* Linv, int ldl, const double* Wr, const double* beta, int M, int Mp,
                     int R, cudaStream_t st);
int tc_build_operands(const TcPrep& t, const double* Linv, int ldl, const double* G, int ldg, int g_is_linv,
                      const double* Lpinv, int ldp, const double* q_sqrt, const double* beta, double* trace_out,
                      const double* Kinv, int parts, int chained, const double* alpha, cudaStream_t st);
// parts: 1 = forward operands, 2 = KL trace + backward operands.  chained: the W blocks 1..R hold C_r^T (C_r = Lm^-1 L_r, or
// L_r when whitened) and the mean rows alpha^T (alpha = Lm^-1 q_mu, or q_mu) for tc_cond_chained, instead of W_r / beta^T.
int tc_pack_z(const TcPrep& t, const double* Z, int M, int L, double inv_ls, cudaStream_t st, const double* hyp = nullptr);

struct TcCondWork {
  TcPrep prep;    // only carved by tc_carve_cond (the conditional() API mirror owns its operands)
  void *Kh, *Kl;  // [Tpad, Mp] fp16 planes of the kernel-matrix rows
  float* kscal;   // device: [0] = K scale, [1] = 1 / K scale
  void *Ah, *Al;  // [Tpad, Mp] fp16 planes of a = Lm^-1 k (chained conditional, ConvLayer only; null otherwise)
  float* ascal;   // device: {scale, 1/scale} of the a planes
  size_t Tpad;
  size_t bytes;
};
void tc_carve_cond(TcCondWork& w, int M, int Mp, int R, size_t T, void* buf);
int tc_split_rows(const float* Kt, int T, int Mp, const TcCondWork& w, cudaStream_t st);
int tc_cond(const TcPrep& prep, const TcCondWork& w, int T, int Mp, int R, float* acc, float* mean, cudaStream_t st);
// chained form (prep.chained): a = K Lm^-T with |a|^2 -> acc[:, 0] and the a planes, then G_r = a C_r (upper-triangular operand)
int tc_cond_chained(const TcPrep& prep, const TcCondWork& w, int T, int Mp, int R, float a_bound, float* acc, float* mean, cudaStream_t st,
                    const double* patch_weights = nullptr, int P = 0);   // patch_weights: the bound on |a| is scaled by max|w| (image-level rows)
bool tc_forward_chained();   // DCGP_FWD_CHAINED != 0 (default on)

// Generic batched NT GEMM on split-fp16 planes: C[b][i,j] = sum_k A[b*a_batch_rows + i, k] * B[b*b_batch_rows + j, k].
// Planes are row-major [rows_total, k_pad] fp16 (k_pad % 64 == 0, m_pad % 128 == 0, n_pad % 64 == 0, zero padded).
struct TcGemm {
  const void *Ah, *Al, *Bh, *Bl;
  long long a_rows_total, b_rows_total;
  int a_batch_rows, b_batch_rows, batch;
  int m, n, m_pad, n_pad, k_pad;
  const float *a_scal, *b_scal;   // device {scale, 1/scale}
  float* C; long long c_batch_stride; int ldc;   // optional fp32 output (unscaled values)
  double* sq_out;                                // optional: += sum of squares of C
  float* absmax_out;                             // optional: atomic max |C| (zero it first)
  int bf16;                                      // planes are bf16 hi/lo (unscaled values allowed) instead of fp16
  int splits; long long c_split_stride;          // split-K: partial C per split (caller reduces); splits <= 1 = off
  int nprod;                                     // split products per k-step (0 = 3)
  int kblocked;                                  // both operands are k-blocked planes [k_pad/64][rows][64] instead of row-major [rows, k_pad]
};
size_t tc_bgemm_workspace_bytes(int batch, int m, int n, int k);
int tc_bgemm_nt(const float* A, const float* B, float* C, int batch, int m, int n, int k, long long a_bstride, long long b_bstride,
                void* ws, cudaStream_t st);
// the same with explicit leading dimensions / batch strides (in elements; strides multiples of the leading dimension; ldc % 4 == 0)
int tc_bgemm_nt_ld(const float* A, int lda, long long a_bstride, const float* B, int ldb, long long b_bstride, float* C, int ldc,
                   long long c_bstride, int batch, int m, int n, int k, void* ws, cudaStream_t st);
int tc_gemm_splits(const TcGemm& g);             // number of splits tc_gemm will really use
int tc_gemm(const TcGemm& g, cudaStream_t st);

struct TcApplyWork {
  TcCondWork kk;   // planes of the patch-level kernel matrix (Kuf rows)
  TcCondWork kz;   // svgp only: planes of the image-level Kzx rows
  size_t bytes;
};
int tc_kuf(const TcPrep& prep, const View& v, const float* X, int n_rows, float variance, float inv_ls, const float* kscal,
           void* Kh, void* Kl, cudaStream_t st);
void tc_carve_apply(TcApplyWork& a, int kind, int M, int Mp, int R, int L, size_t Tk, size_t T, void* buf);
int tc_layer_apply(const dcgp_layer_desc* d, const View& v, const TcPrep& prep, const TcApplyWork& a, const float* zs,
                   float* Kt32, const double* patch_weights, const float* X, int n_rows, float* Kzx, float* acc,
                   float* mean_t, cudaStream_t st);

void tc_set_reserved_sms(int n);
void tc_set_timing(int on);
double tc_kernel_ms(int which);   // 0 = conditional GEMM, 1 = Kuf, 2 = dK (+dd) GEMM, 3 = dQ GEMM (last launch of each)
double tc_kernel_flops(int which);   // executed tensor-pipe flops of that launch (counted by the launcher)

// split products per k-step of the three big T-sized GEMM families (1..4, see TcParams::nprod in dcgp_tc.cu)
struct TcProducts { int cond, dk, dq; };
constexpr int kDefaultProdCond = 3, kDefaultProdDk = 3, kDefaultProdDq = 3;   // measured: tests/test_gpu_bench_size.py, DESIGN.md 'Precision'
const TcProducts& tc_products();
void tc_set_products(int cond, int dk, int dq);   // 0 leaves a value unchanged
constexpr int kPreciseAutoM = 1024;               // mode < 0: stage 1 of the conditional on four accumulators from this M on
void tc_set_precise_stage1(int mode);             // 1 always (default), 0 never, < 0 from kPreciseAutoM on
int tc_get_precise_stage1();
bool tc_precise_stage1(int Mp);

// Workspace of the backward pass of one layer (see dcgp_tc_bwd.inc)
struct TcBwdWork {
  int Jp, Lp, splits2, splits4, splitsb;
  size_t Tpad, Tkpad;
  float *gm, *s, *sT, *gmT32, *gknn, *scal, *dK32, *part2, *partb, *rowdot, *DDZ, *part4;
  double *DDX, *redpart;
  int n_redpart;
  void *GTh, *GTl, *GMh, *GMl, *KTh, *KTl, *DAh, *DAl, *Dh, *Dl, *DTh, *DTl, *PTh, *PTl;   // D*, DT*, PT*: bf16 planes; KT*: a^T
  size_t bytes;
};
void tc_carve_bwd(TcBwdWork& b, int kind, int M, int Mp, int R, int L, size_t Tk, size_t T, int P, void* buf);
int tc_layer_backward(const dcgp_layer_desc* d, const View& v, const TcPrep& prep, const TcApplyWork& a, const TcBwdWork& b,
                      const double* Z, const double* patch_weights, const float* X, int n_rows, int n_rep, const float* g_mean,
                      const float* g_var, float* gX, double* gQB, double* gZ, double* gscal, double* gw, int phases, cudaStream_t st);

// ---- dcgp_chain.cu: the M-only chain rule of one layer (native counterpart of what tf.gradients derives for the minibatch-
// independent part of the layer; see the file header)
struct ChainInputs {
  const double *Kinv, *Li, *Lpinv, *Lm, *alpha;   // Kinv ld M; Li, Lpinv ld ldi; Lm ld M (lower); alpha [M,R] (non-whitened)
  int ldi;
  const float *C32, *S32, *Ct32;                  // [R,Mp,Mp]: C_r, S_r, C_r^T
};
size_t chain_rule_workspace_bytes(int M, int R, int L);
int chain_rule(const dcgp_layer_desc* d, const ChainInputs& in, const double* Z, const double* Z_prior, const double* q_mu,
               const double* q_sqrt, const double* hyp_dev, const double* gQB, const double* gZ_direct, const double* gscal,
               double kl_weight, int parts, double* gZ, double* ghyp, double* g_qmu, double* g_qsqrt, void* ws, cudaStream_t st);

}  // namespace dcgp
