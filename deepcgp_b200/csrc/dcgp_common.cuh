// Shared declarations for libdcgp.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dcgp.h"

namespace dcgp {

void set_error(const char* fmt, ...);
int check_launch(const char* what, int n_launched = 1);   // also counts kernel launches (dcgp_launch_count)
long long launch_count();

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ------------------------------------------------------------------ float64 M-only toolbox (dcgp_f64.cu)
constexpr int NB = 64;  // Cholesky / triangular-inverse block size

// C[m,n] = alpha * op(A) * op(B) + beta * C, row-major, batched over blockIdx.z.
// lowerA / lowerB: treat the STORED operand as lower triangular (entries above the diagonal read as 0).
struct GemmF64 {
  int m, n, k;
  const double* A; int lda; int transA; int lowerA; long long strideA;
  const double* B; int ldb; int transB; int lowerB; long long strideB;
  double* C; int ldc; long long strideC;
  double alpha, beta;
  int batch;
  int lowerC;   // only output tiles that touch the lower triangle are computed (symmetric rank-k updates)
  int ksplit;   // set by gemm_f64(): > 1 = split-k over blockIdx.z with atomic accumulation into a zeroed C
};
int gemm_f64(const GemmF64& g, cudaStream_t st);

int rbf_sym_f64(const double* Z, int M, int L, double variance, double lengthscale, double jitter, double* K,
                cudaStream_t st, const double* hyp = nullptr);
// In-place blocked left-looking Cholesky (lower). invD: [ceil(M/NB)] inverses of the diagonal blocks (NB*NB each).
size_t potrf_ws_bytes(int M);
int potrf_f64(double* A, int lda, int M, double* invD, int* info, cudaStream_t st);
// Linv (ld = Mq, Mq = NB * next_pow2(ceil(M/NB))) = L^-1 by recursive doubling; needs L factor + invD from potrf.
int trtri_pad(int M);
size_t trtri_ws_bytes(int M);
int trtri_f64(const double* L, int lda, int M, const double* invD, double* Linv, void* ws, cudaStream_t st);
// out[slot] = sum of squares / sum of log(diag^2) (deterministic single-CTA reductions)
int sumsq_f64(const double* x, long long rows, int cols, int ld, int lower_period, double* out, cudaStream_t st);
int logdiag2_f64(const double* A, int lda, int M, int batch, long long stride, double* out, cudaStream_t st);

}  // namespace dcgp
