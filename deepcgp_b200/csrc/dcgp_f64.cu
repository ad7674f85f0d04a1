// float64 "M-only" toolbox: everything whose size depends only on the number of inducing patches M.
//   Kuu (layers.py:18-21), Cholesky (conditionals.py:29), L^-1, and the small dense products that
//   build the stacked conditional operand W and the KL terms.  These are latency-bound at M <= 1024
//   (Cholesky at M=512 is 45 MFLOP); tcgen05 has no fp64 kind, so they run on the fp64 CUDA cores.
#include <stdarg.h>
#include <string.h>

#include "dcgp_common.cuh"

namespace dcgp {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }
static long long g_launches = 0;
long long launch_count() { return g_launches; }
int check_launch(const char* what, int n_launched) {
  g_launches += n_launched;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return DCGP_ERR_CUDA;
  }
  return DCGP_OK;
}

// ------------------------------------------------------------------------------------------ GEMM
// 64x64 output tile, 16-deep k-slices, 4x4 outputs per thread; the next slice's global loads are issued into registers
// before the current slice is consumed (software double buffering) because these GEMMs are small and latency-bound.
constexpr int KC = 32;    // k-slice depth staged per iteration
__device__ __forceinline__ void gemm_f64_fetch(const GemmF64& g, const double* __restrict__ A, const double* __restrict__ B,
                                               int i0, int j0, int k0, int tid, double (&ra)[8], double (&rb)[8]) {
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int idx = tid + e * 256;
    int i, kk;
    if (!g.transA) { i = idx >> 5; kk = idx & 31; } else { kk = idx >> 6; i = idx & 63; }
    const int gi = i0 + i, gk = k0 + kk;
    double v = 0.0;
    if (gi < g.m && gk < g.k) {
      const int r = g.transA ? gk : gi, c = g.transA ? gi : gk;
      if (!g.lowerA || c <= r) v = A[(long long)r * g.lda + c];
    }
    ra[e] = v;
    int j, kb;
    if (!g.transB) { kb = idx >> 6; j = idx & 63; } else { j = idx >> 5; kb = idx & 31; }
    const int gj = j0 + j, gkb = k0 + kb;
    double w = 0.0;
    if (gj < g.n && gkb < g.k) {
      const int r = g.transB ? gj : gkb, c = g.transB ? gkb : gj;
      if (!g.lowerB || c <= r) w = B[(long long)r * g.ldb + c];
    }
    rb[e] = w;
  }
}

// D (8x8) += A (8x4, row) * B (4x8, col) on the float64 tensor cores.  Fragments (PTX ISA, mma.m8n8k4 .f64):
//   a: row = lane / 4, col = lane % 4;   b: row (k) = lane % 4, col (n) = lane / 4;   c0, c1: row = lane / 4, cols 2 (lane % 4) + {0, 1}
__device__ __forceinline__ void dmma_884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// 64x64 output tile per CTA, 32-deep k-slices staged through shared memory (register double buffering of the global loads:
// these GEMMs are small and latency-bound), inner product on DMMA: 8 warps as 4 (m) x 2 (n), each 16 x 32 = 2 x 4 m8n8 tiles.
__global__ void __launch_bounds__(256) gemm_f64_kernel(GemmF64 g) {
  __shared__ double As[KC][64 + 2];
  __shared__ double Bs[KC][64 + 2];
  const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
  if (g.lowerC && j0 > i0 + 63) return;   // output tile strictly above the diagonal: not needed
  // blockIdx.z = batch index, or (g.ksplit > 1, single matrix) the k-split: partial products are added atomically into a
  // zeroed C -- these GEMMs are latency-bound chains of 16 k-slices on 64 CTAs; splitting k puts 4x the CTAs to work
  const int bz = g.ksplit > 1 ? 0 : blockIdx.z;
  const double* __restrict__ A = g.A + (long long)bz * g.strideA;
  const double* __restrict__ B = g.B + (long long)bz * g.strideB;
  double* __restrict__ C = g.C + (long long)bz * g.strideC;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp >> 1) * 16, wn = (warp & 1) * 32;       // this warp's 16 x 32 sub-tile
  const int fr = lane >> 2, fc = lane & 3;
  double acc[2][4][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

  // k-range that can be non-zero given triangular operands (block-uniform)
  int kbeg = 0, kend = g.k;
  if (g.lowerA) { if (!g.transA) kend = min(kend, i0 + 64); else kbeg = max(kbeg, i0 & ~(KC - 1)); }
  if (g.lowerB) { if (!g.transB) kbeg = max(kbeg, j0 & ~(KC - 1)); else kend = min(kend, j0 + 64); }
  if (g.ksplit > 1) {                      // this CTA's share of the (trimmed) k-range, in whole slices
    const int slices = (kend - kbeg + KC - 1) / KC, per = (slices + g.ksplit - 1) / g.ksplit;
    const int s0 = (int)blockIdx.z * per;
    kbeg = kbeg + s0 * KC;
    kend = min(kend, kbeg + per * KC);
  }

  double ra[8], rb[8];
  if (kbeg < kend) gemm_f64_fetch(g, A, B, i0, j0, kbeg, tid, ra, rb);
  for (int k0 = kbeg; k0 < kend; k0 += KC) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int idx = tid + e * 256;
      if (!g.transA) As[idx & 31][idx >> 5] = ra[e]; else As[idx >> 6][idx & 63] = ra[e];
      if (!g.transB) Bs[idx >> 6][idx & 63] = rb[e]; else Bs[idx & 31][idx >> 5] = rb[e];
    }
    __syncthreads();
    if (k0 + KC < kend) gemm_f64_fetch(g, A, B, i0, j0, k0 + KC, tid, ra, rb);
#pragma unroll
    for (int kk = 0; kk < KC; kk += 4) {
      double a[2], b[4];
#pragma unroll
      for (int u = 0; u < 2; ++u) a[u] = As[kk + fc][wm + u * 8 + fr];
#pragma unroll
      for (int v = 0; v < 4; ++v) b[v] = Bs[kk + fc][wn + v * 8 + fr];
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) dmma_884(acc[u][v][0], acc[u][v][1], a[u], b[v]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int gi = i0 + wm + u * 8 + fr;
    if (gi >= g.m) continue;
#pragma unroll
    for (int v = 0; v < 4; ++v)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int gj = j0 + wn + v * 8 + 2 * fc + e;
        if (gj >= g.n) continue;
        double* p = C + (long long)gi * g.ldc + gj;
        double r = g.alpha * acc[u][v][e];
        if (g.ksplit > 1) { if (kbeg < kend) atomicAdd(p, r); continue; }
        if (g.beta != 0.0) r += g.beta * (*p);
        *p = r;
      }
  }
}

int gemm_f64(const GemmF64& gin, cudaStream_t st) {
  if (gin.m <= 0 || gin.n <= 0 || gin.batch <= 0) return DCGP_OK;
  GemmF64 g = gin;
  const int tiles = ceil_div(g.n, 64) * ceil_div(g.m, 64);
  g.ksplit = 1;
  if (g.batch == 1 && g.beta == 0.0 && !g.lowerC && g.k >= 256 && tiles <= 96) g.ksplit = g.k >= 512 ? 4 : 2;
  dim3 grid(ceil_div(g.n, 64), ceil_div(g.m, 64), g.ksplit > 1 ? g.ksplit : g.batch);
  if (g.ksplit > 1) cudaMemset2DAsync(g.C, (size_t)g.ldc * sizeof(double), 0, (size_t)g.n * sizeof(double), (size_t)g.m, st);
  gemm_f64_kernel<<<grid, 256, 0, st>>>(g);
  return check_launch("gemm_f64");
}

// ------------------------------------------------------------------------------------------ Kuu
// layers.py:18-21 / kernels.py:135-136: variance * exp(-0.5 * |zi - zj|^2 / l^2) + jitter * I
// hyp (nullable): device {variance, lengthscale} that override the by-value pair (a caller that queues this launch before
// the host knows the step's hyper-parameters: grad.TrainStep)
__global__ void __launch_bounds__(256) rbf_sym_f64_kernel(const double* __restrict__ Z, int M, int L, double variance,
                                                          double inv_ls, double jitter, double* __restrict__ K,
                                                          const double* __restrict__ hyp) {
  __shared__ double Zi[16][17], Zj[16][17];
  if (hyp) { variance = hyp[0]; inv_ls = 1.0 / hyp[1]; }
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i = blockIdx.y * 16 + ty, j = blockIdx.x * 16 + tx;
  double d = 0.0;
  for (int l0 = 0; l0 < L; l0 += 16) {
    const int li = blockIdx.y * 16 + ty, lj = blockIdx.x * 16 + ty;
    Zi[ty][tx] = (li < M && l0 + tx < L) ? Z[(long long)li * L + l0 + tx] * inv_ls : 0.0;
    Zj[ty][tx] = (lj < M && l0 + tx < L) ? Z[(long long)lj * L + l0 + tx] * inv_ls : 0.0;
    __syncthreads();
#pragma unroll
    for (int l = 0; l < 16; ++l) {
      const double t = Zi[ty][l] - Zj[tx][l];
      d = fma(t, t, d);
    }
    __syncthreads();
  }
  if (i < M && j < M) K[(long long)i * M + j] = variance * exp(-0.5 * d) + (i == j ? jitter : 0.0);
}

int rbf_sym_f64(const double* Z, int M, int L, double variance, double lengthscale, double jitter, double* K,
                cudaStream_t st, const double* hyp) {
  dim3 grid(ceil_div(M, 16), ceil_div(M, 16));
  rbf_sym_f64_kernel<<<grid, 256, 0, st>>>(Z, M, L, variance, 1.0 / lengthscale, jitter, K, hyp);
  return check_launch("rbf_sym_f64");
}

// ------------------------------------------------------------------------------------------ Cholesky
// Diagonal block: factor an nb<=64 block held in shared memory `s` (padding rows/cols = identity), write it back to A,
// and form its inverse in `x` -> invD (used for the panel solve and as the seed of the triangular inverse).
// Blocked inside the CTA (16-wide panels: warp-level factor, row-parallel panel solve, rank-16 trailing update), then
// the inverse by recursive doubling on 16 -> 32 -> 64 blocks: a dozen block barriers instead of ~200.
// Must be called by all 256 threads of the CTA; `s` must be fully populated and visible (barrier) on entry.
__device__ void factor_diag_smem(double (*s)[NB + 1], double (*x)[NB + 1], double (*tm)[NB + 1], double* __restrict__ A,
                                 int lda, int j0, int nb, double* __restrict__ invD, int* __restrict__ info) {
  __shared__ double invd[NB];                       // 1 / diag(L): every later division becomes a multiplication
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int e = tid; e < NB * NB; e += 256) x[e >> 6][e & 63] = 0.0;
  __syncthreads();
  for (int k0 = 0; k0 < NB; k0 += 16) {
    if (warp == 0) {
      // 16x16 diagonal block in REGISTERS: lane i (< 16) owns row i; column k is broadcast by shuffles, the pivot's
      // reciprocal square root replaces sqrt + divide.  Fully unrolled so that a[] stays in registers.
      const int i = lane & 15;
      double a[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) a[j] = s[k0 + i][k0 + j];
      int bad = 0;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const double p = __shfl_sync(0xffffffffu, a[k], k);
        if (!(p > 0.0) && !bad) bad = k0 + k + 1;
        const double rs = rsqrt(p);
        a[k] = (i == k) ? p * rs : a[k] * rs;       // rows below the pivot: l_ik = a_ik / l_kk
        if (i == k) invd[k0 + k] = rs;
#pragma unroll
        for (int j = k + 1; j < 16; ++j) {
          const double ljk = __shfl_sync(0xffffffffu, a[k], j);
          a[j] = fma(-a[k], ljk, a[j]);             // meaningful for i >= j
        }
      }
      if (bad && lane == 0) atomicCAS(info, 0, j0 + bad);
      if (lane < 16) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j <= i) s[k0 + i][k0 + j] = a[j];
      }
    }
    __syncthreads();
    // panel below the diagonal block: row i solves x * Lkk^T = a  (forward substitution along the 16 columns)
    if (tid < NB - k0 - 16) {
      const int i = k0 + 16 + tid;
      double r[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) r[j] = s[i][k0 + j];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        double v = r[j];
#pragma unroll
        for (int l = 0; l < j; ++l) v = fma(-r[l], s[k0 + j][k0 + l], v);
        r[j] = v * invd[k0 + j];
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) s[i][k0 + j] = r[j];
    }
    __syncthreads();
    // rank-16 update of the trailing lower triangle
    const int n = NB - k0 - 16;
    for (int e = tid; e < n * n; e += 256) {
      const int i = k0 + 16 + e / n, j = k0 + 16 + e % n;
      if (j <= i) {
        double v = s[i][j];
#pragma unroll
        for (int l = 0; l < 16; ++l) v = fma(-s[i][k0 + l], s[j][k0 + l], v);
        s[i][j] = v;
      }
    }
    __syncthreads();
  }
  for (int e = tid; e < nb * nb; e += 256) {
    const int i = e / nb, j = e % nb;
    A[(long long)(j0 + i) * lda + j0 + j] = (j <= i) ? s[i][j] : 0.0;
  }
  // ---- inverse: 16x16 diagonal blocks by forward substitution (one column per thread) ...
  if (tid < NB) {
    const int b0 = (tid >> 4) << 4, c = tid;
    x[c][c] = invd[c];
    for (int i = c + 1; i < b0 + 16; ++i) {
      double acc = 0.0;
      for (int k = c; k < i; ++k) acc = fma(s[i][k], x[k][c], acc);
      x[i][c] = -acc * invd[i];
    }
  }
  __syncthreads();
  // ... then inv([[A,0],[C,B]]) = [[A^-1,0],[-B^-1 C A^-1, B^-1]] for block sizes 16 and 32
  for (int sz = 16; sz < NB; sz *= 2) {
    const int npairs = NB / (2 * sz);
    for (int e = tid; e < npairs * sz * sz; e += 256) {   // T = C * A^-1
      const int pr = e / (sz * sz), r = (e / sz) % sz, c = e % sz;
      const int a0 = pr * 2 * sz, b0 = a0 + sz;
      double acc = 0.0;
      for (int k = c; k < sz; ++k) acc = fma(s[b0 + r][a0 + k], x[a0 + k][a0 + c], acc);
      tm[b0 + r][a0 + c] = acc;
    }
    __syncthreads();
    for (int e = tid; e < npairs * sz * sz; e += 256) {   // X_ba = -B^-1 * T
      const int pr = e / (sz * sz), r = (e / sz) % sz, c = e % sz;
      const int a0 = pr * 2 * sz, b0 = a0 + sz;
      double acc = 0.0;
      for (int k = 0; k <= r; ++k) acc = fma(x[b0 + r][b0 + k], tm[b0 + k][a0 + c], acc);
      x[b0 + r][a0 + c] = -acc;
    }
    __syncthreads();
  }
  for (int e = tid; e < NB * NB; e += 256) invD[e] = x[e >> 6][e & 63];
}

__global__ void __launch_bounds__(256) potrf_diag_kernel(double* __restrict__ A, int lda, int j0, int nb,
                                                         double* __restrict__ invD, int* __restrict__ info) {
  extern __shared__ double dyn_smem[];
  double (*s)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(dyn_smem);
  double (*x)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(dyn_smem + NB * (NB + 1));
  double (*tm)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(dyn_smem + 2 * NB * (NB + 1));
  for (int e = threadIdx.x; e < NB * NB; e += 256) {
    const int i = e >> 6, j = e & 63;
    double v = (i == j) ? 1.0 : 0.0;
    if (i < nb && j < nb) v = (j <= i) ? A[(long long)(j0 + i) * lda + j0 + j] : 0.0;
    s[i][j] = v;
  }
  __syncthreads();
  factor_diag_smem(s, x, tm, A, lda, j0, nb, invD, info);
}

// One launch per 64-column panel of the right-looking factorisation (after the diagonal block j has been factored and
// inverted): CTA (bi, bk), bi >= bk >= 1 (block offsets past panel j), owns the trailing tile A[j+bi, j+bk] and does
//   X_i = P_i inv(L_jj)^T, X_k = P_k inv(L_jj)^T      (panel solve, recomputed per tile: 2 x 64^3 flops, no extra launch)
//   A[j+bi, j+bk] -= X_i X_k^T                         (rank-64 update)
// the bk == 1 column of CTAs also stores X_i as the final panel of L (transposed, in the upper triangle), and CTA (1, 1)
// goes straight on to factor and invert
// the NEXT diagonal block (look-ahead), so the whole factorisation is nblk launches with no separate diag/solve kernels.
__global__ void __launch_bounds__(256) chol_step_kernel(double* __restrict__ A, int lda, int M, int j0,
                                                        const double* __restrict__ invD, double* __restrict__ invD_next,
                                                        int* __restrict__ info) {
  extern __shared__ double dyn_smem[];
  double (*Pi)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(dyn_smem);
  double (*Pk)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(dyn_smem + NB * (NB + 1));
  double (*D)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(dyn_smem + 2 * NB * (NB + 1));
  // linear tile index -> (bi, bk), bi >= bk >= 1
  int bi = 1, t = blockIdx.x;
  while (t >= bi) { t -= bi; ++bi; }
  const int bk = t + 1;
  const int ri0 = j0 + bi * NB, rk0 = j0 + bk * NB;          // first rows of blocks i and k (rk0 is also the tile's column origin)
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const bool diag = (bi == bk);
  for (int e = tid; e < NB * NB; e += 256) {
    const int i = e >> 6, j = e & 63;
    Pi[i][j] = (ri0 + i < M) ? A[(long long)(ri0 + i) * lda + j0 + j] : 0.0;
    if (!diag) Pk[i][j] = (rk0 + i < M) ? A[(long long)(rk0 + i) * lda + j0 + j] : 0.0;
    D[i][j] = invD[e];
  }
  __syncthreads();
  // panel solve through the inverse of the diagonal block: X[r][c] = sum_k P[r][k] * D[c][k] (D lower triangular, stored
  // with zeros above the diagonal).  4x4 outputs per thread at rows ty + 16u, cols tx + 16v: the 16 lanes that differ in
  // tx read 16 consecutive shared-memory words (no bank conflicts), the two ty values of a warp are broadcasts.
  double xi[4][4], xk[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) { xi[u][v] = 0.0; xk[u][v] = 0.0; }
#pragma unroll 2
  for (int k = 0; k < NB; ++k) {
    double d[4], a[4], b[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) d[v] = D[tx + 16 * v][k];
#pragma unroll
    for (int u = 0; u < 4; ++u) { a[u] = Pi[ty + 16 * u][k]; b[u] = diag ? 0.0 : Pk[ty + 16 * u][k]; }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) { xi[u][v] = fma(a[u], d[v], xi[u][v]); xk[u][v] = fma(b[u], d[v], xk[u][v]); }
  }
  __syncthreads();
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      Pi[ty + 16 * u][tx + 16 * v] = xi[u][v];
      if (!diag) Pk[ty + 16 * u][tx + 16 * v] = xk[u][v];
      // final L panel: parked TRANSPOSED in the (otherwise unused) upper triangle -- other CTAs of this launch still read
      // the raw panel in place; lower_from_upper_kernel moves it home at the end
      if (bk == 1 && ri0 + ty + 16 * u < M) A[(long long)(j0 + tx + 16 * v) * lda + ri0 + ty + 16 * u] = xi[u][v];
    }
  __syncthreads();
  double (*Xk)[NB + 1] = diag ? Pi : Pk;
  // rank-64 update of this tile
  double acc[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int gi = ri0 + ty + 16 * u, gj = rk0 + tx + 16 * v;
      acc[u][v] = (gi < M && gj < M) ? A[(long long)gi * lda + gj] : 0.0;
    }
#pragma unroll 4
  for (int k = 0; k < NB; ++k) {
    double a[4], b[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { a[u] = Pi[ty + 16 * u][k]; b[u] = Xk[tx + 16 * u][k]; }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) acc[u][v] = fma(-a[u], b[v], acc[u][v]);
  }
  if (!(diag && bi == 1)) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int gi = ri0 + ty + 16 * u, gj = rk0 + tx + 16 * v;
        if (gi < M && gj < M) A[(long long)gi * lda + gj] = acc[u][v];
      }
    return;
  }
  // CTA (1, 1): the updated tile is the next diagonal block -> factor + invert it right here (look-ahead)
  __syncthreads();                                   // everyone is done reading Pi / D
  const int jn = j0 + NB, nbn = (M - jn < NB) ? (M - jn) : NB;
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int i = ty + 16 * u, j = tx + 16 * v;
      double val = (i == j) ? 1.0 : 0.0;
      if (i < nbn && j < nbn) val = (j <= i) ? acc[u][v] : 0.0;
      Pi[i][j] = val;
    }
  __syncthreads();
  factor_diag_smem(Pi, Pk, D, A, lda, jn, nbn, invD_next, info);
}

// Off-diagonal blocks of L were parked transposed in the upper triangle (chol_step_kernel): move them home, zero the upper
// triangle.  One thread per (i, j), j < i: it alone touches A[i][j] and A[j][i].
__global__ void lower_from_upper_kernel(double* __restrict__ A, int lda, int M) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (i < M && j < i) {
    if (i / NB != j / NB) A[(long long)i * lda + j] = A[(long long)j * lda + i];
    A[(long long)j * lda + i] = 0.0;
  }
}

size_t potrf_ws_bytes(int M) { return (size_t)ceil_div(M, NB) * NB * NB * sizeof(double); }

constexpr int kDiagSmem = 3 * NB * (NB + 1) * sizeof(double);

// Right-looking blocked Cholesky, ONE launch per 64-column panel (chol_step_kernel): panel solve, rank-64 update of the
// trailing lower triangle (a wide, shallow update that fills the machine, where a left-looking update would be a narrow
// GEMM with a long sequential k loop) and the factor + inverse of the next diagonal block.
int potrf_f64(double* A, int lda, int M, double* invD, int* info, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDiagSmem);
    cudaFuncSetAttribute(chol_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDiagSmem);
    attr_set = true;
  }
  const int nblk = ceil_div(M, NB);
  potrf_diag_kernel<<<1, 256, kDiagSmem, st>>>(A, lda, 0, M < NB ? M : NB, invD, info);
  for (int jb = 0; jb + 1 < nblk; ++jb) {            // panel jb: solve + trailing update + factor of diagonal block jb+1
    const int nt = nblk - 1 - jb;                    // trailing blocks
    chol_step_kernel<<<nt * (nt + 1) / 2, 256, kDiagSmem, st>>>(A, lda, M, jb * NB, invD + (size_t)jb * NB * NB,
                                                                invD + (size_t)(jb + 1) * NB * NB, info);
  }
  lower_from_upper_kernel<<<dim3(ceil_div(M, 256), M), 256, 0, st>>>(A, lda, M);
  return check_launch("potrf_f64", nblk + 1);
}

// ------------------------------------------------------------------------------------------ L^-1
int trtri_pad(int M) {
  int nblk = ceil_div(M, NB), p = 1;
  while (p < nblk) p *= 2;
  return p * NB;
}
size_t trtri_ws_bytes(int M) {
  const size_t Mq = trtri_pad(M);
  return (Mq * Mq + Mq * Mq / 4 + 64) * sizeof(double);
}

__global__ void trtri_init_kernel(const double* __restrict__ L, int lda, int M, int Mq, const double* __restrict__ invD,
                                  double* __restrict__ Lpad, double* __restrict__ Linv) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (i >= Mq || j >= Mq) return;
  double l = (i == j) ? 1.0 : 0.0;
  if (i < M && j < M) l = (j <= i) ? L[(long long)i * lda + j] : 0.0;
  Lpad[(long long)i * Mq + j] = l;
  double v = 0.0;
  const int bi = i / NB, bj = j / NB;
  if (bi == bj) {
    if (bi * NB < M) v = invD[(size_t)bi * NB * NB + (i % NB) * NB + (j % NB)];
    else v = (i == j) ? 1.0 : 0.0;
  }
  Linv[(long long)i * Mq + j] = v;
}

// inv([[A,0],[C,B]]) = [[A^-1,0],[-B^-1 C A^-1, B^-1]], applied level by level (block sizes 64,128,...).
int trtri_f64(const double* L, int lda, int M, const double* invD, double* Linv, void* ws, cudaStream_t st) {
  const int Mq = trtri_pad(M);
  double* Lpad = (double*)ws;
  double* T = Lpad + (size_t)Mq * Mq;
  trtri_init_kernel<<<dim3(ceil_div(Mq, 256), Mq), 256, 0, st>>>(L, lda, M, Mq, invD, Lpad, Linv);
  for (int sz = NB; sz < Mq; sz *= 2) {
    const int npairs = Mq / (2 * sz);
    const long long dstride = 2LL * sz * (Mq + 1);
    GemmF64 g1{};  // T_p = C_p * A_p^-1
    g1.m = g1.n = g1.k = sz;
    g1.A = Lpad + (long long)sz * Mq; g1.lda = Mq; g1.strideA = dstride;
    g1.B = Linv; g1.ldb = Mq; g1.lowerB = 1; g1.strideB = dstride;
    g1.C = T; g1.ldc = sz; g1.strideC = (long long)sz * sz;
    g1.alpha = 1.0; g1.beta = 0.0; g1.batch = npairs;
    int rc = gemm_f64(g1, st);
    if (rc) return rc;
    GemmF64 g2{};  // Linv[b,a] = -B_p^-1 * T_p
    g2.m = g2.n = g2.k = sz;
    g2.A = Linv + (long long)sz * Mq + sz; g2.lda = Mq; g2.lowerA = 1; g2.strideA = dstride;
    g2.B = T; g2.ldb = sz; g2.strideB = (long long)sz * sz;
    g2.C = Linv + (long long)sz * Mq; g2.ldc = Mq; g2.strideC = dstride;
    g2.alpha = -1.0; g2.beta = 0.0; g2.batch = npairs;
    rc = gemm_f64(g2, st);
    if (rc) return rc;
  }
  return check_launch("trtri_f64", 1);
}

// ------------------------------------------------------------------------------------------ reductions
__device__ __forceinline__ double block_sum_1024(double v) {
  __shared__ double sh[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  v = (threadIdx.x < 32) ? sh[threadIdx.x] : 0.0;
  if (threadIdx.x < 32) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  return v;
}

// Multi-CTA: every CTA reduces a slice and adds its partial with one double atomicAdd (out must be zeroed first).
__global__ void __launch_bounds__(1024) sumsq_f64_kernel(const double* __restrict__ x, long long rows, int cols, int ld,
                                                         int lower_period, double* __restrict__ out) {
  double acc = 0.0;
  const long long n = rows * cols;
  for (long long e = blockIdx.x * 1024LL + threadIdx.x; e < n; e += 1024LL * gridDim.x) {
    const long long r = e / cols;
    const int c = (int)(e % cols);
    if (lower_period > 0 && c > (int)(r % lower_period)) continue;
    const double v = x[r * ld + c];
    acc = fma(v, v, acc);
  }
  acc = block_sum_1024(acc);
  if (threadIdx.x == 0) atomicAdd(out, acc);
}

int sumsq_f64(const double* x, long long rows, int cols, int ld, int lower_period, double* out, cudaStream_t st) {
  cudaMemsetAsync(out, 0, sizeof(double), st);
  const long long n = rows * cols;
  int blocks = (int)((n + 8191) / 8192);
  if (blocks < 1) blocks = 1;
  if (blocks > 148) blocks = 148;
  sumsq_f64_kernel<<<blocks, 1024, 0, st>>>(x, rows, cols, ld, lower_period, out);
  return check_launch("sumsq_f64");
}

__global__ void __launch_bounds__(1024) logdiag2_f64_kernel(const double* __restrict__ A, int lda, int M, int batch,
                                                            long long stride, double* __restrict__ out) {
  double acc = 0.0;
  for (int e = threadIdx.x; e < batch * M; e += 1024) {
    const int b = e / M, i = e % M;
    const double d = A[b * stride + (long long)i * lda + i];
    acc += log(d * d);
  }
  acc = block_sum_1024(acc);
  if (threadIdx.x == 0) *out = acc;
}

int logdiag2_f64(const double* A, int lda, int M, int batch, long long stride, double* out, cudaStream_t st) {
  logdiag2_f64_kernel<<<1, 1024, 0, st>>>(A, lda, M, batch, stride, out);
  return check_launch("logdiag2_f64");
}

}  // namespace dcgp
