// M-only chain rule of one layer, native: what tf.gradients derives for the minibatch-independent part of
// conditionals.py:29-58 + the KL (layers.py:137-147, DS/layers.py:231-256), in the order of the forward pass:
//     Lm = chol(Kuu), Li = Lm^-1, a = Li k;   C_r = Li L_r (L_r when whitened), S_r = C_r C_r^T, alpha = Li q_mu (q_mu)
//     mean_r = alpha_r^T a,   var_r = knn - |a|^2 + a^T S_r a.
// Inputs from dcgp_layer_backward: dS_r = sum_t s_r a a^T, dalpha = sum_t a g_mean^T, and the direct paths gZ, gscal.
//     W  = alpha dalpha^T + sum_r 2 (S_r - I) dS_r  [+ sum_r 2 dS_r S_r + dalpha alpha^T  when C_r, alpha move with Li]
//     dLm = tril(-Li^T W),   dKuu = 1/2 Li^T (P + P^T) Li,  P = Phi(Lm^T dLm)     (Cholesky backward; Phi = tril, diagonal / 2)
//     d/dL_r = tril(Li^T 2 dS_r C_r)  (tril(2 dS_r C_r) whitened),   d/dq_mu = Li^T dalpha  (dalpha whitened)
//     then the RBF chain to Z, variance, lengthscale, plus the KL gradient.
// Single-matrix products: float64 on the DMMA GEMM of dcgp_f64.cu; the R-batched M^3 products: the split-fp16 tcgen05 GEMM
// (their inputs -- dS from the tensor-core GEMMs, C_r / S_r from the forward -- carry 22 bits anyway).
// The kernel hyper-parameters are read from DEVICE memory (`hyp`), so that the whole call can sit in a CUDA graph that is
// replayed while the host's copy of them is still one optimiser step behind (grad.TrainStep).
#include <string.h>

#include "dcgp_kernels.cuh"
#include "dcgp_tc.cuh"

namespace dcgp {

struct ChainWork {
  int M, Mp, R, L;
  // persistent between the static (parameter-only) and the dynamic part
  double *Kpinv, *a, *dKL, *sc;      // [Mp,Mp], [M,R], [Mp,Mp], scalars: {gvar_p, gls_p, sH, sHD}
  float *Cb32, *LqT32, *Kp32;        // [R,Mp,Mp] Kp^-1 L_r, [R,Mp,Mp] L_r^T, [Mp,Mp]
  // dynamic
  double *W, *X, *A2, *Ps, *T1, *GU, *Hs, *rs, *HZ, *tmpMR;
  float *gS32, *Tm32, *GCt32, *res32, *LiT32;
  void* bg;                          // workspace of the batched tensor-core GEMM
  size_t bg_bytes;
  size_t bytes;
};

struct Carver2 {
  char* base; size_t off = 0;
  explicit Carver2(void* p) : base((char*)p) {}
  template <typename T> T* take(size_t n) {
    off = align_up(off, 256);
    T* r = (T*)(base ? base + off : nullptr);
    off += n * sizeof(T);
    return r;
  }
};

static ChainWork carve_chain(int M, int R, int L, void* ws) {
  ChainWork c;
  c.M = M; c.Mp = (int)align_up(M, 64); c.R = R; c.L = L;
  const size_t mm = (size_t)c.Mp * c.Mp, rmm = (size_t)R * mm;
  Carver2 k(ws);
  c.Kpinv = k.take<double>(mm); c.a = k.take<double>((size_t)M * R); c.dKL = k.take<double>(mm); c.sc = k.take<double>(8);
  c.Cb32 = k.take<float>(rmm); c.LqT32 = k.take<float>(rmm); c.Kp32 = k.take<float>(mm);
  c.W = k.take<double>(mm); c.X = k.take<double>(mm); c.A2 = k.take<double>(mm); c.Ps = k.take<double>(mm);
  c.T1 = k.take<double>(mm); c.GU = k.take<double>(mm); c.Hs = k.take<double>(mm); c.rs = k.take<double>(c.Mp);
  c.HZ = k.take<double>((size_t)M * L); c.tmpMR = k.take<double>((size_t)M * R);
  c.gS32 = k.take<float>(rmm); c.Tm32 = k.take<float>(rmm); c.GCt32 = k.take<float>(rmm); c.res32 = k.take<float>(rmm);
  c.LiT32 = k.take<float>(mm);
  c.bg_bytes = tc_bgemm_workspace_bytes(R, c.Mp, c.Mp, c.Mp);
  c.bg = k.take<char>(c.bg_bytes);
  c.bytes = align_up(k.off, 256);
  return c;
}
size_t chain_rule_workspace_bytes(int M, int R, int L) { return carve_chain(M, R, L, nullptr).bytes; }

// ------------------------------------------------------------------------------------------ small kernels
#define GRID_STRIDE(e, total) for (long long e = blockIdx.x * 256LL + threadIdx.x; e < (total); e += 256LL * gridDim.x)
// index splits in 32-bit arithmetic (every element count here is below 2^31: chain_rule checks R * Mp * Mp): a 64-bit
// division by a run-time divisor costs ~80 instructions, more than the rest of these element-wise kernels
__device__ __forceinline__ int div32(long long e, int d, int& rem) {
  const unsigned u = (unsigned)e, q = u / (unsigned)d;
  rem = (int)(u - q * (unsigned)d);
  return (int)q;
}
static int blocks_for(long long n) { long long b = (n + 255) / 256; return (int)(b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b)); }

// dst[i, j] = (float) src[i, j] (or its transpose), [M, M] inside [Mp, Mp] with zero padding; `lower`: keep i >= j of the source
__global__ void __launch_bounds__(256) to_f32_kernel(const double* __restrict__ src, int ld, int M, int Mp, int transpose, int lower,
                                                     float* __restrict__ dst) {
  GRID_STRIDE(e, (long long)Mp * Mp) {
    int j;
    const int i = div32(e, Mp, j);
    float v = 0.f;
    if (i < M && j < M) {
      const int r = transpose ? j : i, c = transpose ? i : j;
      if (!lower || c <= r) v = (float)src[(long long)r * ld + c];
    }
    dst[e] = v;
  }
}
// LqT[(r*Mp + j)*Mp + i] = L_r[i, j] (i >= j)
__global__ void __launch_bounds__(256) lqt_kernel(const double* __restrict__ q_sqrt, int M, int Mp, int R, float* __restrict__ out) {
  GRID_STRIDE(e, (long long)R * Mp * Mp) {
    int i, j;
    const int q = div32(e, Mp, i), r = div32(q, Mp, j);
    out[e] = (i < M && j < M && i >= j) ? (float)q_sqrt[((long long)r * M + i) * M + j] : 0.f;
  }
}
// dKL/dKp = 1/2 (-a a^T - sum_r Cb_r Cb_r^T + R Kp^-1)        (CC = the R products Cb_r Cb_r^T, float32 [R, Mp, Mp])
__global__ void __launch_bounds__(256) dkl_kernel(const double* __restrict__ a, const float* __restrict__ CC, const double* __restrict__ Kpinv,
                                                  int ldk, int M, int Mp, int R, double* __restrict__ dKL) {
  GRID_STRIDE(e, (long long)Mp * Mp) {
    int j;
    const int i = div32(e, Mp, j);
    double v = 0.0;
    if (i < M && j < M) {
      double aa = 0.0, cc = 0.0;
      for (int r = 0; r < R; ++r) {
        aa += a[(long long)i * R + r] * a[(long long)j * R + r];
        cc += (double)CC[((long long)r * Mp + i) * Mp + j];
      }
      v = 0.5 * (-aa - cc + (double)R * Kpinv[(long long)i * ldk + j]);
    }
    dKL[e] = v;
  }
}
// gS32[r] = (float) dS_r (blocks 1..R of gQB, ld Mp), zero padded
__global__ void __launch_bounds__(256) gs32_kernel(const double* __restrict__ gQB, int M, int Mp, int R, float* __restrict__ gS32) {
  GRID_STRIDE(e, (long long)R * Mp * Mp) {
    int i, j;
    const int q = div32(e, Mp, j);
    div32(q, Mp, i);
    gS32[e] = (i < M && j < M) ? (float)gQB[(long long)Mp * Mp + e] : 0.f;
  }
}
// W = alpha dalpha^T + 2 (sum_r Tm_r - sum_r dS_r)  [+ 2 (sum_r Tm_r)^T + dalpha alpha^T when not whitened],  Tm_r = S_r dS_r
__global__ void __launch_bounds__(256) w_kernel(const double* __restrict__ alpha, const double* __restrict__ gQB, const float* __restrict__ Tm,
                                                int M, int Mp, int R, int white, double* __restrict__ W) {
  const double* galpha = gQB + (long long)(R + 1) * Mp * Mp;      // row r: dalpha[:, r]
  GRID_STRIDE(e, (long long)Mp * Mp) {
    int j;
    const int i = div32(e, Mp, j);
    double v = 0.0;
    if (i < M && j < M) {
      double ts = 0.0, tst = 0.0, gs = 0.0, ag = 0.0, ga = 0.0;
      for (int r = 0; r < R; ++r) {
        ts += (double)Tm[((long long)r * Mp + i) * Mp + j];
        tst += (double)Tm[((long long)r * Mp + j) * Mp + i];
        gs += gQB[((long long)(r + 1) * Mp + i) * Mp + j];
        ag += alpha[(long long)i * R + r] * galpha[(long long)r * Mp + j];
        ga += galpha[(long long)r * Mp + i] * alpha[(long long)j * R + r];
      }
      v = ag + 2.0 * (ts - gs);
      if (!white) v += 2.0 * tst + ga;
    }
    W[e] = v;
  }
}
// g_qsqrt[r, i, j] (contiguous [R, M, M]) = tril( 2 src ) + KL part, where src = res32[r, i, j] (non-whitened: Li^T dS_r C_r) or
// GCt32[r, j, i] (whitened: dS_r C_r), and the KL part is -klw (Cb_r - diag(1 / diag L_r))  (whitened: Cb_r = L_r)
__global__ void __launch_bounds__(256) gqsqrt_kernel(const float* __restrict__ res32, const float* __restrict__ GCt32,
                                                     const float* __restrict__ Cb32, const double* __restrict__ q_sqrt, int M, int Mp,
                                                     int R, int white, double klw, double* __restrict__ out) {
  GRID_STRIDE(e, (long long)R * M * M) {
    int i, j;
    const int q = div32(e, M, j), r = div32(q, M, i);
    double v = 0.0;
    if (i >= j) {
      const double lij = q_sqrt[e];
      const double g = white ? (double)GCt32[((long long)r * Mp + j) * Mp + i] : (double)res32[((long long)r * Mp + i) * Mp + j];
      const double cb = white ? lij : (double)Cb32[((long long)r * Mp + i) * Mp + j];
      v = 2.0 * g - klw * (cb - (i == j ? 1.0 / lij : 0.0));
    }
    out[e] = v;
  }
}
// out = x + s * y   ([n] doubles)
__global__ void __launch_bounds__(256) axpy_kernel(const double* __restrict__ x, double s, const double* __restrict__ y, long long n,
                                                   double* __restrict__ out) {
  GRID_STRIDE(e, n) out[e] = (x ? x[e] : 0.0) + s * y[e];
}
// g_qmu (whitened) = dalpha - klw q_mu, dalpha stored as rows of gQB
__global__ void __launch_bounds__(256) gqmu_white_kernel(const double* __restrict__ gQB, const double* __restrict__ q_mu, int M, int Mp,
                                                         int R, double klw, double* __restrict__ out) {
  const double* galpha = gQB + (long long)(R + 1) * Mp * Mp;
  GRID_STRIDE(e, (long long)M * R) {
    int r;
    const int i = div32(e, R, r);
    out[e] = galpha[(long long)r * Mp + i] - klw * q_mu[e];
  }
}
// Ps = P + P^T with P = Phi(A2) = tril(A2), diagonal halved
__global__ void __launch_bounds__(256) psym_kernel(const double* __restrict__ A2, int M, int Mp, double* __restrict__ Ps) {
  GRID_STRIDE(e, (long long)Mp * Mp) {
    int j;
    const int i = div32(e, Mp, j);
    Ps[e] = (i < M && j < M) ? (i >= j ? A2[e] : A2[(long long)j * Mp + i]) : 0.0;
  }
}
// RBF chain, part A: Hs[i,j] = (G[i,j] + G[j,i]) K[i,j] with K = var exp(-d/2) recomputed from Z, G = GU (+ gscale * Gadd),
// rs[i] = sum_j Hs[i,j], sc[2] += sum G K, sc[3] += sum G K d       (one 16 x 16 tile per CTA, like rbf_sym_f64_kernel)
__global__ void __launch_bounds__(256) rbf_chain_a_kernel(const double* __restrict__ Z, int M, int L, const double* __restrict__ hyp,
                                                          const double* __restrict__ GU, int ldg, const double* __restrict__ Gadd,
                                                          int lda, double gscale, double* __restrict__ Hs, int ldh,
                                                          double* __restrict__ rs, double* __restrict__ sc) {
  __shared__ double Zi[16][17], Zj[16][17];
  __shared__ double red[8][2];
  const double var = hyp[0], inv_ls = 1.0 / hyp[1];
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
  double h = 0.0, hd = 0.0, hs = 0.0;
  if (i < M && j < M) {
    const double k = var * exp(-0.5 * d);
    double gij = GU[(long long)i * ldg + j], gji = GU[(long long)j * ldg + i];
    if (Gadd) { gij += gscale * Gadd[(long long)i * lda + j]; gji += gscale * Gadd[(long long)j * lda + i]; }
    h = gij * k;
    hd = h * d;
    hs = (gij + gji) * k;
    if (Hs) Hs[(long long)i * ldh + j] = hs;
  }
  // row sums of Hs: reduce over the 16 lanes of a row (tx), then one atomic per row and CTA
  double r = hs;
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  if (rs && tx == 0 && i < M) atomicAdd(&rs[i], r);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { h += __shfl_xor_sync(0xffffffffu, h, o); hd += __shfl_xor_sync(0xffffffffu, hd, o); }
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = h; red[threadIdx.x >> 5][1] = hd; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < 8; ++w) { a += red[w][0]; b += red[w][1]; }
    atomicAdd(&sc[2], a);
    atomicAdd(&sc[3], b);
  }
}
// prior chain of the ConvLayer KL (Z_prior is a constant): sc[0] = gvar_p, sc[1] = gls_p from sc[2], sc[3]; then clears sc[2..3]
__global__ void rbf_prior_scalars_kernel(const double* __restrict__ hyp, double* __restrict__ sc) {
  sc[0] = sc[2] / hyp[0];
  sc[1] = sc[3] / hyp[1];
  sc[2] = sc[3] = 0.0;
}
// RBF chain, part B: gZ = -(rs Z - Hs Z) / ls^2 + gZ_direct;  ghyp = {sH / var + gvar_p + gscal[0], sHD / ls + gls_p + gscal[1]}
__global__ void __launch_bounds__(256) rbf_chain_b_kernel(const double* __restrict__ Z, int M, int L, const double* __restrict__ hyp,
                                                          const double* __restrict__ rs, const double* __restrict__ HZ,
                                                          const double* __restrict__ gZ_direct, const double* __restrict__ gscal,
                                                          const double* __restrict__ sc, double* __restrict__ gZ, double* __restrict__ ghyp) {
  const double var = hyp[0], ls = hyp[1];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    ghyp[0] = sc[2] / var + sc[0] + gscal[0];
    ghyp[1] = sc[3] / ls + sc[1] + gscal[1];
  }
  GRID_STRIDE(e, (long long)M * L) {
    int l_unused;
    const int i = div32(e, L, l_unused);
    gZ[e] = -(rs[i] * Z[e] - HZ[e]) / (ls * ls) + gZ_direct[e];
  }
}
__global__ void set_hyp_kernel(double v, double l, double* __restrict__ hyp) { hyp[0] = v; hyp[1] = l; }

static GemmF64 gemm(int m, int n, int k, const double* A, int lda, int tA, int lowA, const double* B, int ldb, int tB, int lowB,
                    double* C, int ldc, double alpha) {
  GemmF64 g{};
  g.m = m; g.n = n; g.k = k;
  g.A = A; g.lda = lda; g.transA = tA; g.lowerA = lowA;
  g.B = B; g.ldb = ldb; g.transB = tB; g.lowerB = lowB;
  g.C = C; g.ldc = ldc; g.alpha = alpha; g.beta = 0.0; g.batch = 1;
  return g;
}

#define CH_TRY(expr) do { int _rc = (expr); if (_rc != DCGP_OK) return _rc; } while (0)

int chain_rule(const dcgp_layer_desc* d, const ChainInputs& in, const double* Z, const double* Z_prior, const double* q_mu,
               const double* q_sqrt, const double* hyp_dev, const double* gQB, const double* gZ_direct, const double* gscal,
               double kl_weight, int parts, double* gZ, double* ghyp, double* g_qmu, double* g_qsqrt, void* ws, cudaStream_t st) {
  const int M = d->M, R = d->R, L = d->f * d->f * d->C, white = d->white;
  const bool conv = d->kind == DCGP_LAYER_CONV;
  ChainWork c = carve_chain(M, R, L, ws);
  const int Mp = c.Mp;
  const long long mm = (long long)Mp * Mp;
  if ((long long)(R + 2) * mm >= (1LL << 31) || (long long)M * L >= (1LL << 31)) { set_error("chain_rule: R * M^2 too large"); return DCGP_ERR_ARG; }
  const double* hyp = hyp_dev;
  if (!hyp) {                       // the descriptor's values (host-current): stage them in the workspace
    set_hyp_kernel<<<1, 1, 0, st>>>(d->variance, d->lengthscale, c.sc + 6);
    CH_TRY(check_launch("set_hyp"));
    hyp = c.sc + 6;
  }
  if (parts & 1) {
    // ---- static part: the KL gradient (depends on the parameters and on this step's prepare only)
    cudaMemsetAsync(c.sc, 0, 4 * sizeof(double), st);
    if (!white) {
      const double* Kpinv = in.Kinv;
      int ldk = M;
      if (conv && Z_prior) {        // prior = Kuu at the initial Z (layers.py:149-150): Kp^-1 = Lp^-T Lp^-1
        CH_TRY(gemm_f64(gemm(M, M, M, in.Lpinv, in.ldi, 1, 1, in.Lpinv, in.ldi, 0, 1, c.Kpinv, Mp, 1.0), st));
        Kpinv = c.Kpinv; ldk = Mp;
      }
      CH_TRY(gemm_f64(gemm(M, R, M, Kpinv, ldk, 0, 0, q_mu, R, 0, 0, c.a, R, 1.0), st));              // a = Kp^-1 q_mu
      to_f32_kernel<<<blocks_for(mm), 256, 0, st>>>(Kpinv, ldk, M, Mp, 0, 0, c.Kp32);
      lqt_kernel<<<blocks_for(R * mm), 256, 0, st>>>(q_sqrt, M, Mp, R, c.LqT32);
      CH_TRY(check_launch("chain_static_pack", 2));
      // Cb_r = Kp^-1 L_r = Kp32 . (L_r^T)^T
      CH_TRY(tc_bgemm_nt_ld(c.Kp32, Mp, 0, c.LqT32, Mp, mm, c.Cb32, Mp, mm, R, Mp, Mp, Mp, c.bg, st));
      CH_TRY(tc_bgemm_nt_ld(c.Cb32, Mp, mm, c.Cb32, Mp, mm, c.res32, Mp, mm, R, Mp, Mp, Mp, c.bg, st));   // Cb_r Cb_r^T
      dkl_kernel<<<blocks_for(mm), 256, 0, st>>>(c.a, c.res32, Kpinv, ldk, M, Mp, R, c.dKL);
      CH_TRY(check_launch("dkl"));
      if (conv) {                   // d/dvariance, d/dlengthscale through the constant-Z prior: sum (-klw dKL) . dKp/dtheta
        dim3 grid(ceil_div(M, 16), ceil_div(M, 16));
        rbf_chain_a_kernel<<<grid, 256, 0, st>>>(Z_prior ? Z_prior : Z, M, L, hyp, c.dKL, Mp, nullptr, 0, 0.0, nullptr, 0, nullptr, c.sc);
        // (sums of dKL . K; the factor -klw is applied below)
        axpy_kernel<<<1, 32, 0, st>>>(nullptr, -kl_weight, c.sc + 2, 2, c.sc + 2);
        rbf_prior_scalars_kernel<<<1, 1, 0, st>>>(hyp, c.sc);
        CH_TRY(check_launch("rbf_prior", 3));
      }
    }
  }
  if (!(parts & 2)) return DCGP_OK;
  // ---- dynamic part
  const double* alpha = white ? q_mu : in.alpha;
  gs32_kernel<<<blocks_for(R * mm), 256, 0, st>>>(gQB, M, Mp, R, c.gS32);
  CH_TRY(check_launch("gs32"));
  CH_TRY(tc_bgemm_nt_ld(in.S32, Mp, mm, c.gS32, Mp, mm, c.Tm32, Mp, mm, R, Mp, Mp, Mp, c.bg, st));        // Tm_r = S_r dS_r
  w_kernel<<<blocks_for(mm), 256, 0, st>>>(alpha, gQB, c.Tm32, M, Mp, R, white, c.W);
  CH_TRY(check_launch("w"));
  CH_TRY(tc_bgemm_nt_ld(in.Ct32, Mp, mm, c.gS32, Mp, mm, c.GCt32, Mp, mm, R, Mp, Mp, Mp, c.bg, st));      // C_r^T dS_r = (dS_r C_r)^T
  if (!white) {
    to_f32_kernel<<<blocks_for(mm), 256, 0, st>>>(in.Li, in.ldi, M, Mp, 1, 1, c.LiT32);                  // Li^T
    CH_TRY(check_launch("lit32"));
    CH_TRY(tc_bgemm_nt_ld(c.LiT32, Mp, 0, c.GCt32, Mp, mm, c.res32, Mp, mm, R, Mp, Mp, Mp, c.bg, st));    // Li^T (dS_r C_r)
  }
  gqsqrt_kernel<<<blocks_for((long long)R * M * M), 256, 0, st>>>(c.res32, c.GCt32, c.Cb32, q_sqrt, M, Mp, R, white, kl_weight, g_qsqrt);
  CH_TRY(check_launch("gqsqrt"));
  if (white) {
    gqmu_white_kernel<<<blocks_for((long long)M * R), 256, 0, st>>>(gQB, q_mu, M, Mp, R, kl_weight, g_qmu);
    CH_TRY(check_launch("gqmu_white"));
  } else {                          // g_qmu = Li^T dalpha - klw a
    const double* galphaT = gQB + (long long)(R + 1) * mm;    // [R, Mp]: row r = dalpha[:, r]
    CH_TRY(gemm_f64(gemm(M, R, M, in.Li, in.ldi, 1, 1, galphaT, Mp, 1, 0, c.tmpMR, R, 1.0), st));
    axpy_kernel<<<blocks_for((long long)M * R), 256, 0, st>>>(c.tmpMR, -kl_weight, c.a, (long long)M * R, g_qmu);
    CH_TRY(check_launch("gqmu"));
  }
  // Cholesky backward: X = -Li^T W, A2 = Lm^T tril(X), Ps = Phi(A2) + Phi(A2)^T, dKuu = 1/2 Li^T Ps Li
  CH_TRY(gemm_f64(gemm(M, M, M, in.Li, in.ldi, 1, 1, c.W, Mp, 0, 0, c.X, Mp, -1.0), st));
  CH_TRY(gemm_f64(gemm(M, M, M, in.Lm, M, 1, 1, c.X, Mp, 0, 1, c.A2, Mp, 1.0), st));
  psym_kernel<<<blocks_for(mm), 256, 0, st>>>(c.A2, M, Mp, c.Ps);
  CH_TRY(check_launch("psym"));
  CH_TRY(gemm_f64(gemm(M, M, M, c.Ps, Mp, 0, 0, in.Li, in.ldi, 0, 1, c.T1, Mp, 1.0), st));
  CH_TRY(gemm_f64(gemm(M, M, M, in.Li, in.ldi, 1, 1, c.T1, Mp, 0, 0, c.GU, Mp, 0.5), st));
  // RBF chain to Z, variance, lengthscale (SVGP_Layer, not whitened: the KL's dKp joins dKuu, the prior is the current Ku)
  cudaMemsetAsync(c.rs, 0, (size_t)Mp * sizeof(double), st);
  {
    dim3 grid(ceil_div(M, 16), ceil_div(M, 16));
    const bool add_kl = !white && !conv;
    rbf_chain_a_kernel<<<grid, 256, 0, st>>>(Z, M, L, hyp, c.GU, Mp, add_kl ? c.dKL : nullptr, Mp, -kl_weight, c.Hs, Mp, c.rs, c.sc);
    CH_TRY(check_launch("rbf_chain_a"));
  }
  CH_TRY(gemm_f64(gemm(M, L, M, c.Hs, Mp, 0, 0, Z, L, 0, 0, c.HZ, L, 1.0), st));
  rbf_chain_b_kernel<<<blocks_for((long long)M * L), 256, 0, st>>>(Z, M, L, hyp, c.rs, c.HZ, gZ_direct, gscal, c.sc, gZ, ghyp);
  CH_TRY(check_launch("rbf_chain_b"));
  // sc[2..3] are consumed: clear them for the next dynamic call that follows a static call (which leaves them at 0)
  cudaMemsetAsync(c.sc + 2, 0, 2 * sizeof(double), st);
  return DCGP_OK;
}

}  // namespace dcgp
