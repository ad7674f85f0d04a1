// float32 CUDA-core kernels for the minibatch-sized ("T-sized") work, plus the small glue kernels.
// The fp32 Kuf / conditional-GEMM kernels here are the validation path (DCGP_ALGO_SIMT); the product path
// replaces them with the tcgen05 kernels of dcgp_tc.cu.  Glue kernels (patch gather, patch-mean, Kdiag,
// finalize/reparameterise, RobustMax expectations, packing) are shared by both paths.
#include <cuda_fp16.h>
#include <math.h>

#include <mutex>

#include "dcgp_kernels.cuh"

namespace dcgp {

// ------------------------------------------------------------------------------------------ a2: patches
// views.py:40-54.  layout 0: [P,N,L], layout 1: [N,P,L].  One thread per output element.
__global__ void patches_kernel(const float* __restrict__ X, View v, int N, int layout, float* __restrict__ out) {
  const long long total = (long long)N * v.P * v.L;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int l = (int)(e % v.L);
    long long r = e / v.L;
    int n, p;
    if (layout == 0) { n = (int)(r % N); p = (int)(r / N); } else { p = (int)(r % v.P); n = (int)(r / v.P); }
    out[e] = X[(long long)n * v.HWC + v.patch_base(p) + v.elem_off(l)];
  }
}

int launch_patches(const float* X, const View& v, int N, int layout, float* out, cudaStream_t st) {
  const long long total = (long long)N * v.P * v.L;
  const int blocks = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
  patches_kernel<<<blocks > 0 ? blocks : 1, 256, 0, st>>>(X, v, N, layout, out);
  return check_launch("patches");
}

// ------------------------------------------------------------------------------------------ a4: Kuf (fp32)
// Fused im2col + squared distance + RBF (layers.py:23-32 with GPflow RBF.K).  64 patches x 64 inducing
// points per CTA; the patch tile is gathered straight from the NHWC image (never materialised).
// d = sum_l (x_l/ls - z_l/ls)^2 is formed as differences (no |x|^2+|z|^2-2xz cancellation in fp32).
// LAYOUT 2 writes the split-fp16 planes (hi, lo of k * kscal[0]) the tensor-core conditional GEMM consumes, rows
// t in [T, Tpad) zero-filled.
template <int LAYOUT>
__global__ void __launch_bounds__(256) kuf_simt_kernel(const float* __restrict__ X, View v, int T, int N,
                                                       const float* __restrict__ zs, int M, float variance,
                                                       float inv_ls, int ldo, float* __restrict__ out,
                                                       const float* __restrict__ kscal = nullptr,
                                                       __half* __restrict__ oh = nullptr, __half* __restrict__ ol = nullptr,
                                                       long long Tpad = 0) {
  constexpr int LC = 32;
  __shared__ float Xs[LC][65];
  __shared__ float Zs[LC][65];
  __shared__ int base_s[64];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int t0 = blockIdx.x * 64, m0 = blockIdx.y * 64;
  if (tid < 64) {
    const int t = t0 + tid;
    int b = -1;
    if (t < T) {
      const int n = t / v.P, p = t - n * v.P;
      b = n * v.HWC + v.patch_base(p);  // fits int: rows*HWC < 2^31 checked on the host
    }
    base_s[tid] = b;
  }
  __syncthreads();
  float d[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) d[a][b] = 0.f;

  for (int l0 = 0; l0 < v.L; l0 += LC) {
#pragma unroll
    for (int e = 0; e < (64 * LC) / 256; ++e) {
      const int idx = tid + e * 256;
      const int l = idx & (LC - 1), r = idx >> 5;
      const int gl = l0 + l;
      float xv = 0.f, zv = 0.f;
      if (gl < v.L) {
        const int b = base_s[r];
        if (b >= 0) xv = X[(long long)b + v.elem_off(gl)] * inv_ls;
        if (m0 + r < M) zv = zs[(long long)(m0 + r) * v.L + gl];
      }
      Xs[l][r] = xv;
      Zs[l][r] = zv;
    }
    __syncthreads();
#pragma unroll 8
    for (int l = 0; l < LC; ++l) {
      float a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { a[u] = Xs[l][ty * 4 + u]; b[u] = Zs[l][tx * 4 + u]; }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const float t = a[u] - b[w];
          d[u][w] = fmaf(t, t, d[u][w]);
        }
    }
    __syncthreads();
  }
  if (LAYOUT == 2) {
    const float ks = kscal[0];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long t = t0 + ty * 4 + u;
      if (t >= Tpad) continue;
      __half hi[4], lo[4];
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const int m = m0 + tx * 4 + w;
        const float k = (t < T && m < M) ? ks * variance * expf(-0.5f * d[u][w]) : 0.f;
        hi[w] = __float2half_rn(k);
        lo[w] = __float2half_rn(k - __half2float(hi[w]));
      }
      const int m = m0 + tx * 4;
      if (m < ldo) {   // ldo is a multiple of 64: the 4 columns are all inside; 8-byte vector stores
        *reinterpret_cast<uint2*>(oh + t * ldo + m) = *reinterpret_cast<uint2*>(hi);
        *reinterpret_cast<uint2*>(ol + t * ldo + m) = *reinterpret_cast<uint2*>(lo);
      }
    }
    return;
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int t = t0 + ty * 4 + u;
    if (t >= T) continue;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int m = m0 + tx * 4 + w;
      const float k = variance * expf(-0.5f * d[u][w]);
      if (LAYOUT == 1) {
        if (m < ldo) out[(long long)t * ldo + m] = (m < M) ? k : 0.f;
      } else {
        if (m < M) {
          const int n = t / v.P, p = t - n * v.P;
          out[((long long)p * M + m) * N + n] = k;
        }
      }
    }
  }
}

int launch_kuf_simt(const float* X, const View& v, int n_rows, const float* zs, int M, float variance, float inv_ls,
                    int layout, int ldo, float* out, cudaStream_t st) {
  const int T = n_rows * v.P;
  const int mcols = (layout == 1) ? ldo : M;
  dim3 grid(ceil_div(T, 64), ceil_div(mcols, 64));
  if (layout == 1)
    kuf_simt_kernel<1><<<grid, 256, 0, st>>>(X, v, T, n_rows, zs, M, variance, inv_ls, ldo, out);
  else
    kuf_simt_kernel<0><<<grid, 256, 0, st>>>(X, v, T, n_rows, zs, M, variance, inv_ls, ldo, out);
  return check_launch("kuf_simt");
}

int launch_kuf_simt_planes(const float* X, const View& v, int n_rows, const float* zs, int M, float variance, float inv_ls,
                           int ldo, const float* kscal, void* Kh, void* Kl, long long Tpad, cudaStream_t st) {
  const int T = n_rows * v.P;
  dim3 grid((unsigned)((Tpad + 63) / 64), ceil_div(ldo, 64));
  kuf_simt_kernel<2><<<grid, 256, 0, st>>>(X, v, T, n_rows, zs, M, variance, inv_ls, ldo, nullptr, kscal, (__half*)Kh,
                                          (__half*)Kl, Tpad);
  return check_launch("kuf_simt_planes");
}

// [P,M,N] (reference Kuf layout) -> [N*P, ldo] rows t = n*P + p (conditional GEMM operand layout)
__global__ void pmn_to_tm_kernel(const float* __restrict__ Kmn, int P, int M, int N, int ldo, float* __restrict__ out) {
  const long long total = (long long)N * P * ldo;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(e % ldo);
    const long long t = e / ldo;
    const int n = (int)(t / P), p = (int)(t % P);
    out[e] = (m < M) ? Kmn[((long long)p * M + m) * N + n] : 0.f;
  }
}
int launch_pmn_to_tm(const float* Kmn, int P, int M, int N, int ldo, float* out, cudaStream_t st) {
  pmn_to_tm_kernel<<<148 * 8, 256, 0, st>>>(Kmn, P, M, N, ldo, out);
  return check_launch("pmn_to_tm");
}

// ------------------------------------------------------------------------------------------ a5: conditional GEMM (fp32)
// G[t, j] = sum_m Kt[t, m] * W[j, m];  acc[t, blk] = sum_{j in blk} G[t, j]^2  (blk = 0..R), and for the extra
// "mean" block the products themselves: mean[t, r] = sum_m Kt[t, m] * Wmean[r, m].
// Replaces conditionals.py:31-33,40,44-47,50-51,55-58,65 in the single-solve form of SURVEY App. A.4.
__global__ void __launch_bounds__(256) cond_simt_kernel(const float* __restrict__ Kt, int T, int ld, int Mp,
                                                        const float* __restrict__ W, const float* __restrict__ Wmean,
                                                        int R, float* __restrict__ acc_out, float* __restrict__ mean_out) {
  __shared__ float As[16][128 + 4];
  __shared__ float Bs[16][64 + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int t0 = blockIdx.x * 128;
  const int blk = blockIdx.y;
  const bool is_mean = (blk == R + 1);
  const float* __restrict__ Wb = is_mean ? Wmean : W + (long long)blk * Mp * Mp;
  const int njt = is_mean ? 1 : Mp / 64;
  float ssq[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) ssq[u] = 0.f;

  for (int jt = 0; jt < njt; ++jt) {
    float acc[8][4];
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int w = 0; w < 4; ++w) acc[u][w] = 0.f;
    for (int k0 = 0; k0 < Mp; k0 += 16) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int idx = tid + e * 256;
        const int kk = idx & 15, r = idx >> 4;
        const int t = t0 + r;
        As[kk][r] = (t < T) ? Kt[(long long)t * ld + k0 + kk] : 0.f;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int idx = tid + e * 256;
        const int kk = idx & 15, r = idx >> 4;
        Bs[kk][r] = Wb[(long long)(jt * 64 + r) * Mp + k0 + kk];
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) {
        float a[8], b[4];
#pragma unroll
        for (int u = 0; u < 8; ++u) a[u] = As[kk][ty * 8 + u];
#pragma unroll
        for (int w = 0; w < 4; ++w) b[w] = Bs[kk][tx * 4 + w];
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
          for (int w = 0; w < 4; ++w) acc[u][w] = fmaf(a[u], b[w], acc[u][w]);
      }
      __syncthreads();
    }
    if (is_mean) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int t = t0 + ty * 8 + u;
        if (t >= T) continue;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const int r = tx * 4 + w;
          if (r < R) mean_out[(long long)t * R + r] = acc[u][w];
        }
      }
    } else {
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int w = 0; w < 4; ++w) ssq[u] = fmaf(acc[u][w], acc[u][w], ssq[u]);
    }
  }
  if (!is_mean) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      float s = ssq[u];
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);  // 16 lanes share a row
      const int t = t0 + ty * 8 + u;
      if (tx == 0 && t < T) acc_out[(long long)t * (R + 1) + blk] = s;
    }
  }
}

int launch_cond_simt(const float* Kt, int T, int ld, int Mp, const float* W, const float* Wmean, int R, float* acc,
                     float* mean, cudaStream_t st) {
  if (R > 64) { set_error("cond_simt: R > 64 unsupported"); return DCGP_ERR_ARG; }
  dim3 grid(ceil_div(T, 128), R + 2);
  cond_simt_kernel<<<grid, 256, 0, st>>>(Kt, T, ld, Mp, W, Wmean, R, acc, mean);
  return check_launch("cond_simt");
}

// ------------------------------------------------------------------------------------------ finalize (+ a8 sample)
// var[t,r] = Knn(t) - acc[t,0] + acc[t,1+r]   (conditionals.py:40,65),  output layouts of layers.py:128-131,
// then DS/utils.py:41: sample = mean + z * sqrt(var + jitter).  `n_rep` replicates rows (DS/dgp.py:63).
constexpr float kVarFloor = 1e-12f;   // floor of var + jitter under the square root (reparameterised sample and its gradient)
__global__ void finalize_kernel(const float* __restrict__ acc, const float* __restrict__ mean_t, int T, int R,
                                float knn_const, const float* __restrict__ knn_vec, int n_rep,
                                const float* __restrict__ z, float jitter, float* __restrict__ mean,
                                float* __restrict__ var, float* __restrict__ sample) {
  const long long per = (long long)T * R;
  const long long total = per * n_rep;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    long long i;
    int t, r;
    if (total < (1LL << 32)) {                    // 32-bit index arithmetic (a 64-bit division costs more than the rest of the loop)
      const unsigned iu = (unsigned)e % (unsigned)per;
      i = iu; t = (int)(iu / (unsigned)R); r = (int)(iu - (unsigned)t * (unsigned)R);
    } else {
      i = e % per; t = (int)(i / R); r = (int)(i % R);
    }
    const float knn = knn_vec ? knn_vec[t] : knn_const;
    const float m = mean_t[i];
    const float v = knn - acc[(long long)t * (R + 1)] + acc[(long long)t * (R + 1) + 1 + r];
    mean[e] = m;
    var[e] = v;
    // The float64 reference cannot see var + jitter < 0; the fp32-class variance can (cancellation of size ~1e-4 sigma^2
    // against a true variance ~ 0): clamp the argument of the square root only, `var` is reported as computed.
    if (sample) sample[e] = m + z[e] * sqrtf(fmaxf(v + jitter, kVarFloor));
  }
}
int launch_finalize(const float* acc, const float* mean_t, int T, int R, float knn_const, const float* knn_vec,
                    int n_rep, const float* z, float jitter, float* mean, float* var, float* sample, cudaStream_t st) {
  const long long total = (long long)T * R * n_rep;
  const int blocks = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
  finalize_kernel<<<blocks > 0 ? blocks : 1, 256, 0, st>>>(acc, mean_t, T, R, knn_const, knn_vec, n_rep, z, jitter, mean,
                                                          var, z ? sample : nullptr);
  return check_launch("finalize");
}

// conditional() API mirror: fmean[N,P,R] is mean_t as is; fvar[R,P,N] = Knn[p,n] - acc0 + acc_r  (conditionals.py:40-41,65)
__global__ void finalize_ref_layout_kernel(const float* __restrict__ acc, const float* __restrict__ Knn, int P, int N, int R,
                                           float* __restrict__ fvar) {
  const long long total = (long long)R * P * N;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(e % N);
    const int p = (int)((e / N) % P);
    const int r = (int)(e / ((long long)N * P));
    const long long t = (long long)n * P + p;
    fvar[e] = Knn[(long long)p * N + n] - acc[t * (R + 1)] + acc[t * (R + 1) + 1 + r];
  }
}
int launch_finalize_ref_layout(const float* acc, const float* Knn, int P, int N, int R, float* fvar, cudaStream_t st) {
  finalize_ref_layout_kernel<<<148 * 4, 256, 0, st>>>(acc, Knn, P, N, R, fvar);
  return check_launch("finalize_ref_layout");
}

__global__ void reparam_kernel(const float* __restrict__ mean, const float* __restrict__ var, const float* __restrict__ z,
                               size_t n, float jitter, float* __restrict__ out) {
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x)
    out[e] = mean[e] + z[e] * sqrtf(fmaxf(var[e] + jitter, kVarFloor));
}
int launch_reparam(const float* mean, const float* var, const float* z, size_t n, float jitter, float* out, cudaStream_t st) {
  reparam_kernel<<<148 * 8, 256, 0, st>>>(mean, var, z, n, jitter, out);
  return check_launch("reparam");
}

// ------------------------------------------------------------------------------------------ a7: ConvKernel pieces
// kernels.py:127-133: Kzx_t[n, m] = (1/P) sum_p w_p K[(n*P+p), m]   (trans=1 writes the reference's [M,N] layout)
__global__ void patch_mean_kernel(const float* __restrict__ Kt, int n_rows, int P, int ld, int M, const double* __restrict__ w,
                                  int trans, int ldo, float* __restrict__ out) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (m >= ld) return;
  float acc = 0.f;
  const float* src = Kt + (long long)n * P * ld + m;
  for (int p = 0; p < P; ++p) acc = fmaf(w ? (float)w[p] : 1.f, src[(long long)p * ld], acc);
  acc /= (float)P;
  if (trans) { if (m < M) out[(long long)m * n_rows + n] = acc; }
  else if (m < ldo) out[(long long)n * ldo + m] = (m < M) ? acc : 0.f;
}
int launch_patch_mean(const float* Kt, int n_rows, int P, int ld, int M, const double* w, int trans, int ldo, float* out,
                      cudaStream_t st) {
  dim3 grid(ceil_div(ld, 128), n_rows);
  patch_mean_kernel<<<grid, 128, 0, st>>>(Kt, n_rows, P, ld, M, w, trans, ldo, out);
  return check_launch("patch_mean");
}

// kernels.py:106-115: Kdiag[n] = (1/P^2) sum_{p,p'} w_p w_p' k(x_np, x_np').  One CTA per image; the image lives in
// shared memory (when it fits) and every patch pair is formed by im2col indexing; symmetric pairs counted twice.
__global__ void __launch_bounds__(256) kdiag_kernel(const float* __restrict__ X, View v, const double* __restrict__ w,
                                                    float variance, float inv_ls2, int use_smem, float* __restrict__ out) {
  extern __shared__ float img_s[];
  __shared__ double red[8];
  const int n = blockIdx.x;
  const float* __restrict__ img = X + (long long)n * v.HWC;
  const float* Dm = nullptr;                       // use_smem == 2: pair distances from the tiled pass (patch_pair_sqdist)
  if (use_smem) {
    for (int e = threadIdx.x; e < v.HWC; e += 256) img_s[e] = img[e];
    if (use_smem == 2) {
      float* dm = img_s + v.HWC;                   // [P][P + 1]
      int* pb = reinterpret_cast<int*>(dm + v.P * (v.P + 1));
      int* off = pb + v.P;
      for (int e = threadIdx.x; e < v.P; e += 256) pb[e] = v.patch_base(e);
      for (int e = threadIdx.x; e < v.L; e += 256) off[e] = v.elem_off(e);
      __syncthreads();
      patch_pair_sqdist(img_s, v.P, v.L, pb, off, dm, v.P + 1);
      Dm = dm;
    }
    __syncthreads();
    img = img_s;
  }
  const int fC = v.f * v.C, rowstride = v.W * v.C;
  double acc = 0.0;
  const long long npairs = (long long)v.P * (v.P + 1) / 2;
  for (long long e = threadIdx.x; e < npairs; e += 256) {
    // unrank (p >= q) from e = p(p+1)/2 + q
    int p = (int)((sqrt(8.0 * (double)e + 1.0) - 1.0) * 0.5);
    while ((long long)p * (p + 1) / 2 > e) --p;
    while ((long long)(p + 1) * (p + 2) / 2 <= e) ++p;
    const int q = (int)(e - (long long)p * (p + 1) / 2);
    float k;
    if (p == q) {
      k = variance;
    } else if (Dm) {
      k = 2.f * variance * expf(-0.5f * Dm[p * (v.P + 1) + q] * inv_ls2);
    } else {
      const float* a = img + v.patch_base(p);
      const float* b = img + v.patch_base(q);
      float d = 0.f;
      for (int dy = 0; dy < v.f; ++dy) {
        for (int c = 0; c < fC; ++c) {
          const float t = a[c] - b[c];
          d = fmaf(t, t, d);
        }
        a += rowstride;
        b += rowstride;
      }
      k = 2.f * variance * expf(-0.5f * d * inv_ls2);
    }
    const double ww = w ? w[p] * w[q] : 1.0;
    acc += ww * (double)k;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += red[i];
    out[n] = (float)(s / ((double)v.P * (double)v.P));
  }
}
int launch_kdiag(const float* X, const View& v, int n_rows, const double* w, float variance, float inv_ls2, float* out,
                 cudaStream_t st) {
  size_t bytes = (size_t)v.HWC * sizeof(float);
  int use_smem = bytes <= 200 * 1024;
  const size_t tiled = bytes + ((size_t)v.P * (v.P + 1) + v.P + v.L) * sizeof(float);
  if (tiled <= 96 * 1024) { use_smem = 2; bytes = tiled; }     // image + pair-distance matrix + im2col tables
  if (use_smem && bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kdiag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { set_error("kdiag smem attr: %s", cudaGetErrorString(e)); return DCGP_ERR_CUDA; }
  }
  kdiag_kernel<<<n_rows, 256, use_smem ? bytes : 0, st>>>(X, v, w, variance, inv_ls2, use_smem, out);
  return check_launch("kdiag");
}

// ------------------------------------------------------------------------------------------ a8: N(0,1) draws
// tf.random_normal of DS/layers.py:104, as a COUNTER-BASED generator (Philox-4x32-10 + Box-Muller) keyed by
// (seed, step, layer) and indexed by (sample, GLOBAL image index, output): the draw of an image does not depend on which
// rank holds it or on how many ranks there are, so 1/2/4/8-GPU runs of the same step see identical noise.
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

__global__ void __launch_bounds__(256) randn_kernel(float* __restrict__ z, int S, int n_local, int D, long long n_global,
                                                    long long n0, unsigned long long seed, unsigned long long step, int layer) {
  const long long total = (long long)S * n_local * D;
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += 256LL * gridDim.x) {
    int d, n, s;
    if (total < (1LL << 32)) {
      const unsigned q = (unsigned)e / (unsigned)D;
      d = (int)((unsigned)e - q * (unsigned)D);
      s = (int)(q / (unsigned)n_local);
      n = (int)(q - (unsigned)s * (unsigned)n_local);
    } else {
      d = (int)(e % D);
      const long long q = e / D;
      n = (int)(q % n_local); s = (int)(q / n_local);
    }
    const unsigned long long idx = ((unsigned long long)s * (unsigned long long)n_global + (unsigned long long)(n0 + n)) * D + d;
    uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)step, (uint32_t)(step >> 32) ^ ((uint32_t)layer << 16)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const float u1 = ((float)(c[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);      // (0, 1)
    const float u2 = ((float)(c[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    z[e] = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
  }
}

int launch_randn(float* z, int S, int n_local, int D, long long n_global, long long n0, unsigned long long seed,
                 unsigned long long step, int layer, cudaStream_t st) {
  const long long total = (long long)S * n_local * D;
  if (total <= 0) return DCGP_OK;
  long long blocks = (total + 1023) / 1024;
  if (blocks > 148 * 8) blocks = 148 * 8;
  randn_kernel<<<(unsigned)blocks, 256, 0, st>>>(z, S, n_local, D, n_global, n0, seed, step, layer);
  return check_launch("randn");
}

// ------------------------------------------------------------------------------------------ a9: likelihood
__constant__ double c_gh_x[20];
__constant__ double c_gh_w[20];  // already divided by sqrt(pi)

// GPflow MultiClass(RobustMax).variational_expectations with 20-point Gauss-Hermite.
// One thread per (s, n) row; K <= 16 classes.
__global__ void __launch_bounds__(128) varexp_kernel(const float* __restrict__ Fmu, const float* __restrict__ Fvar,
                                                     const int32_t* __restrict__ Y, int SN, int N, int K, double log1meps,
                                                     double logepsk, double* __restrict__ varexp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= SN) return;
  const int y = min(max(Y[i % N], 0), K - 1);   // labels are validated by the host (likelihoods.py); never index out of range
  double mu[16], isd[16];
  for (int k = 0; k < K; ++k) {
    mu[k] = (double)Fmu[(long long)i * K + k];
    const double v = fmax((double)Fvar[(long long)i * K + k], 1e-10);
    isd[k] = 1.0 / sqrt(v);
  }
  const double mu_y = mu[y];
  const double sd2 = sqrt(fmax(2.0 * (double)Fvar[(long long)i * K + y], 1e-10));
  double p = 0.0;
  for (int g = 0; g < 20; ++g) {
    const double x = mu_y + c_gh_x[g] * sd2;
    double prod = 1.0;
    for (int k = 0; k < K; ++k) {
      if (k == y) continue;
      const double dist = (x - mu[k]) * isd[k];
      double cdf = 0.5 * (1.0 + erf(dist * 0.70710678118654752440));
      cdf = cdf * (1.0 - 2e-4) + 1e-4;
      prod *= cdf;
    }
    p += prod * c_gh_w[g];
  }
  varexp[i] = p * log1meps + (1.0 - p) * logepsk;
}

// GPflow MultiClass(RobustMax).predict_mean_and_var / predict_density (the prediction path of DS/dgp.py:116-126): for every
// class c, p_c = P(f_c is the largest) by the same 20-point quadrature, then ps = p_c (1 - eps) + (1 - p_c) eps/(K-1).
// One thread per (row, class).  pmean / pvar [SN, K] (var = ps - ps^2) and, when Y is given, logdens[SN] = log ps[y].
__global__ void __launch_bounds__(128) multiclass_predict_kernel(const float* __restrict__ Fmu, const float* __restrict__ Fvar,
                                                                 const int32_t* __restrict__ Y, int SN, int N, int K, double eps,
                                                                 double* __restrict__ pmean, double* __restrict__ pvar,
                                                                 double* __restrict__ logdens) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= (long long)SN * K) return;
  const int i = (int)(e / K), c = (int)(e - (long long)i * K);
  double mu[16], isd[16];
  for (int k = 0; k < K; ++k) {
    mu[k] = (double)Fmu[(long long)i * K + k];
    const double v = fmax((double)Fvar[(long long)i * K + k], 1e-10);
    isd[k] = 1.0 / sqrt(v);
  }
  const double mu_c = mu[c];
  const double sd2 = sqrt(fmax(2.0 * (double)Fvar[(long long)i * K + c], 1e-10));
  double p = 0.0;
  for (int g = 0; g < 20; ++g) {
    const double x = mu_c + c_gh_x[g] * sd2;
    double prod = 1.0;
    for (int k = 0; k < K; ++k) {
      if (k == c) continue;
      const double dist = (x - mu[k]) * isd[k];
      double cdf = 0.5 * (1.0 + erf(dist * 0.70710678118654752440));
      cdf = cdf * (1.0 - 2e-4) + 1e-4;
      prod *= cdf;
    }
    p += prod * c_gh_w[g];
  }
  const double ps = p * (1.0 - eps) + (1.0 - p) * (eps / (K - 1.0));
  if (pmean) pmean[e] = ps;
  if (pvar) pvar[e] = ps - ps * ps;
  if (logdens && Y && Y[i % N] == c) logdens[i] = log(ps);
}

__global__ void __launch_bounds__(1024) sum_f64_kernel(const double* __restrict__ x, int n, double* __restrict__ out) {
  __shared__ double sh[32];
  double v = 0.0;
  for (int e = threadIdx.x; e < n; e += 1024) v += x[e];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = sh[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) *out = v;
  }
}

// Gradient of the expected log-likelihood w.r.t. Fmu, Fvar (times `coef`), same quadrature as varexp_kernel:
//   p = sum_g w_g prod_{k!=y} c(u_gk),  u_gk = (mu_y + x_g sqrt(2 v_y) - mu_k) / sqrt(v_k),  c = Phi*(1-2e-4)+1e-4
__global__ void __launch_bounds__(128) varexp_grad_kernel(const float* __restrict__ Fmu, const float* __restrict__ Fvar,
                                                          const int32_t* __restrict__ Y, int SN, int N, int K, double dlog,
                                                          double coef, float* __restrict__ gmu, float* __restrict__ gvar) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= SN) return;
  const int y = min(max(Y[i % N], 0), K - 1);   // labels are validated by the host (likelihoods.py); never index out of range
  double mu[16], isd[16], vk[16], dmu[16], dv[16];
  bool vclip[16];
  for (int k = 0; k < K; ++k) {
    mu[k] = (double)Fmu[(long long)i * K + k];
    const double v = (double)Fvar[(long long)i * K + k];
    vclip[k] = !(v > 1e-10);
    vk[k] = fmax(v, 1e-10);
    isd[k] = 1.0 / sqrt(vk[k]);
    dmu[k] = 0.0;
    dv[k] = 0.0;
  }
  const double vy2 = 2.0 * (double)Fvar[(long long)i * K + y];
  const bool yclip = !(vy2 > 1e-10);
  const double sd2 = sqrt(fmax(vy2, 1e-10));
  for (int g = 0; g < 20; ++g) {
    const double x = mu[y] + c_gh_x[g] * sd2;
    double c[16], u[16], prod = 1.0;
    for (int k = 0; k < K; ++k) {
      if (k == y) continue;
      u[k] = (x - mu[k]) * isd[k];
      c[k] = 0.5 * (1.0 + erf(u[k] * 0.70710678118654752440)) * (1.0 - 2e-4) + 1e-4;
      prod *= c[k];
    }
    for (int k = 0; k < K; ++k) {
      if (k == y) continue;
      const double phi = 0.3989422804014327 * exp(-0.5 * u[k] * u[k]) * (1.0 - 2e-4);
      const double A = c_gh_w[g] * (prod / c[k]) * phi;       // d p / d u_gk (weighted)
      dmu[k] -= A * isd[k];
      if (!vclip[k]) dv[k] -= A * u[k] / (2.0 * vk[k]);
      dmu[y] += A * isd[k];
      if (!yclip) dv[y] += A * isd[k] * c_gh_x[g] / sd2;       // d x_g / d v_y = gh_x / sqrt(2 v_y)
    }
  }
  const double f = coef * dlog;   // dlog = log(1-eps) - log(eps/(K-1))
  for (int k = 0; k < K; ++k) {
    gmu[(long long)i * K + k] = (float)(f * dmu[k]);
    gvar[(long long)i * K + k] = (float)(f * dv[k]);
  }
}

// DS/utils.py:41 backward: F = mean + z sqrt(var + jitter)  ->  g_mean = gF, g_var = gF * z / (2 sqrt(var + jitter))
__global__ void sample_backward_kernel(const float* __restrict__ gF, const float* __restrict__ z, const float* __restrict__ var,
                                       size_t n, float jitter, float* __restrict__ g_mean, float* __restrict__ g_var) {
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const float g = gF[e];
    g_mean[e] = g;
    const float sv = var[e] + jitter;      // where the forward clamped the square root the sample does not depend on var
    g_var[e] = sv > kVarFloor ? g * z[e] * 0.5f * rsqrtf(sv) : 0.f;
  }
}
int launch_sample_backward(const float* gF, const float* z, const float* var, size_t n, float jitter, float* g_mean, float* g_var,
                           cudaStream_t st) {
  sample_backward_kernel<<<148 * 8, 256, 0, st>>>(gF, z, var, n, jitter, g_mean, g_var);
  return check_launch("sample_backward");
}

// Adam (experiment.py:97-99, tf.train.AdamOptimizer semantics) on a flat float64 parameter vector; maximize=1 ascends.
__global__ void adam_kernel(double* __restrict__ p, const double* __restrict__ g, double* __restrict__ m, double* __restrict__ v,
                            size_t n, double lr_t, double b1, double b2, double eps, double sign) {
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const double ge = sign * g[e];
    const double me = b1 * m[e] + (1.0 - b1) * ge;
    const double ve = b2 * v[e] + (1.0 - b2) * ge * ge;
    m[e] = me;
    v[e] = ve;
    p[e] -= lr_t * me / (sqrt(ve) + eps);
  }
}
int launch_adam(double* param, const double* grad, double* m, double* v, size_t n, double lr, double b1, double b2, double eps,
                int step, int maximize, cudaStream_t st) {
  const double lr_t = lr * sqrt(1.0 - pow(b2, step)) / (1.0 - pow(b1, step));   // TF's bias-corrected step size
  adam_kernel<<<148 * 4, 256, 0, st>>>(param, grad, m, v, n, lr_t, b1, b2, eps, maximize ? -1.0 : 1.0);
  return check_launch("adam");
}

static void gauss_hermite_20(double* x, double* w) {
  // Newton iteration on the orthonormal Hermite recurrence (classic `gauher`), n = 20.
  const int n = 20;
  const double pim4 = 0.7511255444649425;
  double z = 0.0;
  for (int i = 0; i < (n + 1) / 2; ++i) {
    if (i == 0) z = sqrt((double)(2 * n + 1)) - 1.85575 * pow((double)(2 * n + 1), -0.16667);
    else if (i == 1) z -= 1.14 * pow((double)n, 0.426) / z;
    else if (i == 2) z = 1.86 * z - 0.86 * x[0];
    else if (i == 3) z = 1.91 * z - 0.91 * x[1];
    else z = 2.0 * z - x[i - 2];
    double pp = 0.0;
    for (int its = 0; its < 100; ++its) {
      double p1 = pim4, p2 = 0.0;
      for (int j = 0; j < n; ++j) {
        const double p3 = p2;
        p2 = p1;
        p1 = z * sqrt(2.0 / (j + 1)) * p2 - sqrt((double)j / (j + 1)) * p3;
      }
      pp = sqrt(2.0 * n) * p2;
      const double z1 = z;
      z = z1 - p1 / pp;
      if (fabs(z - z1) <= 1e-15) break;
    }
    x[i] = z;
    x[n - 1 - i] = -z;
    w[i] = 2.0 / (pp * pp);
    w[n - 1 - i] = w[i];
  }
}

static void init_gh();
int launch_varexp_grad(const float* Fmu, const float* Fvar, const int32_t* Y, int S, int N, int K, double epsilon, double coef,
                       float* gmu, float* gvar, cudaStream_t st) {
  if (K > 16 || K < 2) { set_error("varexp_grad: K must be in [2,16]"); return DCGP_ERR_ARG; }
  init_gh();
  const int SN = S * N;
  varexp_grad_kernel<<<ceil_div(SN, 128), 128, 0, st>>>(Fmu, Fvar, Y, SN, N, K, log(1.0 - epsilon) - log(epsilon / (K - 1.0)), coef,
                                                        gmu, gvar);
  return check_launch("varexp_grad");
}

static void init_gh() {
  // __constant__ memory is per device: load the tables once for every device this process uses
  static std::mutex mu;
  static bool done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  if (dev < 0 || dev >= 64 || done[dev]) return;
  double x[20], w[20];
  gauss_hermite_20(x, w);
  for (int i = 0; i < 20; ++i) w[i] /= sqrt(M_PI);
  if (cudaMemcpyToSymbol(c_gh_x, x, sizeof(x)) != cudaSuccess || cudaMemcpyToSymbol(c_gh_w, w, sizeof(w)) != cudaSuccess) {
    set_error("Gauss-Hermite tables: %s", cudaGetErrorString(cudaGetLastError()));
    return;
  }
  done[dev] = true;
}

int launch_varexp(const float* Fmu, const float* Fvar, const int32_t* Y, int S, int N, int K, double epsilon, double* varexp,
                  double* sum, cudaStream_t st) {
  if (K > 16 || K < 2) { set_error("varexp: K must be in [2,16]"); return DCGP_ERR_ARG; }
  init_gh();
  const int SN = S * N;
  varexp_kernel<<<ceil_div(SN, 128), 128, 0, st>>>(Fmu, Fvar, Y, SN, N, K, log(1.0 - epsilon), log(epsilon / (K - 1.0)), varexp);
  sum_f64_kernel<<<1, 1024, 0, st>>>(varexp, SN, sum);
  return check_launch("varexp", 2);
}

int launch_multiclass_predict(const float* Fmu, const float* Fvar, const int32_t* Y, int S, int N, int K, double epsilon,
                              double* pmean, double* pvar, double* logdens, cudaStream_t st) {
  if (K > 16 || K < 2) { set_error("multiclass_predict: K must be in [2,16]"); return DCGP_ERR_ARG; }
  init_gh();
  const long long n = (long long)S * N * K;
  multiclass_predict_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(Fmu, Fvar, Y, S * N, N, K, epsilon, pmean, pvar, logdens);
  return check_launch("multiclass_predict");
}

__global__ void elbo_kernel(const double* sum_varexp, int S, double scale, const double* kls, int n_layers, double* elbo) {
  double kl = 0.0;
  for (int i = 0; i < n_layers; ++i) kl += kls[i];
  *elbo = (*sum_varexp / S) * scale - kl;
}
int launch_elbo(const double* sum_varexp, int S, double scale, const double* kls, int n_layers, double* elbo, cudaStream_t st) {
  elbo_kernel<<<1, 1, 0, st>>>(sum_varexp, S, scale, kls, n_layers, elbo);
  return check_launch("elbo");
}

// ------------------------------------------------------------------------------------------ packing (f64 -> f32 operands)
__global__ void pack_z_kernel(const double* __restrict__ Z, long long n, double inv_ls, float* __restrict__ zs,
                              const double* __restrict__ hyp) {
  if (hyp) inv_ls = 1.0 / hyp[1];
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    zs[e] = (float)(Z[e] * inv_ls);
}
int launch_pack_z(const double* Z, long long n, double inv_ls, float* zs, cudaStream_t st, const double* hyp) {
  pack_z_kernel<<<148, 256, 0, st>>>(Z, n, inv_ls, zs, hyp);
  return check_launch("pack_z");
}

// W32[blk][i][j]: blk 0 = Linv (ld ldl), blk r>=1 = Wr[r-1] (M x M, ld M); zero padded to Mp.
__global__ void pack_w_kernel(const double* __restrict__ Linv, int ldl, const double* __restrict__ Wr, int M, int Mp, int R,
                              float* __restrict__ W) {
  const long long total = (long long)(R + 1) * Mp * Mp;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(e % Mp);
    const int i = (int)((e / Mp) % Mp);
    const int blk = (int)(e / ((long long)Mp * Mp));
    double v = 0.0;
    if (i < M && j < M) v = (blk == 0) ? Linv[(long long)i * ldl + j] : Wr[((long long)(blk - 1) * M + i) * M + j];
    W[e] = (float)v;
  }
}
int launch_pack_w(const double* Linv, int ldl, const double* Wr, int M, int Mp, int R, float* W, cudaStream_t st) {
  pack_w_kernel<<<148 * 8, 256, 0, st>>>(Linv, ldl, Wr, M, Mp, R, W);
  return check_launch("pack_w");
}

// Wmean[r][m] = beta[m][r] (beta is [M,R]); rows >= R and cols >= M are zero.
__global__ void pack_wmean_kernel(const double* __restrict__ beta, int M, int Mp, int R, int RP, float* __restrict__ Wm) {
  const int total = RP * Mp;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int m = e % Mp, r = e / Mp;
    Wm[e] = (r < R && m < M) ? (float)beta[(long long)m * R + r] : 0.f;
  }
}
int launch_pack_wmean(const double* beta, int M, int Mp, int R, int RP, float* Wm, cudaStream_t st) {
  pack_wmean_kernel<<<148, 256, 0, st>>>(beta, M, Mp, R, RP, Wm);
  return check_launch("pack_wmean");
}

// KL from its four reductions (GPflow gauss_kl / DS/layers.py:242-256):  sc = {mahal, trace, logdet_q, logdet_p}
__global__ void kl_kernel(const double* sc, int M, int R, int white, double* kl) {
  double two = sc[0] - (double)M * R - sc[2] + sc[1];
  if (!white) two += (double)R * sc[3];
  *kl = 0.5 * two;
}
int launch_kl(const double* sc, int M, int R, int white, double* kl, cudaStream_t st) {
  kl_kernel<<<1, 1, 0, st>>>(sc, M, R, white, kl);
  return check_launch("kl");
}

}  // namespace dcgp
