// Kernel-launcher declarations shared by dcgp_api.cu, dcgp_simt.cu and dcgp_tc.cu.
#pragma once
#include "dcgp_common.cuh"

namespace dcgp {

// views.py:20-30,56-68 FullView geometry, plus the im2col index math that replaces tf.extract_image_patches.
struct View {
  int H, W, C, f, s, OH, OW, P, L, HWC;
  __host__ __device__ int patch_base(int p) const {  // offset of patch p's top-left pixel inside one image
    const int oy = p / OW, ox = p - oy * OW;
    return (oy * s * W + ox * s) * C;
  }
  __host__ __device__ int elem_off(int l) const {  // offset of patch element l = (dy*f+dx)*C+c from that pixel
    const int fC = f * C;
    const int dy = l / fC;
    return dy * W * C + (l - dy * fC);
  }
};
static inline View make_view(int H, int W, int C, int f, int s) {
  View v;
  v.H = H; v.W = W; v.C = C; v.f = f; v.s = s;
  v.OH = (H - f) / s + 1;  // views.py:65-68
  v.OW = (W - f) / s + 1;
  v.P = v.OH * v.OW;
  v.L = f * f * C;
  v.HWC = H * W * C;
  return v;
}

// Squared distances between all pairs of patches of ONE image held in shared memory (kernels.py:106-115, Kdiag and its
// backward): Dm[p * ldm + q] = |x_p - x_q|^2 for q < p.  4 x 4 patch tiles -- eight loads feed sixteen differences per patch
// element instead of two loads per difference -- each tile by four adjacent lanes that take a quarter of the patch elements
// each and add up by shuffle in a fixed order (deterministic).  pb[p]: image offset of patch p, off[l]: image offset of patch
// element l (both in shared memory); every thread of the CTA must call (blockDim.x a multiple of 32).
#ifdef __CUDACC__
static __device__ __forceinline__ void patch_pair_sqdist(const float* __restrict__ img, int P, int L, const int* __restrict__ pb,
                                                         const int* __restrict__ off, float* __restrict__ Dm, int ldm) {
  const int nb = (P + 3) >> 2, ntiles = nb * (nb + 1) / 2, per_round = blockDim.x >> 2;
  const int sub = threadIdx.x & 3;
  const int l0 = (L * sub) >> 2, l1 = (L * (sub + 1)) >> 2;
  for (int t0 = 0; t0 < ntiles; t0 += per_round) {
    const int tile = t0 + (threadIdx.x >> 2);
    const bool live = tile < ntiles;
    int bp = 0, bq = 0;
    if (live) {                                      // unrank (bp >= bq) from tile = bp (bp + 1) / 2 + bq
      bp = (int)((sqrtf(8.f * (float)tile + 1.f) - 1.f) * 0.5f);
      while (bp * (bp + 1) / 2 > tile) --bp;
      while ((bp + 1) * (bp + 2) / 2 <= tile) ++bp;
      bq = tile - bp * (bp + 1) / 2;
    }
    int pa[4], qb[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { pa[u] = pb[min(4 * bp + u, P - 1)]; qb[u] = pb[min(4 * bq + u, P - 1)]; }
    float d[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) d[i] = 0.f;
    if (live) {
      for (int l = l0; l < l1; ++l) {
        const int o = off[l];
        float xa[4], xb[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { xa[u] = img[pa[u] + o]; xb[u] = img[qb[u] + o]; }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const float t = xa[u] - xb[w];
            d[4 * u + w] = fmaf(t, t, d[4 * u + w]);
          }
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      d[i] += __shfl_xor_sync(0xffffffffu, d[i], 1);
      d[i] += __shfl_xor_sync(0xffffffffu, d[i], 2);
    }
    if (live) {
      const int p = 4 * bp + sub;                    // lane `sub` stores row `sub` of the tile
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const int q = 4 * bq + w;
        // (d[4 * sub + w] with a compile-time index: select instead of a dynamically indexed register array)
        const float val = sub == 0 ? d[w] : (sub == 1 ? d[4 + w] : (sub == 2 ? d[8 + w] : d[12 + w]));
        if (p < P && q < p) Dm[p * ldm + q] = val;
      }
    }
  }
}
#endif

// ---- dcgp_simt.cu
int launch_patches(const float* X, const View& v, int N, int layout, float* out, cudaStream_t st);
int launch_kuf_simt(const float* X, const View& v, int n_rows, const float* zs, int M, float variance, float inv_ls,
                    int layout, int ldo, float* out, cudaStream_t st);
int launch_pmn_to_tm(const float* Kmn, int P, int M, int N, int ldo, float* out, cudaStream_t st);
int launch_cond_simt(const float* Kt, int T, int ld, int Mp, const float* W, const float* Wmean, int R, float* acc,
                     float* mean, cudaStream_t st);
int launch_finalize(const float* acc, const float* mean_t, int T, int R, float knn_const, const float* knn_vec,
                    int n_rep, const float* z, float jitter, float* mean, float* var, float* sample, cudaStream_t st);
int launch_finalize_ref_layout(const float* acc, const float* Knn, int P, int N, int R, float* fvar, cudaStream_t st);
int launch_randn(float* z, int S, int n_local, int D, long long n_global, long long n0, unsigned long long seed,
                 unsigned long long step, int layer, cudaStream_t st);
int launch_reparam(const float* mean, const float* var, const float* z, size_t n, float jitter, float* out, cudaStream_t st);
int launch_patch_mean(const float* Kt, int n_rows, int P, int ld, int M, const double* w, int trans, int ldo, float* out,
                      cudaStream_t st);
int launch_kdiag(const float* X, const View& v, int n_rows, const double* w, float variance, float inv_ls2, float* out,
                 cudaStream_t st);
int launch_varexp(const float* Fmu, const float* Fvar, const int32_t* Y, int S, int N, int K, double epsilon, double* varexp,
                  double* sum, cudaStream_t st);
int launch_elbo(const double* sum_varexp, int S, double scale, const double* kls, int n_layers, double* elbo, cudaStream_t st);
int launch_pack_z(const double* Z, long long n, double inv_ls, float* zs, cudaStream_t st, const double* hyp = nullptr);
int launch_pack_w(const double* Linv, int ldl, const double* Wr, int M, int Mp, int R, float* W, cudaStream_t st);
int launch_pack_wmean(const double* beta, int M, int Mp, int R, int RP, float* Wm, cudaStream_t st);
int launch_multiclass_predict(const float* Fmu, const float* Fvar, const int32_t* Y, int S, int N, int K, double epsilon,
                              double* pmean, double* pvar, double* logdens, cudaStream_t st);
int launch_varexp_grad(const float* Fmu, const float* Fvar, const int32_t* Y, int S, int N, int K, double epsilon, double coef,
                       float* gmu, float* gvar, cudaStream_t st);
int launch_sample_backward(const float* gF, const float* z, const float* var, size_t n, float jitter, float* g_mean, float* g_var,
                           cudaStream_t st);
int launch_adam(double* param, const double* grad, double* m, double* v, size_t n, double lr, double b1, double b2, double eps,
                int step, int maximize, cudaStream_t st);
int launch_kl(const double* sc, int M, int R, int white, double* kl, cudaStream_t st);

}  // namespace dcgp
