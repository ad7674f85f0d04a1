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
