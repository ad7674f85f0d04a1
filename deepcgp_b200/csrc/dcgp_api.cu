// C ABI of libdcgp.so (see include/dcgp.h): host-side orchestration of the kernels.
// No device allocation happens here: every buffer comes from the caller; all work is stream-ordered.
#include <string.h>

#include <map>
#include <mutex>
#include <utility>

#include "dcgp_kernels.cuh"
#include "dcgp_tc.cuh"

namespace dcgp {
const char* last_error();

#define DCGP_TRY(expr)            \
  do {                            \
    int _rc = (expr);             \
    if (_rc != DCGP_OK) return _rc; \
  } while (0)

// Carves a caller-provided workspace into aligned sub-buffers (256 B) and tracks the high-water mark.
struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* p) : base((char*)p) {}
  template <typename T>
  T* take(size_t n) {
    off = align_up(off, 256);
    T* r = (T*)(base ? base + off : nullptr);
    off += n * sizeof(T);
    return r;
  }
};

// ---------------------------------------------------------------------------------------------- prepared layer
// Opaque `prep` buffer consumed by dcgp_layer_apply: the minibatch-independent operands of one layer.
struct Prep {
  int M, Mp, R, RP, L;
  float* zs;      // [M, L]  Z / lengthscale, fp32 (SIMT Kuf operand)
  float* W;       // [(R+1)*Mp, Mp] fp32: block 0 = Lm^-1, block r = L_r^T G   (SIMT conditional operand)
  float* Wmean;   // [RP, Mp]      fp32: rows r < R = (G^T q_mu)[:, r]
  TcPrep tc;      // split-fp16 planes of the same operands (tensor-core path)
  size_t bytes;
};

static Prep carve_prep(const dcgp_layer_desc* d, void* buf) {
  Prep p;
  p.M = d->M;
  p.Mp = (int)align_up(d->M, 64);
  p.R = d->R;
  p.RP = 64;
  p.L = d->f * d->f * d->C;
  Carver c(buf);
  p.zs = c.take<float>((size_t)p.M * p.L);
  p.W = c.take<float>((size_t)(p.R + 1) * p.Mp * p.Mp);
  p.Wmean = c.take<float>((size_t)p.RP * p.Mp);
  tc_carve_prep(p.tc, p.M, p.Mp, p.R, p.L, c.base ? c.base + align_up(c.off, 1024) : nullptr);
  c.off = align_up(c.off, 1024) + p.tc.bytes;
  p.bytes = align_up(c.off, 256);
  return p;
}

struct F64Work {
  int M, Mq, R;
  double *Kuu, *invD, *Linv, *Kinv, *Wr, *beta, *alpha, *Lpinv, *invDp, *tmpMR, *tmpK, *sc;
  void *trws, *trws2;
  size_t bytes;
};

static F64Work carve_f64(int M, int R, void* ws) {
  F64Work w;
  w.M = M;
  w.R = R;
  w.Mq = trtri_pad(M);
  Carver c(ws);
  w.Kuu = c.take<double>((size_t)M * M);
  w.invD = c.take<double>(potrf_ws_bytes(M) / sizeof(double));
  w.Linv = c.take<double>((size_t)w.Mq * w.Mq);
  w.trws = c.take<double>(trtri_ws_bytes(M) / sizeof(double));
  w.trws2 = c.take<double>(trtri_ws_bytes(M) / sizeof(double));
  w.Kinv = c.take<double>((size_t)M * M);      // also holds Kuu(Z_prior) / its Cholesky factor for the KL
  w.Wr = c.take<double>((size_t)R * M * M);    // also reused as Lp^-1 L_r for the KL trace
  w.beta = c.take<double>((size_t)M * R);
  w.Lpinv = c.take<double>((size_t)w.Mq * w.Mq);
  w.invDp = c.take<double>(potrf_ws_bytes(M) / sizeof(double));
  w.tmpMR = c.take<double>((size_t)M * R);
  w.tmpK = c.take<double>((size_t)M * M);
  w.sc = c.take<double>(8);
  w.alpha = c.take<double>((size_t)M * R);     // Lm^-1 q_mu: mean operand of the chained conditional
  w.bytes = align_up(c.off, 256);
  return w;
}

// Kuu (in w.Kuu) -> Lm (in place), Lm^-1, and the stacked conditional operand:
//   non-white: G = Kuu^-1,  W_r = L_r^T G (= C_r^T Lm^-1, C_r = Lm^-1 L_r),  beta = G q_mu        (conditionals.py:44-58)
//   white    : G = Lm^-1,   W_r = L_r^T G,                                   beta = G^T q_mu
struct GInfo { const double* G; int ldg; };

static int factor_only(const F64Work& w, int* info, cudaStream_t st) {
  DCGP_TRY(potrf_f64(w.Kuu, w.M, w.M, w.invD, info, st));
  return trtri_f64(w.Kuu, w.M, w.M, w.invD, w.Linv, w.trws, st);
}

static int g_and_beta(const F64Work& w, int white, const double* q_mu, GInfo* gi, cudaStream_t st) {
  const int M = w.M, R = w.R, Mq = w.Mq;
  gi->G = w.Linv;
  gi->ldg = Mq;
  if (!white) {
    GemmF64 g{};
    g.m = g.n = g.k = M;
    g.A = w.Linv; g.lda = Mq; g.transA = 1; g.lowerA = 1;
    g.B = w.Linv; g.ldb = Mq; g.lowerB = 1;
    g.C = w.Kinv; g.ldc = M; g.alpha = 1.0; g.batch = 1;
    DCGP_TRY(gemm_f64(g, st));
    gi->G = w.Kinv;
    gi->ldg = M;
  }
  GemmF64 g{};  // beta = G^T q_mu
  g.m = M; g.n = R; g.k = M;
  g.A = gi->G; g.lda = gi->ldg; g.transA = 1; g.lowerA = white ? 1 : 0;
  g.B = q_mu; g.ldb = R;
  g.C = w.beta; g.ldc = R; g.alpha = 1.0; g.batch = 1;
  return gemm_f64(g, st);
}

static int factor_and_G(const F64Work& w, int white, const double* q_mu, int* info, GInfo* gi, cudaStream_t st) {
  DCGP_TRY(factor_only(w, info, st));
  return g_and_beta(w, white, q_mu, gi, st);
}

static int wr_f64(const F64Work& w, int white, const GInfo& gi, const double* q_sqrt, cudaStream_t st) {
  const int M = w.M, R = w.R;
  GemmF64 g{};  // W_r = L_r^T G
  g.m = g.n = g.k = M;
  g.A = q_sqrt; g.lda = M; g.transA = 1; g.lowerA = 1; g.strideA = (long long)M * M;
  g.B = gi.G; g.ldb = gi.ldg; g.lowerB = white ? 1 : 0; g.strideB = 0;
  g.C = w.Wr; g.ldc = M; g.strideC = (long long)M * M;
  g.alpha = 1.0; g.batch = R;
  return gemm_f64(g, st);
}

static int build_operands(const F64Work& w, int white, const double* q_mu, const double* q_sqrt, int* info,
                          cudaStream_t st) {
  GInfo gi;
  DCGP_TRY(factor_and_G(w, white, q_mu, info, &gi, st));
  return wr_f64(w, white, gi, q_sqrt, st);
}

// KL[q(u) || p(u)]: GPflow gauss_kl (layers.py:145-147) == the hand-written DS/layers.py:242-256.
// Lp/Lpinv: Cholesky factor of the prior covariance and its inverse (ignored when white).
static int kl_terms(const F64Work& w, int white, const double* Lp, int ldp, const double* Lpinv, int ldpi,
                    const double* q_mu, const double* q_sqrt, double* kl, cudaStream_t st, bool have_trace = false) {
  const int M = w.M, R = w.R;
  if (white) {
    DCGP_TRY(sumsq_f64(q_mu, M, R, R, 0, w.sc + 0, st));
    DCGP_TRY(sumsq_f64(q_sqrt, (long long)R * M, M, M, M, w.sc + 1, st));
  } else {
    GemmF64 g{};  // Lp^-1 q_mu
    g.m = M; g.n = R; g.k = M;
    g.A = Lpinv; g.lda = ldpi; g.lowerA = 1;
    g.B = q_mu; g.ldb = R;
    g.C = w.tmpMR; g.ldc = R; g.alpha = 1.0; g.batch = 1;
    DCGP_TRY(gemm_f64(g, st));
    DCGP_TRY(sumsq_f64(w.tmpMR, M, R, R, 0, w.sc + 0, st));
    if (!have_trace) {   // (the tensor-core path has already accumulated the trace into sc[1])
      GemmF64 h{};  // Lp^-1 L_r
      h.m = h.n = h.k = M;
      h.A = Lpinv; h.lda = ldpi; h.lowerA = 1; h.strideA = 0;
      h.B = q_sqrt; h.ldb = M; h.lowerB = 1; h.strideB = (long long)M * M;
      h.C = w.Wr; h.ldc = M; h.strideC = (long long)M * M;
      h.alpha = 1.0; h.batch = R;
      DCGP_TRY(gemm_f64(h, st));
      DCGP_TRY(sumsq_f64(w.Wr, (long long)R * M, M, M, 0, w.sc + 1, st));
    }
    DCGP_TRY(logdiag2_f64(Lp, ldp, M, 1, 0, w.sc + 3, st));
  }
  DCGP_TRY(logdiag2_f64(q_sqrt, M, M, R, (long long)M * M, w.sc + 2, st));
  return launch_kl(w.sc, M, R, white, kl, st);
}

// The KL prior of a ConvLayer needs a second, independent Cholesky + inverse (Kuu at the initial Z).  Both chains are
// latency-bound (a handful of CTAs), so the prior chain is forked onto an auxiliary stream and joined before the KL.
// One (stream, 2 events) triple per calling stream slot, created lazily; this is the only device-side state the library keeps.
struct AuxFork {
  cudaStream_t aux = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  bool ok = false;
};
// keyed by (device, calling stream); guarded by a mutex; never evicted (a process uses a handful of streams)
static AuxFork& aux_for(cudaStream_t st) {
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, AuxFork> pool;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  AuxFork& a = pool[std::make_pair(dev, st)];
  if (!a.ok) {
    a.ok = cudaStreamCreateWithFlags(&a.aux, cudaStreamNonBlocking) == cudaSuccess &&
           cudaEventCreateWithFlags(&a.fork, cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&a.join, cudaEventDisableTiming) == cudaSuccess;
    if (!a.ok) set_error("auxiliary stream / events: %s", cudaGetErrorString(cudaGetLastError()));
  }
  return a;
}

static int check_desc(const dcgp_layer_desc* d) {
  if (!d) { set_error("null layer descriptor"); return DCGP_ERR_ARG; }
  if (d->kind != DCGP_LAYER_CONV && d->kind != DCGP_LAYER_SVGP_CONV) { set_error("bad layer kind %d", d->kind); return DCGP_ERR_ARG; }
  if (d->H < d->f || d->W < d->f || d->f < 1 || d->s < 1 || d->C < 1) { set_error("bad view geometry"); return DCGP_ERR_ARG; }
  if (d->M < 1 || d->M > 4096 || d->R < 1 || d->R > 64) { set_error("need 1<=M<=4096 and 1<=R<=64"); return DCGP_ERR_ARG; }
  if (!(d->variance > 0) || !(d->lengthscale > 0)) { set_error("variance / lengthscale must be positive"); return DCGP_ERR_ARG; }
  return DCGP_OK;
}

}  // namespace dcgp

using namespace dcgp;

extern "C" {

const char* dcgp_last_error(void) { return dcgp::last_error(); }
int dcgp_version(void) { return 100; }
long long dcgp_launch_count(void) { return dcgp::launch_count(); }
void dcgp_set_kernel_timing(int on) { dcgp::tc_set_timing(on); }
void dcgp_set_reserved_sms(int n) { dcgp::tc_set_reserved_sms(n); }
double dcgp_kernel_ms(int which) { return dcgp::tc_kernel_ms(which); }
double dcgp_kernel_tensor_flops(int which) { return dcgp::tc_kernel_flops(which); }
void dcgp_set_products(int cond, int dk, int dq) { dcgp::tc_set_products(cond, dk, dq); }
void dcgp_set_precise_stage1(int mode) { dcgp::tc_set_precise_stage1(mode); }
int dcgp_get_precise_stage1(void) { return dcgp::tc_get_precise_stage1(); }
void dcgp_get_products(int* cond, int* dk, int* dq) {
  const dcgp::TcProducts& t = dcgp::tc_products();
  if (cond) *cond = t.cond;
  if (dk) *dk = t.dk;
  if (dq) *dq = t.dq;
}

int dcgp_view_geometry(int H, int W, int C, int f, int s, int* OH, int* OW, int* P, int* L) {
  if (H < f || W < f || f < 1 || s < 1 || C < 1) { set_error("bad view geometry"); return DCGP_ERR_ARG; }
  View v = make_view(H, W, C, f, s);
  if (OH) *OH = v.OH;
  if (OW) *OW = v.OW;
  if (P) *P = v.P;
  if (L) *L = v.L;
  return DCGP_OK;
}

int dcgp_extract_patches(const float* X, int N, int H, int W, int C, int f, int s, int layout, float* out, void* stream) {
  if (!X || !out || N < 0 || (layout != 0 && layout != 1)) { set_error("extract_patches: bad argument"); return DCGP_ERR_ARG; }
  if (H < f || W < f || f < 1 || s < 1 || C < 1) { set_error("bad view geometry"); return DCGP_ERR_ARG; }
  if (N == 0) return DCGP_OK;
  return launch_patches(X, make_view(H, W, C, f, s), N, layout, out, (cudaStream_t)stream);
}

int dcgp_kuu(const double* Z, int M, int L, double variance, double lengthscale, double jitter, double* Kuu, void* stream) {
  if (!Z || !Kuu || M < 1 || L < 1) { set_error("kuu: bad argument"); return DCGP_ERR_ARG; }
  return rbf_sym_f64(Z, M, L, variance, lengthscale, jitter, Kuu, (cudaStream_t)stream);
}

size_t dcgp_kuf_workspace_bytes(int M, int L) { return align_up((size_t)M * L * sizeof(float), 256) + 256; }

int dcgp_kuf(const float* X, int N, int H, int W, int C, int f, int s, const double* Z, int M, double variance,
             double lengthscale, int layout, int ldo, float* out, void* ws, size_t ws_bytes, void* stream) {
  if (!X || !Z || !out || N < 0 || M < 1) { set_error("kuf: bad argument"); return DCGP_ERR_ARG; }
  if (H < f || W < f || f < 1 || s < 1 || C < 1) { set_error("bad view geometry"); return DCGP_ERR_ARG; }
  if (layout != 0 && layout != 1) { set_error("kuf: bad layout"); return DCGP_ERR_ARG; }
  if (layout == 1 && ldo < M) { set_error("kuf: ldo < M"); return DCGP_ERR_ARG; }
  if (N == 0) return DCGP_OK;
  View v = make_view(H, W, C, f, s);
  if (!ws || ws_bytes < dcgp_kuf_workspace_bytes(M, v.L)) { set_error("kuf: workspace too small"); return DCGP_ERR_WORKSPACE; }
  if ((long long)N * v.HWC >= (1LL << 31) || (long long)N * v.P >= (1LL << 31)) { set_error("kuf: too many rows"); return DCGP_ERR_ARG; }
  Carver c(ws);
  float* zs = c.take<float>((size_t)M * v.L);
  cudaStream_t st = (cudaStream_t)stream;
  DCGP_TRY(launch_pack_z(Z, (long long)M * v.L, 1.0 / lengthscale, zs, st));
  return launch_kuf_simt(X, v, N, zs, M, (float)variance, (float)(1.0 / lengthscale), layout, ldo, out, st);
}

size_t dcgp_cholesky_workspace_bytes(int M) { return potrf_ws_bytes(M) + 256; }

int dcgp_cholesky(double* A, int M, void* ws, size_t ws_bytes, int* info, void* stream) {
  if (!A || !info || M < 1) { set_error("cholesky: bad argument"); return DCGP_ERR_ARG; }
  if (!ws || ws_bytes < dcgp_cholesky_workspace_bytes(M)) { set_error("cholesky: workspace too small"); return DCGP_ERR_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(info, 0, sizeof(int), st);
  Carver c(ws);
  return potrf_f64(A, M, M, c.take<double>(potrf_ws_bytes(M) / sizeof(double)), info, st);
}

// ---------------------------------------------------------------------------------------------- conditional()
struct CondWork {
  F64Work f64;
  float *W, *Wmean, *Kt, *acc;
  TcCondWork tc;
  int Mp;
  size_t bytes;
};
static CondWork carve_cond(int P, int M, int N, int R, void* ws) {
  CondWork cw;
  cw.Mp = (int)align_up(M, 64);
  cw.f64 = carve_f64(M, R, ws);
  Carver c(ws);
  c.off = cw.f64.bytes;
  const size_t T = (size_t)P * N;
  cw.W = c.take<float>((size_t)(R + 1) * cw.Mp * cw.Mp);
  cw.Wmean = c.take<float>((size_t)64 * cw.Mp);
  cw.Kt = c.take<float>(T * cw.Mp);
  cw.acc = c.take<float>(T * (R + 1));
  tc_carve_cond(cw.tc, M, cw.Mp, R, T, c.base ? c.base + align_up(c.off, 1024) : nullptr);
  c.off = align_up(c.off, 1024) + cw.tc.bytes;
  cw.bytes = align_up(c.off, 256);
  return cw;
}

size_t dcgp_conditional_workspace_bytes(int P, int M, int N, int R) { return carve_cond(P, M, N, R, nullptr).bytes; }

int dcgp_conditional(const float* Kmn, const double* Kmm, const float* Knn, const double* f, const double* q_sqrt,
                     int white, int P, int M, int N, int R, int algo, float* fmean, float* fvar, void* ws,
                     size_t ws_bytes, int* info, void* stream) {
  if (!Kmn || !Kmm || !Knn || !f || !q_sqrt || !fmean || !fvar || !info) { set_error("conditional: null argument"); return DCGP_ERR_ARG; }
  if (P < 1 || M < 1 || N < 1 || R < 1 || R > 64) { set_error("conditional: bad shape"); return DCGP_ERR_ARG; }
  CondWork cw = carve_cond(P, M, N, R, ws);
  if (!ws || ws_bytes < cw.bytes) { set_error("conditional: workspace too small (%zu < %zu)", ws_bytes, cw.bytes); return DCGP_ERR_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(info, 0, sizeof(int), st);
  cudaMemcpyAsync(cw.f64.Kuu, Kmm, (size_t)M * M * sizeof(double), cudaMemcpyDeviceToDevice, st);
  DCGP_TRY(build_operands(cw.f64, white, f, q_sqrt, info, st));
  const int T = P * N;
  DCGP_TRY(launch_pmn_to_tm(Kmn, P, M, N, cw.Mp, cw.Kt, st));
  if (algo == DCGP_ALGO_TC) {
    DCGP_TRY(tc_pack_operands(cw.tc.prep, cw.f64.Linv, cw.f64.Mq, cw.f64.Wr, cw.f64.beta, M, cw.Mp, R, st));
    DCGP_TRY(tc_split_rows(cw.Kt, T, cw.Mp, cw.tc, st));
    DCGP_TRY(tc_cond(cw.tc.prep, cw.tc, T, cw.Mp, R, cw.acc, fmean, st));
  } else {
    DCGP_TRY(launch_pack_w(cw.f64.Linv, cw.f64.Mq, cw.f64.Wr, M, cw.Mp, R, cw.W, st));
    DCGP_TRY(launch_pack_wmean(cw.f64.beta, M, cw.Mp, R, 64, cw.Wmean, st));
    DCGP_TRY(launch_cond_simt(cw.Kt, T, cw.Mp, cw.Mp, cw.W, cw.Wmean, R, cw.acc, fmean, st));
  }
  return launch_finalize_ref_layout(cw.acc, Knn, P, N, R, fvar, st);
}

// ---------------------------------------------------------------------------------------------- layer prepare
size_t dcgp_prepare_bytes(const dcgp_layer_desc* d) {
  if (check_desc(d)) return 0;
  return carve_prep(d, nullptr).bytes;
}
size_t dcgp_prepare_workspace_bytes(const dcgp_layer_desc* d) {
  if (check_desc(d)) return 0;
  return carve_f64(d->M, d->R, nullptr).bytes;
}

int dcgp_layer_prepare_ev(const dcgp_layer_desc* d, const double* Z, const double* Z_prior, const double* q_mu,
                          const double* q_sqrt, int algo, void* prep_buf, double* kl, void* ws, size_t ws_bytes, int* info,
                          void* fwd_ready_event, void* stream) {
  return dcgp_layer_prepare_hyp(d, Z, Z_prior, q_mu, q_sqrt, algo, prep_buf, kl, ws, ws_bytes, info, fwd_ready_event, nullptr,
                                stream);
}

int dcgp_layer_prepare_hyp(const dcgp_layer_desc* d, const double* Z, const double* Z_prior, const double* q_mu,
                           const double* q_sqrt, int algo, void* prep_buf, double* kl, void* ws, size_t ws_bytes, int* info,
                           void* fwd_ready_event, const double* hyp, void* stream) {
  DCGP_TRY(check_desc(d));
  if (!Z || !q_mu || !q_sqrt || !prep_buf || !kl || !info) { set_error("layer_prepare: null argument"); return DCGP_ERR_ARG; }
  F64Work w = carve_f64(d->M, d->R, ws);
  if (!ws || ws_bytes < w.bytes) { set_error("layer_prepare: workspace too small (%zu < %zu)", ws_bytes, w.bytes); return DCGP_ERR_WORKSPACE; }
  Prep p = carve_prep(d, prep_buf);
  cudaStream_t st = (cudaStream_t)stream;
  const int M = d->M, R = d->R, L = p.L;
  cudaMemsetAsync(info, 0, sizeof(int), st);
  DCGP_TRY(rbf_sym_f64(Z, M, L, d->variance, d->lengthscale, d->jitter, w.Kuu, st, hyp));   // layers.py:18-21 / DS/layers.py:184
  DCGP_TRY(launch_pack_z(Z, (long long)M * L, 1.0 / d->lengthscale, p.zs, st, hyp));
  // KL prior.  ConvLayer: Kuu at the *initial* Z with the live kernel hyper-parameters (layers.py:149-150, SURVEY
  // App. C3); SVGP_Layer: the current Ku (DS/layers.py:242-256).
  const bool own_prior = (d->kind == DCGP_LAYER_CONV) && Z_prior && Z_prior != Z && !d->white;
  GInfo gi;
  const double* Lp = w.Kuu;        // Cholesky factor of the prior covariance and its inverse
  const double* Lpinv = w.Linv;
  double* Kp = w.Wr;               // scratch that is free at this point on both paths
  AuxFork* ax = nullptr;
  if (own_prior) {                 // forked: runs concurrently with the Kuu chain below
    ax = &aux_for(st);
    if (!ax->ok) return DCGP_ERR_CUDA;
    cudaEventRecord(ax->fork, st);
    cudaStreamWaitEvent(ax->aux, ax->fork, 0);
    DCGP_TRY(rbf_sym_f64(Z_prior, M, L, d->variance, d->lengthscale, d->jitter, Kp, ax->aux, hyp));
    DCGP_TRY(potrf_f64(Kp, M, M, w.invDp, info, ax->aux));
    DCGP_TRY(trtri_f64(Kp, M, M, w.invDp, w.Lpinv, w.trws2, ax->aux));
    cudaEventRecord(ax->join, ax->aux);
    Lp = Kp;
    Lpinv = w.Lpinv;
  }
  if (algo == DCGP_ALGO_TC) {
    // Forward operands first; the caller's event marks the point from which dcgp_layer_apply may run -- the KL (which joins
    // the prior chain) and the backward operands follow on the same stream, off the forward's critical path.
    // chained (default): a = Lm^-1 k, then G_r = C_r^T a: needs Lm^-1, C_r^T and alpha only.
    const int chained = tc_forward_chained() ? 1 : 0;
    DCGP_TRY(factor_only(w, info, st));
    const double* alpha = q_mu;          // whitened: mean = a^T q_mu
    if (!d->white) {                     // alpha = Lm^-1 q_mu
      GemmF64 g{};
      g.m = M; g.n = R; g.k = M;
      g.A = w.Linv; g.lda = w.Mq; g.lowerA = 1;
      g.B = q_mu; g.ldb = R;
      g.C = w.alpha; g.ldc = R; g.alpha = 1.0; g.batch = 1;
      DCGP_TRY(gemm_f64(g, st));
      alpha = w.alpha;
    }
    if (chained) {
      DCGP_TRY(tc_build_operands(p.tc, w.Linv, w.Mq, nullptr, 0, d->white ? 1 : 0, nullptr, w.Mq, q_sqrt, w.beta, w.sc + 1,
                                 nullptr, 1, 1, alpha, st));
      DCGP_TRY(tc_pack_z(p.tc, Z, M, L, 1.0 / d->lengthscale, st, hyp));
      if (fwd_ready_event) cudaEventRecord((cudaEvent_t)fwd_ready_event, st);
      DCGP_TRY(g_and_beta(w, d->white, q_mu, &gi, st));
    } else {                             // dense single-stack forward (diagnostic): W_r = L_r^T G, beta
      DCGP_TRY(g_and_beta(w, d->white, q_mu, &gi, st));
      DCGP_TRY(tc_build_operands(p.tc, w.Linv, w.Mq, gi.G, gi.ldg, d->white ? 1 : 0, nullptr, w.Mq, q_sqrt, w.beta, w.sc + 1,
                                 nullptr, 1, 0, nullptr, st));
      DCGP_TRY(tc_pack_z(p.tc, Z, M, L, 1.0 / d->lengthscale, st, hyp));
      if (fwd_ready_event) cudaEventRecord((cudaEvent_t)fwd_ready_event, st);
    }
    if (d->white) {   // Kuu^-1 is read by the host's chain rule even when G = Lm^-1
      GemmF64 g{};
      g.m = g.n = g.k = M;
      g.A = w.Linv; g.lda = w.Mq; g.transA = 1; g.lowerA = 1;
      g.B = w.Linv; g.ldb = w.Mq; g.lowerB = 1;
      g.C = w.Kinv; g.ldc = M; g.alpha = 1.0; g.batch = 1;
      DCGP_TRY(gemm_f64(g, st));
    }
    if (own_prior) cudaStreamWaitEvent(st, ax->join, 0);
    DCGP_TRY(tc_build_operands(p.tc, w.Linv, w.Mq, gi.G, gi.ldg, d->white ? 1 : 0, d->white ? nullptr : Lpinv, w.Mq, q_sqrt,
                               w.beta, w.sc + 1, w.Kinv, 2, chained, alpha, st));
    return kl_terms(w, d->white, Lp, M, Lpinv, w.Mq, q_mu, q_sqrt, kl, st, /*have_trace=*/!d->white);
  }
  DCGP_TRY(factor_and_G(w, d->white, q_mu, info, &gi, st));
  if (own_prior) cudaStreamWaitEvent(st, ax->join, 0);
  if (own_prior) {   // the fp64 path needs w.Wr for W_r first: move the prior factor out of the way
    cudaMemcpyAsync(w.Kinv == gi.G ? w.tmpK : w.Kinv, Kp, (size_t)M * M * sizeof(double), cudaMemcpyDeviceToDevice, st);
    Lp = (w.Kinv == gi.G) ? w.tmpK : w.Kinv;
  }
  DCGP_TRY(wr_f64(w, d->white, gi, q_sqrt, st));
  DCGP_TRY(launch_pack_w(w.Linv, w.Mq, w.Wr, M, p.Mp, R, p.W, st));
  DCGP_TRY(launch_pack_wmean(w.beta, M, p.Mp, R, p.RP, p.Wmean, st));
  if (fwd_ready_event) cudaEventRecord((cudaEvent_t)fwd_ready_event, st);
  return kl_terms(w, d->white, Lp, M, Lpinv, w.Mq, q_mu, q_sqrt, kl, st);
}

int dcgp_layer_prepare(const dcgp_layer_desc* d, const double* Z, const double* Z_prior, const double* q_mu,
                       const double* q_sqrt, int algo, void* prep_buf, double* kl, void* ws, size_t ws_bytes, int* info,
                       void* stream) {
  return dcgp_layer_prepare_ev(d, Z, Z_prior, q_mu, q_sqrt, algo, prep_buf, kl, ws, ws_bytes, info, nullptr, stream);
}

int dcgp_prepare_workspace_layout(const dcgp_layer_desc* d, size_t* off_kinv, size_t* off_linv, size_t* off_lpinv, int* ld_inv) {
  if (check_desc(d)) return DCGP_ERR_ARG;
  char* const fake = (char*)(uintptr_t)(1u << 20);       // carve against a fake (256-B aligned) base to read the offsets back
  F64Work w = carve_f64(d->M, d->R, fake);
  if (off_kinv) *off_kinv = (size_t)((char*)w.Kinv - fake);
  if (off_linv) *off_linv = (size_t)((char*)w.Linv - fake);
  if (off_lpinv) *off_lpinv = (size_t)((char*)w.Lpinv - fake);
  if (ld_inv) *ld_inv = w.Mq;
  return DCGP_OK;
}

int dcgp_prepare_workspace_layout2(const dcgp_layer_desc* d, size_t* off_kinv, size_t* off_linv, size_t* off_lpinv, size_t* off_lm,
                                   int* ld_inv) {
  if (check_desc(d)) return DCGP_ERR_ARG;
  char* const fake = (char*)(uintptr_t)(1u << 20);
  F64Work w = carve_f64(d->M, d->R, fake);
  if (off_lm) *off_lm = (size_t)((char*)w.Kuu - fake);    // Kuu is factored in place: its lower triangle is Lm (ld M)
  return dcgp_prepare_workspace_layout(d, off_kinv, off_linv, off_lpinv, ld_inv);
}

int dcgp_prepare_layout2(const dcgp_layer_desc* d, size_t* off_c32, size_t* off_s32, int* ld) {
  if (check_desc(d)) return DCGP_ERR_ARG;
  char* const fake = (char*)(uintptr_t)(1u << 20);
  Prep p = carve_prep(d, fake);
  if (off_c32) *off_c32 = (size_t)((char*)p.tc.Br32 - fake);
  if (off_s32) *off_s32 = (size_t)((char*)p.tc.Qr32 - fake);
  if (ld) *ld = p.Mp;
  return DCGP_OK;
}

int dcgp_prepare_layout(const dcgp_layer_desc* d, size_t* off_b32, int* ld_b) {
  if (check_desc(d)) return DCGP_ERR_ARG;
  char* const fake = (char*)(uintptr_t)(1u << 20);
  Prep p = carve_prep(d, fake);
  if (off_b32) *off_b32 = (size_t)((char*)p.tc.Br32 - fake);
  if (ld_b) *ld_b = p.Mp;
  return DCGP_OK;
}

// ---------------------------------------------------------------------------------------------- layer apply
struct ApplyWork {
  float *Kt, *acc, *mean_t, *Kzx, *kdiag;
  TcApplyWork tc;
  size_t bytes;
};
static ApplyWork carve_apply(const dcgp_layer_desc* d, int n_rows, void* ws) {
  ApplyWork a;
  View v = make_view(d->H, d->W, d->C, d->f, d->s);
  const size_t Mp = align_up(d->M, 64);
  const size_t Tk = (size_t)n_rows * v.P;                                   // rows of the patch-level kernel matrix
  const size_t T = (d->kind == DCGP_LAYER_CONV) ? Tk : (size_t)n_rows;      // columns the conditional runs over
  Carver c(ws);
  a.Kt = c.take<float>(Tk * Mp);
  a.acc = c.take<float>(T * (d->R + 1));
  a.mean_t = c.take<float>(T * d->R);
  a.Kzx = (d->kind == DCGP_LAYER_CONV) ? nullptr : c.take<float>((size_t)n_rows * Mp);
  a.kdiag = (d->kind == DCGP_LAYER_CONV) ? nullptr : c.take<float>((size_t)n_rows);
  tc_carve_apply(a.tc, d->kind, d->M, (int)Mp, d->R, v.L, Tk, T, c.base ? c.base + align_up(c.off, 1024) : nullptr);
  c.off = align_up(c.off, 1024) + a.tc.bytes;
  a.bytes = align_up(c.off, 256);
  return a;
}

size_t dcgp_apply_workspace_bytes(const dcgp_layer_desc* d, int n_rows, int n_rep) {
  (void)n_rep;
  if (check_desc(d) || n_rows < 1) return 0;
  return carve_apply(d, n_rows, nullptr).bytes;
}

int dcgp_layer_apply(const dcgp_layer_desc* d, const void* prep_buf, const double* patch_weights, const float* X,
                     int n_rows, int n_rep, const float* z, int algo, float* mean, float* var, float* sample, void* ws,
                     size_t ws_bytes, void* stream) {
  DCGP_TRY(check_desc(d));
  if (!prep_buf || !X || !mean || !var || n_rows < 1 || n_rep < 1 || (z && !sample)) { set_error("layer_apply: bad argument"); return DCGP_ERR_ARG; }
  View v = make_view(d->H, d->W, d->C, d->f, d->s);
  if ((long long)n_rows * v.HWC >= (1LL << 31) || (long long)n_rows * v.P >= (1LL << 31)) { set_error("layer_apply: too many rows"); return DCGP_ERR_ARG; }
  ApplyWork a = carve_apply(d, n_rows, ws);
  if (!ws || ws_bytes < a.bytes) { set_error("layer_apply: workspace too small (%zu < %zu)", ws_bytes, a.bytes); return DCGP_ERR_WORKSPACE; }
  Prep p = carve_prep(d, const_cast<void*>(prep_buf));
  cudaStream_t st = (cudaStream_t)stream;
  const float variance = (float)d->variance, inv_ls = (float)(1.0 / d->lengthscale);
  const int Tk = n_rows * v.P;
  const bool conv = d->kind == DCGP_LAYER_CONV;
  const int T = conv ? Tk : n_rows;
  if (algo == DCGP_ALGO_TC) {
    DCGP_TRY(tc_layer_apply(d, v, p.tc, a.tc, p.zs, a.Kt, patch_weights, X, n_rows, a.Kzx, a.acc, a.mean_t, st));
  } else {
    DCGP_TRY(launch_kuf_simt(X, v, n_rows, p.zs, p.M, variance, inv_ls, 1, p.Mp, a.Kt, st));       // layers.py:112 / kernels.py:123
    const float* Kcols = a.Kt;
    if (!conv) {
      DCGP_TRY(launch_patch_mean(a.Kt, n_rows, v.P, p.Mp, p.M, patch_weights, 0, p.Mp, a.Kzx, st)); // kernels.py:127-133
      Kcols = a.Kzx;
    }
    DCGP_TRY(launch_cond_simt(Kcols, T, p.Mp, p.Mp, p.W, p.Wmean, p.R, a.acc, a.mean_t, st));       // conditionals.py:31-65
  }
  if (!conv) DCGP_TRY(launch_kdiag(X, v, n_rows, patch_weights, variance, inv_ls * inv_ls, a.kdiag, st));  // kernels.py:106-115
  return launch_finalize(a.acc, a.mean_t, T, p.R, variance, conv ? nullptr : a.kdiag, n_rep, z, (float)d->jitter, mean, var,
                         sample, st);
}

// ---------------------------------------------------------------------------------------------- layer backward
static void bwd_dims(const dcgp_layer_desc* d, int n_rows, View& v, size_t& Tk, size_t& T, int& Mp) {
  v = make_view(d->H, d->W, d->C, d->f, d->s);
  Mp = (int)align_up(d->M, 64);
  Tk = (size_t)n_rows * v.P;
  T = (d->kind == DCGP_LAYER_CONV) ? Tk : (size_t)n_rows;
}

size_t dcgp_backward_workspace_bytes(const dcgp_layer_desc* d, int n_rows, int n_rep) {
  (void)n_rep;
  if (check_desc(d) || n_rows < 1) return 0;
  View v; size_t Tk, T; int Mp;
  bwd_dims(d, n_rows, v, Tk, T, Mp);
  TcBwdWork b;
  tc_carve_bwd(b, d->kind, d->M, Mp, d->R, v.L, Tk, T, v.P, nullptr);
  return b.bytes + 1024;
}

int dcgp_layer_backward_phases(const dcgp_layer_desc* d, const void* prep_buf, const void* apply_ws, const double* Z,
                        const double* patch_weights, const float* X, int n_rows, int n_rep, const float* g_mean,
                        const float* g_var, float* gX, double* gQB, double* gZ, double* gscal, double* gw, void* ws,
                        size_t ws_bytes, int phases, void* stream) {
  DCGP_TRY(check_desc(d));
  if (!prep_buf || !apply_ws || !Z || !X || !g_mean || !g_var || !gQB || !gZ || !gscal || n_rows < 1 || n_rep < 1) {
    set_error("layer_backward: bad argument");
    return DCGP_ERR_ARG;
  }
  if (phases < 1 || phases > 3) { set_error("layer_backward: phases must be 1, 2 or 3"); return DCGP_ERR_ARG; }
  if (d->kind == DCGP_LAYER_SVGP_CONV && !gw) { set_error("layer_backward: gw required for the ConvKernel layer"); return DCGP_ERR_ARG; }
  View v; size_t Tk, T; int Mp;
  bwd_dims(d, n_rows, v, Tk, T, Mp);
  TcBwdWork b;
  void* wsa = (void*)align_up((size_t)ws, 1024);
  tc_carve_bwd(b, d->kind, d->M, Mp, d->R, v.L, Tk, T, v.P, wsa);
  if (!ws || ws_bytes < b.bytes + 1024) { set_error("layer_backward: workspace too small (%zu < %zu)", ws_bytes, b.bytes + 1024); return DCGP_ERR_WORKSPACE; }
  Prep p = carve_prep(d, const_cast<void*>(prep_buf));
  ApplyWork a = carve_apply(d, n_rows, const_cast<void*>(apply_ws));
  return tc_layer_backward(d, v, p.tc, a.tc, b, Z, patch_weights, X, n_rows, n_rep, g_mean, g_var, gX, gQB, gZ, gscal, gw,
                           phases, (cudaStream_t)stream);
}

int dcgp_layer_backward(const dcgp_layer_desc* d, const void* prep_buf, const void* apply_ws, const double* Z,
                        const double* patch_weights, const float* X, int n_rows, int n_rep, const float* g_mean,
                        const float* g_var, float* gX, double* gQB, double* gZ, double* gscal, double* gw, void* ws,
                        size_t ws_bytes, void* stream) {
  return dcgp_layer_backward_phases(d, prep_buf, apply_ws, Z, patch_weights, X, n_rows, n_rep, g_mean, g_var, gX, gQB, gZ, gscal,
                                    gw, ws, ws_bytes, 3, stream);
}

size_t dcgp_chain_rule_workspace_bytes(const dcgp_layer_desc* d) {
  if (check_desc(d)) return 0;
  return chain_rule_workspace_bytes(d->M, d->R, d->f * d->f * d->C) + 256;
}

int dcgp_layer_chain_rule(const dcgp_layer_desc* d, const void* prep_buf, const void* prepare_ws, const double* Z,
                          const double* Z_prior, const double* q_mu, const double* q_sqrt, const double* hyp, const double* gQB,
                          const double* gZ_direct, const double* gscal, double kl_weight, int parts, double* gZ, double* ghyp,
                          double* g_qmu, double* g_qsqrt, void* ws, size_t ws_bytes, void* stream) {
  DCGP_TRY(check_desc(d));
  if (!prep_buf || !prepare_ws || !Z || !q_mu || !q_sqrt || parts < 1 || parts > 3) { set_error("chain_rule: bad argument"); return DCGP_ERR_ARG; }
  if ((parts & 2) && (!gQB || !gZ_direct || !gscal || !gZ || !ghyp || !g_qmu || !g_qsqrt)) { set_error("chain_rule: null gradient buffer"); return DCGP_ERR_ARG; }
  if (!ws || ws_bytes < dcgp_chain_rule_workspace_bytes(d)) { set_error("chain_rule: workspace too small"); return DCGP_ERR_WORKSPACE; }
  F64Work w = carve_f64(d->M, d->R, const_cast<void*>(prepare_ws));
  Prep p = carve_prep(d, const_cast<void*>(prep_buf));
  ChainInputs in;
  in.Kinv = w.Kinv; in.Li = w.Linv; in.Lpinv = w.Lpinv; in.Lm = w.Kuu; in.alpha = w.alpha; in.ldi = w.Mq;
  in.C32 = p.tc.Br32; in.S32 = p.tc.Qr32; in.Ct32 = p.tc.Wr32;
  // (a ConvLayer whose KL prior is the current Z shares the factor of Kuu: Lp^-1 = Lm^-1)
  const bool own_prior = (d->kind == DCGP_LAYER_CONV) && Z_prior && Z_prior != Z && !d->white;
  if (!own_prior) in.Lpinv = w.Linv;
  return chain_rule(d, in, Z, own_prior ? Z_prior : nullptr, q_mu, q_sqrt, hyp, gQB, gZ_direct, gscal, kl_weight, parts, gZ, ghyp,
                    g_qmu, g_qsqrt, (void*)align_up((size_t)ws, 256), (cudaStream_t)stream);
}

size_t dcgp_bgemm_workspace_bytes(int batch, int m, int n, int k) {
  if (batch < 1 || m < 1 || n < 1 || k < 1) return 0;
  return tc_bgemm_workspace_bytes(batch, m, n, k);
}

int dcgp_bgemm_nt(const float* A, const float* B, float* C, int batch, int m, int n, int k, long long a_bstride,
                  long long b_bstride, void* ws, size_t ws_bytes, void* stream) {
  if (!A || !B || !C || batch < 1 || m < 1 || n < 1 || k < 1 || a_bstride < 0 || b_bstride < 0 || (n % 4) != 0) {
    set_error("bgemm_nt: bad argument (n must be a multiple of 4)");
    return DCGP_ERR_ARG;
  }
  if (!ws || ws_bytes < dcgp_bgemm_workspace_bytes(batch, m, n, k)) { set_error("bgemm_nt: workspace too small"); return DCGP_ERR_WORKSPACE; }
  return tc_bgemm_nt(A, B, C, batch, m, n, k, a_bstride, b_bstride, ws, (cudaStream_t)stream);
}

int dcgp_multiclass_varexp_grad(const float* Fmu, const float* Fvar, const int32_t* Y, int S, int N, int K, double epsilon,
                                double coef, float* gmu, float* gvar, void* stream) {
  if (!Fmu || !Fvar || !Y || !gmu || !gvar || S < 1 || N < 1) { set_error("varexp_grad: bad argument"); return DCGP_ERR_ARG; }
  return launch_varexp_grad(Fmu, Fvar, Y, S, N, K, epsilon, coef, gmu, gvar, (cudaStream_t)stream);
}

int dcgp_sample_backward(const float* gF, const float* z, const float* var, size_t n, double jitter, float* g_mean,
                         float* g_var, void* stream) {
  if (!gF || !z || !var || !g_mean || !g_var) { set_error("sample_backward: null argument"); return DCGP_ERR_ARG; }
  if (n == 0) return DCGP_OK;
  return launch_sample_backward(gF, z, var, n, (float)jitter, g_mean, g_var, (cudaStream_t)stream);
}

int dcgp_adam(double* param, const double* grad, double* m, double* v, size_t n, double lr, double beta1, double beta2,
              double eps, int step, int maximize, void* stream) {
  if (!param || !grad || !m || !v || step < 1) { set_error("adam: bad argument"); return DCGP_ERR_ARG; }
  if (n == 0) return DCGP_OK;
  return launch_adam(param, grad, m, v, n, lr, beta1, beta2, eps, step, maximize, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------- ConvKernel API mirror
size_t dcgp_convkernel_kzx_workspace_bytes(const dcgp_layer_desc* d, int N) {
  if (check_desc(d) || N < 1) return 0;
  View v = make_view(d->H, d->W, d->C, d->f, d->s);
  const size_t Mp = align_up(d->M, 64);
  return align_up((size_t)N * v.P * Mp * sizeof(float), 256) + align_up((size_t)d->M * v.L * sizeof(float), 256) + 256;
}

int dcgp_convkernel_kzx(const dcgp_layer_desc* d, const double* Z, const double* patch_weights, const float* X, int N,
                        float* out, void* ws, size_t ws_bytes, void* stream) {
  DCGP_TRY(check_desc(d));
  if (!Z || !X || !out || N < 1) { set_error("kzx: bad argument"); return DCGP_ERR_ARG; }
  if (!ws || ws_bytes < dcgp_convkernel_kzx_workspace_bytes(d, N)) { set_error("kzx: workspace too small"); return DCGP_ERR_WORKSPACE; }
  View v = make_view(d->H, d->W, d->C, d->f, d->s);
  const int Mp = (int)align_up(d->M, 64);
  Carver c(ws);
  float* Kt = c.take<float>((size_t)N * v.P * Mp);
  float* zs = c.take<float>((size_t)d->M * v.L);
  cudaStream_t st = (cudaStream_t)stream;
  DCGP_TRY(launch_pack_z(Z, (long long)d->M * v.L, 1.0 / d->lengthscale, zs, st));
  DCGP_TRY(launch_kuf_simt(X, v, N, zs, d->M, (float)d->variance, (float)(1.0 / d->lengthscale), 1, Mp, Kt, st));
  return launch_patch_mean(Kt, N, v.P, Mp, d->M, patch_weights, 1, 0, out, st);
}

int dcgp_convkernel_kdiag(const dcgp_layer_desc* d, const double* patch_weights, const float* X, int N, float* out,
                          void* stream) {
  DCGP_TRY(check_desc(d));
  if (!X || !out || N < 1) { set_error("kdiag: bad argument"); return DCGP_ERR_ARG; }
  View v = make_view(d->H, d->W, d->C, d->f, d->s);
  const float inv_ls = (float)(1.0 / d->lengthscale);
  return launch_kdiag(X, v, N, patch_weights, (float)d->variance, inv_ls * inv_ls, out, (cudaStream_t)stream);
}

int dcgp_randn(float* z, int S, int n_local, int D, long long n_global, long long n0, unsigned long long seed,
               unsigned long long step, int layer, void* stream) {
  if (!z || S < 1 || n_local < 0 || D < 1 || n_global < n_local || n0 < 0 || n0 + n_local > n_global) {
    set_error("randn: bad argument");
    return DCGP_ERR_ARG;
  }
  return launch_randn(z, S, n_local, D, n_global, n0, seed, step, layer, (cudaStream_t)stream);
}

int dcgp_reparameterize(const float* mean, const float* var, const float* z, size_t n, double jitter, float* out,
                        void* stream) {
  if (!mean || !var || !z || !out) { set_error("reparameterize: null argument"); return DCGP_ERR_ARG; }
  if (n == 0) return DCGP_OK;
  return launch_reparam(mean, var, z, n, (float)jitter, out, (cudaStream_t)stream);
}

int dcgp_multiclass_varexp(const float* Fmu, const float* Fvar, const int32_t* Y, int S, int N, int K, double epsilon,
                           double* varexp, double* sum, void* stream) {
  if (!Fmu || !Fvar || !Y || !varexp || !sum || S < 1 || N < 1) { set_error("varexp: bad argument"); return DCGP_ERR_ARG; }
  return launch_varexp(Fmu, Fvar, Y, S, N, K, epsilon, varexp, sum, (cudaStream_t)stream);
}

int dcgp_multiclass_predict(const float* Fmu, const float* Fvar, const int32_t* Y, int S, int N, int K, double epsilon,
                            double* pmean, double* pvar, double* logdens, void* stream) {
  if (!Fmu || !Fvar || S < 1 || N < 1 || (!pmean && !pvar && !logdens) || (logdens && !Y)) {
    set_error("multiclass_predict: bad argument");
    return DCGP_ERR_ARG;
  }
  return launch_multiclass_predict(Fmu, Fvar, Y, S, N, K, epsilon, pmean, pvar, logdens, (cudaStream_t)stream);
}

int dcgp_elbo(const double* sum_varexp, int S, double num_data, double n_global, const double* kls, int n_layers,
              double* elbo, void* stream) {
  if (!sum_varexp || !kls || !elbo || S < 1 || n_layers < 1 || !(n_global > 0)) { set_error("elbo: bad argument"); return DCGP_ERR_ARG; }
  return launch_elbo(sum_varexp, S, num_data / n_global, kls, n_layers, elbo, (cudaStream_t)stream);
}

}  // extern "C"
