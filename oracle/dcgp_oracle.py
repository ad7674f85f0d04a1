"""TEST INFRASTRUCTURE ONLY -- float64 NumPy/SciPy restatement of the reference's conv-GP
doubly-stochastic forward pass (the hot path of BASELINE.json:north_star).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product (deepcgp_b200) never does and fails loudly without its CUDA library.

Pinning status
  * Everything the reference itself owns (views.py, layers.py, conditionals.py, kernels.py,
    DS/layers.py, DS/dgp.py, DS/utils.py) is pinned: tests/golden/*.npz were produced by executing
    those *unmodified source files* from /root/reference on top of oracle/refshim (numpy stand-ins for
    the TensorFlow ops and GPflow classes they call) -- see tests/golden/make_golden.py -- and
    tests/test_oracle_golden.py checks this restatement against them to 1e-12.
  * GPflow-1.2.0-owned formulas (RBF.K, gauss_kl, MultiClass/RobustMax, LowerTriangular) and the
    TensorFlow op semantics (extract_image_patches ordering) are third-party code absent from
    /root/reference (requirements.txt:1-2); they are restated from their published definitions in
    both the shim and here: that part is "parity unpinned" (SURVEY.md 8c, App. A.5).

All citations are relative to /root/reference; DS/ = submodules/Doubly-Stochastic-DGP/doubly_stochastic_dgp/.
"""
import numpy as np
import scipy.linalg as sla
import scipy.special as ssp

JITTER = 1e-3  # gpflowrc:11 (numerics.jitter_level), float64 per gpflowrc:7


# ------------------------------------------------------------------ a1/a2: views.py
def out_image_size(H, W, f, s):
    """views.py:65-68."""
    return (H - f) // s + 1, (W - f) // s + 1


def extract_patches(X_nhwc, f, s):
    """views.py:46-54 (+ :32-38 tf.extract_image_patches VALID): [N,H,W,C] -> [N,P,L];
    patch vector order (dy, dx, c) with c fastest, p = oy*OW + ox."""
    N, H, W, C = X_nhwc.shape
    OH, OW = out_image_size(H, W, f, s)
    out = np.empty((N, OH, OW, f, f, C), dtype=X_nhwc.dtype)
    for dy in range(f):
        for dx in range(f):
            out[:, :, :, dy, dx, :] = X_nhwc[:, dy:dy + (OH - 1) * s + 1:s, dx:dx + (OW - 1) * s + 1:s, :]
    return out.reshape(N, OH * OW, f * f * C)


def extract_patches_PNL(X_nhwc, f, s):
    """views.py:40-44: the same tensor transposed to [P,N,L]."""
    return np.ascontiguousarray(extract_patches(X_nhwc, f, s).transpose(1, 0, 2))


# ------------------------------------------------------------------ GPflow RBF (App. A.2 / A.5)
def rbf_K(X, X2, variance, lengthscale):
    """GPflow-1.2.0 RBF.K via Stationary.square_dist: expansion form, no clamp."""
    X = X / lengthscale
    Xs = np.sum(X * X, axis=1)
    if X2 is None:
        d = -2.0 * X @ X.T + Xs[:, None] + Xs[None, :]
    else:
        X2 = X2 / lengthscale
        X2s = np.sum(X2 * X2, axis=1)
        d = -2.0 * X @ X2.T + Xs[:, None] + X2s[None, :]
    return variance * np.exp(-0.5 * d)


# ------------------------------------------------------------------ a3/a4/a4': layers.py:12-50
def mo_Kuu(Z, variance, lengthscale, jitter=JITTER):
    """layers.py:18-21."""
    return rbf_K(Z, None, variance, lengthscale) + jitter * np.eye(Z.shape[0])


def mo_Kuf(Z, PNL, variance, lengthscale):
    """layers.py:23-32: per patch position p, RBF.K(Z, X_p) -> [P,M,N]."""
    return np.stack([rbf_K(Z, PNL[p], variance, lengthscale) for p in range(PNL.shape[0])])


def mo_Kdiag(PNL, variance):
    """layers.py:43-50: RBF.Kdiag = variance, [P,N]."""
    return np.full(PNL.shape[:2], float(variance))


# ------------------------------------------------------------------ a5: conditionals.py:6-67 (diag case)
def conditional(Kmn, Kmm, Knn, f, q_sqrt=None, white=False):
    """Literal restatement, materialising A [P,M,N] and LTA [R,M,P,N] exactly as the TF graph does.
    Returns fmean [N,P,R], fvar [R,P,N]."""
    P, M, N = Kmn.shape
    R = f.shape[1]
    Lm = np.linalg.cholesky(Kmm)                                           # :29
    A = np.stack([sla.solve_triangular(Lm, Kmn[p], lower=True) for p in range(P)])  # :31-33
    fvar = Knn - np.sum(A * A, axis=1)                                     # :40
    fvar = np.tile(fvar[None], (R, 1, 1))                                  # :41
    if not white:                                                          # :44-47
        A = np.stack([sla.solve_triangular(Lm.T, A[p], lower=False) for p in range(P)])
    fmean = np.tensordot(A, f, [[1], [0]]).transpose(1, 0, 2)              # :50-51
    if q_sqrt is not None:
        L = np.tril(q_sqrt)                                                # :55
        LTA = np.tensordot(L, A, [[1], [1]])                               # :58  [R,M,P,N]
        fvar = fvar + np.sum(LTA * LTA, axis=1)                            # :65
    return fmean, fvar


def conditional_single_solve(Kmn, Kmm, Knn, f, q_sqrt, white=False):
    """algebraically identical device form (one solve, no [R,M,P,N] tensor).
    Used for large-size CPU baselines and as a cross-check of `conditional` (<= 1e-12)."""
    P, M, N = Kmn.shape
    R = f.shape[1]
    Lm = np.linalg.cholesky(Kmm)
    L = np.tril(q_sqrt)
    if white:
        alpha, C = f, L
    else:
        alpha = sla.solve_triangular(Lm, f, lower=True)
        C = np.stack([sla.solve_triangular(Lm, L[r], lower=True) for r in range(R)])
    a = sla.solve_triangular(Lm, Kmn.transpose(1, 0, 2).reshape(M, P * N), lower=True)  # [M, P*N]
    base = Knn.reshape(-1) - np.sum(a * a, axis=0)
    fmean = (a.T @ alpha).reshape(P, N, R).transpose(1, 0, 2)
    fvar = np.stack([base + np.sum((C[r].T @ a) ** 2, axis=0) for r in range(R)]).reshape(R, P, N)
    return fmean, fvar


# ------------------------------------------------------------------ a6: layers.py:96-135
def convlayer_conditional_ND(X_ND, lay, jitter=JITTER):
    """ConvLayer.conditional_ND (full_cov=False, Zero mean function).
    lay: dict(H,W,C,f,s,Z,variance,lengthscale,q_mu[M,R],q_sqrt[R,M,M],white).
    Returns mean, var of shape [N, P*R], flat index p*R + r."""
    N = X_ND.shape[0]
    X = X_ND.reshape(N, lay["H"], lay["W"], lay["C"])                      # :108
    PNL = extract_patches_PNL(X, lay["f"], lay["s"])                       # :109
    Kuu = mo_Kuu(lay["Z"], lay["variance"], lay["lengthscale"], jitter)   # :111
    Kuf = mo_Kuf(lay["Z"], PNL, lay["variance"], lay["lengthscale"])      # :112
    Knn = mo_Kdiag(PNL, lay["variance"])                                   # :117
    fmean, fvar = conditional(Kuf, Kuu, Knn, lay["q_mu"], q_sqrt=lay["q_sqrt"], white=lay["white"])
    P, R = PNL.shape[0], lay["q_mu"].shape[1]
    var = fvar.transpose(2, 1, 0).reshape(N, P * R)                        # :128-129
    mean = fmean.reshape(N, P * R)                                         # :131
    return mean, var                                                       # + Zero() :133-134


def convlayer_conditional_ND_fast(X_ND, lay, jitter=JITTER):
    """Same outputs through `conditional_single_solve` (for sizes where LTA would not fit)."""
    N = X_ND.shape[0]
    X = X_ND.reshape(N, lay["H"], lay["W"], lay["C"])
    PNL = extract_patches_PNL(X, lay["f"], lay["s"])
    Kuu = mo_Kuu(lay["Z"], lay["variance"], lay["lengthscale"], jitter)
    Kuf = mo_Kuf(lay["Z"], PNL, lay["variance"], lay["lengthscale"])
    Knn = mo_Kdiag(PNL, lay["variance"])
    fmean, fvar = conditional_single_solve(Kuf, Kuu, Knn, lay["q_mu"], lay["q_sqrt"], white=lay["white"])
    P, R = PNL.shape[0], lay["q_mu"].shape[1]
    return fmean.reshape(N, P * R), fvar.transpose(2, 1, 0).reshape(N, P * R)


def gauss_kl(q_mu, q_sqrt, K=None):
    """GPflow-1.2.0 gauss_kl (App. A.5); K=None is the whitened case."""
    M, R = q_mu.shape
    Lq = np.tril(q_sqrt)
    if K is None:
        alpha = q_mu
        trace = np.sum(Lq * Lq)
        logdet_p = 0.0
    else:
        Lp = np.linalg.cholesky(K)
        alpha = sla.solve_triangular(Lp, q_mu, lower=True)
        trace = sum(np.sum(sla.solve_triangular(Lp, Lq[r], lower=True) ** 2) for r in range(R))
        logdet_p = R * np.sum(np.log(np.diag(Lp) ** 2))
    logdet_q = np.sum(np.log(np.diagonal(Lq, axis1=1, axis2=2) ** 2))
    return 0.5 * (np.sum(alpha * alpha) - M * R - logdet_q + trace + logdet_p)


def convlayer_KL(lay, jitter=JITTER):
    """layers.py:137-147; the non-white prior is Kuu at the *initial* Z (layers.py:149-150)."""
    if lay["white"]:
        return gauss_kl(lay["q_mu"], lay["q_sqrt"], None)
    Zp = lay.get("Z_prior", lay["Z"])
    return gauss_kl(lay["q_mu"], lay["q_sqrt"], mo_Kuu(Zp, lay["variance"], lay["lengthscale"], jitter))


# ------------------------------------------------------------------ a7: kernels.py:79-136,172-178
def convkernel_Kzx(Z, X_ND, lay):
    """kernels.py:117-133: (1/P) sum_p w_p k(z_m, x_{n,p}) -> [M,N]."""
    N = X_ND.shape[0]
    X = X_ND.reshape(N, lay["H"], lay["W"], lay["C"])                      # :30-32
    patches = extract_patches(X, lay["f"], lay["s"])                       # [N,P,L]
    P, L = patches.shape[1:]
    K = rbf_K(Z, patches.reshape(N * P, L), lay["variance"], lay["lengthscale"])  # :123
    K = K.reshape(Z.shape[0], N, P) * lay["patch_weights"]                 # :127-130
    return K.sum(axis=2) / P                                               # :132-133


def convkernel_Kdiag(X_ND, lay):
    """kernels.py:106-115: (1/P^2) sum_{p,p'} w_p w_p' k(x_np, x_np') -> [N]."""
    N = X_ND.shape[0]
    X = X_ND.reshape(N, lay["H"], lay["W"], lay["C"])
    patches = extract_patches(X, lay["f"], lay["s"])
    P = patches.shape[1]
    w = lay["patch_weights"]
    W = w[None, :] * w[:, None]
    return np.array([np.sum(rbf_K(patches[n], None, lay["variance"], lay["lengthscale"]) * W)
                     for n in range(N)]) / (P ** 2)


def convkernel_Kuu(Z, lay, jitter=JITTER):
    """kernels.py:135-136 + dispatch :172-174 (jitter passed by DS/layers.py:184)."""
    return rbf_K(Z, None, lay["variance"], lay["lengthscale"]) + jitter * np.eye(Z.shape[0])


# ------------------------------------------------------------------ a7': DS/layers.py:181-256
def svgp_conditional_ND(X_ND, lay, jitter=JITTER):
    """SVGP_Layer.conditional_ND with a ConvKernel (full_cov=False, Zero mean). -> mean,var [N,R]."""
    Ku = convkernel_Kuu(lay["Z"], lay, jitter)                             # :184
    Lu = np.linalg.cholesky(Ku)                                            # :185
    Kuf = convkernel_Kzx(lay["Z"], X_ND, lay)                              # :194
    A = sla.solve_triangular(Lu, Kuf, lower=True)                          # :196
    if not lay["white"]:
        A = sla.solve_triangular(Lu.T, A, lower=False)                     # :198
    mean = A.T @ lay["q_mu"]                                               # :200
    R = lay["q_mu"].shape[1]
    M = Ku.shape[0]
    q_sqrt = np.tril(lay["q_sqrt"])
    var = np.empty((X_ND.shape[0], R))
    Kff = convkernel_Kdiag(X_ND, lay)                                      # :222
    for r in range(R):
        SK = -np.eye(M) if lay["white"] else -Ku                           # :205-208
        SK = SK + q_sqrt[r] @ q_sqrt[r].T                                  # :211
        B = SK @ A                                                         # :214
        var[:, r] = Kff + np.sum(A * B, axis=0)                            # :221,225
    return mean, var                                                       # :226-229


def svgp_KL(lay, jitter=JITTER):
    """SVGP_Layer.KL, DS/layers.py:242-256."""
    M, R = lay["q_mu"].shape
    q_sqrt = np.tril(lay["q_sqrt"])
    KL = -0.5 * R * M
    KL -= 0.5 * np.sum(np.log(np.diagonal(q_sqrt, axis1=1, axis2=2) ** 2))
    if not lay["white"]:
        Lu = np.linalg.cholesky(convkernel_Kuu(lay["Z"], lay, jitter))
        KL += np.sum(np.log(np.diag(Lu))) * R
        KL += 0.5 * sum(np.sum(sla.solve_triangular(Lu, q_sqrt[r], lower=True) ** 2) for r in range(R))
        Kinv_m = sla.cho_solve((Lu, True), lay["q_mu"])
        KL += 0.5 * np.sum(lay["q_mu"] * Kinv_m)
    else:
        KL += 0.5 * np.sum(q_sqrt ** 2)
        KL += 0.5 * np.sum(lay["q_mu"] ** 2)
    return KL


# ------------------------------------------------------------------ a8: DS/layers.py:72-105, DS/utils.py:40-41, DS/dgp.py:61-76
def reparameterize(mean, var, z, jitter=JITTER):
    """DS/utils.py:41 -- jitter sits inside the square root."""
    return mean + z * (var + jitter) ** 0.5


def layer_conditional_ND(X_ND, lay, jitter=JITTER, fast=False):
    if lay["type"] == "conv":
        return (convlayer_conditional_ND_fast if fast else convlayer_conditional_ND)(X_ND, lay, jitter)
    return svgp_conditional_ND(X_ND, lay, jitter)


def layer_KL(lay, jitter=JITTER):
    return convlayer_KL(lay, jitter) if lay["type"] == "conv" else svgp_KL(lay, jitter)


def propagate(layers, X, S, zs, jitter=JITTER, fast=False):
    """DGP_Base.propagate (DS/dgp.py:61-76) with explicit z per layer ([S,N,D_l])."""
    N = X.shape[0]
    F = np.tile(X[None], (S, 1, 1))                                        # :63
    Fs, Fmeans, Fvars = [], [], []
    for lay, z in zip(layers, zs):
        mean, var = layer_conditional_ND(F.reshape(S * N, -1), lay, jitter, fast)  # DS/layers.py:72-76
        D = mean.shape[1]
        mean, var = mean.reshape(S, N, D), var.reshape(S, N, D)
        F = reparameterize(mean, var, z, jitter)                           # DS/layers.py:105
        Fs.append(F), Fmeans.append(mean), Fvars.append(var)
    return Fs, Fmeans, Fvars


# ------------------------------------------------------------------ a9: likelihood + ELBO
def robustmax_varexp(Fmu, Fvar, Y, num_classes=10, epsilon=1e-3, n_gh=20):
    """GPflow-1.2.0 MultiClass(RobustMax).variational_expectations (App. A.5). Fmu,Fvar [n,K], Y [n]."""
    gh_x, gh_w = np.polynomial.hermite.hermgauss(n_gh)
    Y = np.asarray(Y).astype(np.int64).reshape(-1)
    n = Fmu.shape[0]
    oh = np.zeros((n, num_classes))
    oh[np.arange(n), Y] = 1.0
    mu_s = np.sum(oh * Fmu, 1)
    var_s = np.sum(oh * Fvar, 1)
    X = mu_s[:, None] + gh_x[None, :] * np.sqrt(np.clip(2.0 * var_s, 1e-10, np.inf))[:, None]
    dist = (X[:, None, :] - Fmu[:, :, None]) / np.sqrt(np.clip(Fvar, 1e-10, np.inf))[:, :, None]
    cdfs = 0.5 * (1.0 + ssp.erf(dist / np.sqrt(2.0)))
    cdfs = cdfs * (1 - 2e-4) + 1e-4
    cdfs = cdfs * (1.0 - oh)[:, :, None] + oh[:, :, None]
    p = np.prod(cdfs, axis=1) @ (gh_w / np.sqrt(np.pi))
    return p * np.log(1.0 - epsilon) + (1.0 - p) * np.log(epsilon / (num_classes - 1.0))


def robustmax_predict(Fmu, Fvar, num_classes=10, epsilon=1e-3, n_gh=20):
    """GPflow-1.2.0 MultiClass(RobustMax).predict_mean_and_var: ps [n,K] = p_c (1-eps) + (1-p_c) eps/(K-1), and ps - ps^2."""
    gh_x, gh_w = np.polynomial.hermite.hermgauss(n_gh)
    n = Fmu.shape[0]
    sd = np.sqrt(np.clip(Fvar, 1e-10, np.inf))
    ps = np.empty((n, num_classes))
    for c in range(num_classes):
        X = Fmu[:, c:c + 1] + gh_x[None, :] * np.sqrt(np.clip(2.0 * Fvar[:, c], 1e-10, np.inf))[:, None]
        cdfs = 0.5 * (1.0 + ssp.erf((X[:, None, :] - Fmu[:, :, None]) / sd[:, :, None] / np.sqrt(2.0)))
        cdfs = cdfs * (1 - 2e-4) + 1e-4
        cdfs[:, c, :] = 1.0
        p = np.prod(cdfs, axis=1) @ (gh_w / np.sqrt(np.pi))
        ps[:, c] = p * (1.0 - epsilon) + (1.0 - p) * (epsilon / (num_classes - 1.0))
    return ps, ps - ps * ps


def dgp_predict_y(layers, X, zs, S, jitter=JITTER, fast=False):
    """DGP_Base.predict_y (DS/dgp.py:116-119) with explicit samples zs -> (mean, var), each [S,N,K]."""
    N = X.shape[0]
    _, Fmeans, Fvars = propagate(layers, X, S, zs, jitter, fast)
    K = Fmeans[-1].shape[2]
    m, v = robustmax_predict(Fmeans[-1].reshape(S * N, K), Fvars[-1].reshape(S * N, K), K)
    return m.reshape(S, N, K), v.reshape(S, N, K)


def dgp_predict_density(layers, X, Y, zs, S, jitter=JITTER, fast=False):
    """DGP_Base.predict_density (DS/dgp.py:121-126) -> [N,1]."""
    N = X.shape[0]
    m, _ = dgp_predict_y(layers, X, zs, S, jitter, fast)
    y = np.asarray(Y).astype(np.int64).reshape(-1)
    l = np.log(m[:, np.arange(N), y])                                     # [S,N]
    mx = l.max(axis=0)
    return (mx + np.log(np.exp(l - mx).sum(axis=0)) - np.log(float(S)))[:, None]


def dgp_elbo(layers, X, Y, zs, num_data, S, jitter=JITTER, fast=False):
    """DGP_Base._build_likelihood (DS/dgp.py:83-98) with BroadcastingLikelihood (DS/utils.py:71-93)."""
    N = X.shape[0]
    _, Fmeans, Fvars = propagate(layers, X, S, zs, jitter, fast)
    Fmu, Fvar = Fmeans[-1], Fvars[-1]
    K = Fmu.shape[2]
    ve = robustmax_varexp(Fmu.reshape(S * N, K), Fvar.reshape(S * N, K), np.tile(np.asarray(Y).reshape(-1), S), K)
    L = np.sum(np.mean(ve.reshape(S, N), axis=0))                          # DS/dgp.py:90,94
    KL = sum(layer_KL(lay, jitter) for lay in layers)                      # :95
    return L * (float(num_data) / N) - KL                                  # :96-98
