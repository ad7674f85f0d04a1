"""TEST INFRASTRUCTURE ONLY -- the float64 oracle (oracle/dcgp_oracle.py, single-solve form of SURVEY App. A.4) restated
with torch (CPU, float64) so that autograd supplies the reference gradients of the ELBO, the way TensorFlow autodiff does in
the reference (tf.gradients inside GPflow's AdamOptimizer, experiment.py:84-108).  Values are cross-checked against the
NumPy oracle in tests/test_oracle_golden.py; only tests/ and bench.py's CPU legs may import this module."""
import math

import numpy as np
import torch

JITTER = 1e-3
PARAMS = ("Z", "variance", "lengthscale", "q_mu", "q_sqrt", "patch_weights")


def to_torch_layers(layers, requires_grad=True):
    out = []
    for lay in layers:
        t = dict(lay)
        for k in PARAMS:
            if k in lay:
                t[k] = torch.tensor(np.asarray(lay[k], dtype=np.float64), dtype=torch.float64, requires_grad=requires_grad)
        if "Z_prior" in lay:
            t["Z_prior"] = torch.tensor(np.asarray(lay["Z_prior"]), dtype=torch.float64)
        out.append(t)
    return out


def extract_patches(X, f, s):
    """[N,H,W,C] -> [N,P,L], (dy,dx,c) order (views.py:32-54)."""
    N, H, W, C = X.shape
    OH, OW = (H - f) // s + 1, (W - f) // s + 1
    cols = []
    for dy in range(f):
        for dx in range(f):
            cols.append(X[:, dy:dy + (OH - 1) * s + 1:s, dx:dx + (OW - 1) * s + 1:s, :])
    return torch.stack(cols, dim=3).reshape(N, OH * OW, f * f * C)


def rbf(X, X2, variance, lengthscale):
    X = X / lengthscale
    Xs = (X * X).sum(1)
    if X2 is None:
        d = -2.0 * X @ X.T + Xs[:, None] + Xs[None, :]
    else:
        X2 = X2 / lengthscale
        d = -2.0 * X @ X2.T + Xs[:, None] + (X2 * X2).sum(1)[None, :]
    return variance * torch.exp(-0.5 * d)


def _cond(Kuf_TM, Kuu, knn, q_mu, q_sqrt, white):
    """Kuf_TM [T,M] -> mean [T,R], var [T,R]."""
    Lm = torch.linalg.cholesky(Kuu)
    L = torch.tril(q_sqrt)
    a = torch.linalg.solve_triangular(Lm, Kuf_TM.T, upper=False)            # [M,T]
    if white:
        alpha, C = q_mu, L
    else:
        alpha = torch.linalg.solve_triangular(Lm, q_mu, upper=False)
        C = torch.linalg.solve_triangular(Lm, L, upper=False)               # [R,M,M]
    base = knn - (a * a).sum(0)
    mean = a.T @ alpha
    CTa = torch.matmul(C.transpose(1, 2), a)                                # [R,M,T]
    var = base[:, None] + (CTa * CTa).sum(1).T
    return mean, var


def gauss_kl(q_mu, q_sqrt, K):
    M, R = q_mu.shape
    Lq = torch.tril(q_sqrt)
    logdet_q = torch.log(torch.diagonal(Lq, dim1=1, dim2=2) ** 2).sum()
    if K is None:
        return 0.5 * ((q_mu ** 2).sum() - M * R - logdet_q + (Lq ** 2).sum())
    Lp = torch.linalg.cholesky(K)
    alpha = torch.linalg.solve_triangular(Lp, q_mu, upper=False)
    LpiLq = torch.linalg.solve_triangular(Lp, Lq, upper=False)
    return 0.5 * ((alpha ** 2).sum() - M * R - logdet_q + (LpiLq ** 2).sum() + R * torch.log(torch.diagonal(Lp) ** 2).sum())


def layer_forward(X_ND, lay, jitter=JITTER):
    N = X_ND.shape[0]
    X = X_ND.reshape(N, lay["H"], lay["W"], lay["C"])
    pat = extract_patches(X, lay["f"], lay["s"])                            # [N,P,L]
    P, L = pat.shape[1:]
    M = lay["Z"].shape[0]
    Kuu = rbf(lay["Z"], None, lay["variance"], lay["lengthscale"]) + jitter * torch.eye(M, dtype=torch.float64)
    K = rbf(pat.reshape(N * P, L), lay["Z"], lay["variance"], lay["lengthscale"])   # [N*P, M]
    R = lay["q_mu"].shape[1]
    if lay["type"] == "conv":
        mean, var = _cond(K, Kuu, lay["variance"], lay["q_mu"], lay["q_sqrt"], lay["white"])
        return mean.reshape(N, P * R), var.reshape(N, P * R)
    w = lay["patch_weights"]
    Kzx = (K.reshape(N, P, M) * w[None, :, None]).sum(1) / P                # kernels.py:117-133
    W2 = w[None, :] * w[:, None]
    Kpp = torch.stack([rbf(pat[n], None, lay["variance"], lay["lengthscale"]) for n in range(N)])
    kdiag = (Kpp * W2[None]).sum((1, 2)) / (P * P)                           # kernels.py:106-115
    return _cond(Kzx, Kuu, kdiag, lay["q_mu"], lay["q_sqrt"], lay["white"])


def layer_kl(lay, jitter=JITTER):
    if lay["white"]:
        return gauss_kl(lay["q_mu"], lay["q_sqrt"], None)
    # ConvLayer: the prior is Kuu at the Z VALUE the layer was built with (layers.py:149-150) -- a constant, no gradient
    Zp = lay.get("Z_prior", lay["Z"]).detach() if lay["type"] == "conv" else lay["Z"]
    M = Zp.shape[0]
    K = rbf(Zp, None, lay["variance"], lay["lengthscale"]) + jitter * torch.eye(M, dtype=torch.float64)
    return gauss_kl(lay["q_mu"], lay["q_sqrt"], K)


def robustmax_varexp(Fmu, Fvar, Y, num_classes=10, epsilon=1e-3):
    gh_x, gh_w = np.polynomial.hermite.hermgauss(20)
    gh_x, gh_w = torch.tensor(gh_x), torch.tensor(gh_w / math.sqrt(math.pi))
    n = Fmu.shape[0]
    oh = torch.zeros((n, num_classes), dtype=torch.float64)
    oh[torch.arange(n), torch.as_tensor(Y, dtype=torch.long).reshape(-1)] = 1.0
    mu_s, var_s = (oh * Fmu).sum(1), (oh * Fvar).sum(1)
    X = mu_s[:, None] + gh_x[None, :] * torch.sqrt(torch.clamp(2.0 * var_s, min=1e-10))[:, None]
    dist = (X[:, None, :] - Fmu[:, :, None]) / torch.sqrt(torch.clamp(Fvar, min=1e-10))[:, :, None]
    cdfs = 0.5 * (1.0 + torch.erf(dist / math.sqrt(2.0)))
    cdfs = cdfs * (1 - 2e-4) + 1e-4
    cdfs = cdfs * (1.0 - oh)[:, :, None] + oh[:, :, None]
    p = cdfs.prod(1) @ gh_w
    return p * math.log(1.0 - epsilon) + (1.0 - p) * math.log(epsilon / (num_classes - 1.0))


def dgp_elbo(layers, X, Y, zs, num_data, S, jitter=JITTER, n_global=None, keep=None):
    """ELBO (DS/dgp.py:92-98) as a differentiable torch scalar; layers from to_torch_layers().
    keep: optional list that receives (mean, var) [S,N,D] numpy arrays of every layer (DS/dgp.py:61-76 Fmeans, Fvars)."""
    X = torch.as_tensor(X, dtype=torch.float64)
    N = X.shape[0]
    F = X[None].repeat(S, 1, 1)
    for lay, z in zip(layers, zs):
        mean, var = layer_forward(F.reshape(S * N, -1), lay, jitter)
        D = mean.shape[1]
        mean, var = mean.reshape(S, N, D), var.reshape(S, N, D)
        if keep is not None:
            keep.append((mean.detach().numpy().copy(), var.detach().numpy().copy()))
        F = mean + torch.as_tensor(z, dtype=torch.float64) * torch.sqrt(var + jitter)
    K = mean.shape[2]
    ve = robustmax_varexp(mean.reshape(S * N, K), var.reshape(S * N, K), np.tile(np.asarray(Y).reshape(-1), S), K)
    Lsum = ve.reshape(S, N).mean(0).sum()
    KL = sum(layer_kl(lay, jitter) for lay in layers)
    return Lsum * (float(num_data) / float(n_global or N)) - KL


def elbo_and_grads(layers_np, X, Y, zs, num_data, S, jitter=JITTER, keep=None):
    """Returns (elbo, [ {param: dELBO/dparam (numpy)} per layer ]); q_sqrt gradients are lower-triangular.
    keep: see dgp_elbo."""
    layers = to_torch_layers(layers_np)
    elbo = dgp_elbo(layers, X, Y, zs, num_data, S, jitter, keep=keep)
    elbo.backward()
    grads = []
    for lay in layers:
        g = {}
        for k in PARAMS:
            if k in lay and isinstance(lay[k], torch.Tensor) and lay[k].grad is not None:
                g[k] = lay[k].grad.numpy().copy()
        if "q_sqrt" in g:
            g["q_sqrt"] = np.tril(g["q_sqrt"])
        grads.append(g)
    return float(elbo.item()), grads
