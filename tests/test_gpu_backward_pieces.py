"""GPU: the backward of ONE layer piece by piece against float64 autograd.

dcgp_layer_backward returns the gradient of  obj = sum(g_mean * mean) + sum(g_var * var)  w.r.t. the operands of the forward in
its own order (a = Lm^-1 k;  mean_r = alpha_r^T a,  var_r = knn - |a|^2 + a^T S_r a,  S_r = C_r C_r^T): dS_r, dalpha, and --
along the direct path through Kuf / Kdiag with Lm^-1 held fixed -- the gradients w.r.t. Z, variance, lengthscale, the patch
weights and the layer input.  Restating the layer in that form in torch (float64, operands as independent leaves) gives every
piece exactly; the M-only chain rule (deepcgp_b200/grad.py) is checked with exact pieces as its input, and the two together
against autograd of the layer w.r.t. its parameters."""
import numpy as np
import pytest
import torch

import bench
from tests.test_gpu_parity import build_conv, build_last, dev, npy

pytestmark = pytest.mark.gpu


def _rbf64(X, Z, var, ls):
    Xs, Zs = X / ls, Z / ls
    d = (Xs * Xs).sum(1)[:, None] + (Zs * Zs).sum(1)[None, :] - 2.0 * Xs @ Zs.T
    return var * torch.exp(-0.5 * d)


def _patches64(X, lay):
    from oracle.dcgp_oracle_torch import extract_patches
    N = X.shape[0]
    return extract_patches(X.reshape(N, lay["H"], lay["W"], lay["C"]), lay["f"], lay["s"])       # [N, P, L]


def _t(a, rg=True):
    return torch.tensor(np.asarray(a, dtype=np.float64), requires_grad=rg)


def _layer_outputs(lay, X, Z, var, ls, w, Li, S, alpha):
    """mean, var [N, D] of the layer from a = Li k and the (S_r, alpha) operands (torch float64)."""
    M, R = lay["M"], lay["R"]
    pat = _patches64(X, lay)
    N, P, L = pat.shape
    K = _rbf64(pat.reshape(N * P, L), Z, var, ls)                                                   # [N*P, M]
    if lay["type"] == "conv":
        a = K @ Li.T
        quad = torch.einsum("tm,rmn,tn->tr", a, S, a)
        v = var - (a * a).sum(1, keepdim=True) + quad
        return (a @ alpha).reshape(N, P * R), v.reshape(N, P * R)
    Kzx = (K.reshape(N, P, M) * w[None, :, None]).sum(1) / P
    Kpp = torch.stack([_rbf64(pat[n], pat[n], var, ls) for n in range(N)])
    kdiag = (Kpp * (w[None, :] * w[:, None])[None]).sum((1, 2)) / (P * P)
    a = Kzx @ Li.T
    quad = torch.einsum("tm,rmn,tn->tr", a, S, a)
    return a @ alpha, kdiag[:, None] - (a * a).sum(1, keepdim=True) + quad


def _operands(lay, Z, var, ls, q_mu, q_sqrt):
    """Li, C_r, S_r, alpha as torch functions of the parameters (non-whitened: conditionals.py:29-58)."""
    M = lay["M"]
    Kuu = _rbf64(Z, Z, var, ls) + 1e-3 * torch.eye(M, dtype=torch.float64)
    Lm = torch.linalg.cholesky(Kuu)
    Li = torch.linalg.solve_triangular(Lm, torch.eye(M, dtype=torch.float64), upper=False)
    if lay["white"]:
        C, alpha = torch.tril(q_sqrt), q_mu
    else:
        C, alpha = Li @ torch.tril(q_sqrt), Li @ q_mu
    return Li, Lm, C, C @ C.transpose(1, 2), alpha


def _pieces_reference(lay, X32, g_mean, g_var):
    """float64 autograd of obj w.r.t. (X, Z, variance, lengthscale, S_r, alpha, patch_weights) with Lm^-1 held fixed."""
    X, Z, var, ls = _t(X32), _t(lay["Z"]), _t(lay["variance"]), _t(lay["lengthscale"])
    with torch.no_grad():
        Li, Lm, C, S, alpha = _operands(lay, _t(lay["Z"], False), _t(lay["variance"], False), _t(lay["lengthscale"], False),
                                        _t(lay["q_mu"], False), _t(lay["q_sqrt"], False))
    S, alpha = S.clone().requires_grad_(True), alpha.clone().requires_grad_(True)
    w = _t(lay["patch_weights"]) if lay["type"] != "conv" else None
    mean, v = _layer_outputs(lay, X, Z, var, ls, w, Li, S, alpha)
    obj = (_t(g_mean, False) * mean).sum() + (_t(g_var, False) * v).sum()
    leaves = [X, Z, var, ls, S, alpha] + ([w] if w is not None else [])
    grads = torch.autograd.grad(obj, leaves)
    names = ["X", "Z", "variance", "lengthscale", "S", "alpha", "patch_weights"]
    return {n: g.numpy() for n, g in zip(names, grads)}


def _params_reference(lay, X32, g_mean, g_var):
    """float64 autograd of obj w.r.t. the layer's parameters (everything moves: Kuu, Lm, C_r, alpha, Kuf, Kdiag); no KL."""
    Z, var, ls, q_mu, q_sqrt = _t(lay["Z"]), _t(lay["variance"]), _t(lay["lengthscale"]), _t(lay["q_mu"]), _t(lay["q_sqrt"])
    w = _t(lay["patch_weights"]) if lay["type"] != "conv" else None
    Li, Lm, C, S, alpha = _operands(lay, Z, var, ls, q_mu, q_sqrt)
    mean, v = _layer_outputs(lay, _t(X32, False), Z, var, ls, w, Li, S, alpha)
    obj = (_t(g_mean, False) * mean).sum() + (_t(g_var, False) * v).sum()
    leaves = [Z, var, ls, q_mu, q_sqrt] + ([w] if w is not None else [])
    grads = torch.autograd.grad(obj, leaves)
    out = {n: g.numpy() for n, g in zip(["Z", "variance", "lengthscale", "q_mu", "q_sqrt", "patch_weights"], grads)}
    out["q_sqrt"] = np.tril(out["q_sqrt"])
    return out


def _nw(x, ref):
    return float(np.max(np.abs(np.asarray(x, dtype=np.float64).reshape(ref.shape) - ref)) / max(np.max(np.abs(ref)), 1e-300))


def _run_layer(li, N, seed=31, cfg_name="cfg3", white=False):
    from deepcgp_b200.grad import LayerBackward
    from oracle import dcgp_oracle as O
    cfg = bench.CONFIGS[cfg_name]
    layers = bench.synth_params(cfg)
    rng = np.random.RandomState(seed)
    F = rng.standard_normal((N, cfg["H"] * cfg["W"] * cfg["C"]))
    for lay in layers[:li]:
        m, v = O.convlayer_conditional_ND_fast(F, lay)
        F = m + rng.standard_normal(m.shape) * np.sqrt(v + 1e-3)
    lay = dict(layers[li])
    if white:       # the same q(f) in the whitened parameterisation
        import scipy.linalg as sla
        Lm = np.linalg.cholesky(O.mo_Kuu(lay["Z"], lay["variance"], lay["lengthscale"]))
        lay.update(white=True, q_mu=sla.solve_triangular(Lm, lay["q_mu"], lower=True),
                   q_sqrt=np.stack([sla.solve_triangular(Lm, lay["q_sqrt"][r], lower=True) for r in range(lay["R"])]))
    X32 = F.astype(np.float32)
    layer = build_conv(lay, "tc") if lay["type"] == "conv" else build_last(lay, "tc")
    D = layer.num_outputs
    g_mean = (rng.standard_normal((N, D)) * 3.0).astype(np.float32)
    g_var = (rng.standard_normal((N, D)) * 2.0).astype(np.float32)
    ref = _pieces_reference(lay, X32, g_mean, g_var)
    Xd = torch.as_tensor(X32, device=dev())
    layer.prepare()
    layer._hold = True
    layer._conditional(Xd)
    lb = LayerBackward(layer)
    gX = lb.t_sized(Xd, 1, torch.as_tensor(g_mean, device=dev()), torch.as_tensor(g_var, device=dev()), True)
    layer._hold = False
    torch.cuda.synchronize()
    M, R, Mp = lay["M"], lay["R"], lb.Mp
    gS = npy(lb.gQB[Mp:(R + 1) * Mp].reshape(R, Mp, Mp)[:, :M, :M])
    galpha = npy(lb.gQB[(R + 1) * Mp:(R + 1) * Mp + R, :M]).T
    got = {"X": npy(gX), "Z": npy(lb.gZ), "variance": float(lb.gscal[0].item()), "lengthscale": float(lb.gscal[1].item()),
           "S": gS, "alpha": galpha}
    if lay["type"] != "conv":
        got["patch_weights"] = npy(lb.gw)
    errs = {k: _nw(np.asarray(got[k]), np.asarray(ref[k])) for k in got}
    return errs, lb, layer, lay, ref, (X32, g_mean, g_var)


@pytest.mark.parametrize("cfg_name,li,N", [("cfg3", 0, 3), ("cfg3", 1, 6), ("cfg3", 2, 6), ("cfg4", 1, 4), ("cfg4", 2, 4)])
def test_layer_backward_pieces_vs_float64_autograd(cfg_name, li, N):
    errs, _, _, _, _, _ = _run_layer(li, N, cfg_name=cfg_name)
    print("\n%s layer %d pieces, normwise: %s" % (cfg_name, li, {k: "%.1e" % v for k, v in errs.items()}))
    for k, e in errs.items():
        assert e <= (1e-4 if cfg_name == "cfg3" else 3e-4), (k, e)


@pytest.mark.parametrize("cfg_name,li,N,white", [("cfg3", 1, 6, False), ("cfg3", 2, 6, False), ("cfg3", 1, 4, True), ("cfg4", 1, 4, False)])
def test_layer_parameter_gradients_vs_float64_autograd(cfg_name, li, N, white):
    """dcgp_layer_backward + grad.LayerBackward.m_only (no KL) against autograd of the layer w.r.t. its parameters; and the chain
    rule alone, fed with the exact float64 pieces."""
    errs, lb, layer, lay, ref, (X32, g_mean, g_var) = _run_layer(li, N, cfg_name=cfg_name, white=white)
    refp = _params_reference(lay, X32, g_mean, g_var)
    get = lambda g, k: npy(g[k]) if isinstance(g[k], torch.Tensor) else np.asarray(g[k])
    got = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in lb.m_only(kl_weight=0.0).items()}
    e1 = {k: _nw(get(got, k), refp[k]) for k in refp}
    # the native chain rule (dcgp_layer_chain_rule) against its torch restatement, KL gradient included
    nat = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in lb.m_only(kl_weight=0.5).items()}
    tor = lb.m_only_torch(kl_weight=0.5)
    e0 = {k: _nw(get(nat, k), get(tor, k)) for k in refp}
    print("\n%s layer %d white=%s native vs torch chain rule (with KL), normwise: %s" % (cfg_name, li, white, {k: "%.1e" % v for k, v in e0.items()}))
    for k, e in e0.items():
        assert e <= 2e-5, ("native vs torch", k, e)
    print("\n%s layer %d white=%s parameter gradients, normwise: %s" % (cfg_name, li, white, {k: "%.1e" % v for k, v in e1.items()}))
    M, R, Mp = lay["M"], lay["R"], lb.Mp
    lb.gQB.zero_()
    lb.gQB[Mp:(R + 1) * Mp].reshape(R, Mp, Mp)[:, :M, :M] = torch.as_tensor(ref["S"], device=dev())
    lb.gQB[(R + 1) * Mp:(R + 1) * Mp + R, :M] = torch.as_tensor(ref["alpha"].T.copy(), device=dev())
    lb.gZ.copy_(torch.as_tensor(ref["Z"], device=dev()))
    lb.gscal[0], lb.gscal[1] = float(ref["variance"]), float(ref["lengthscale"])
    if lay["type"] != "conv":
        lb.gw.copy_(torch.as_tensor(ref["patch_weights"], device=dev()))
    got2 = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in lb.m_only(kl_weight=0.0).items()}
    e2 = {k: _nw(get(got2, k), refp[k]) for k in refp}
    print("%s layer %d white=%s chain rule with exact pieces, normwise: %s" % (cfg_name, li, white, {k: "%.1e" % v for k, v in e2.items()}))
    for k, e in e2.items():
        assert e <= 1e-4, ("chain rule", k, e)
    for k, e in e1.items():
        assert e <= 3e-4, ("layer", k, e)
