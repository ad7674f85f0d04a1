"""CPU: the float64 oracle restatement vs. golden vectors produced by the reference's own source
(tests/golden/make_golden.py), plus the self-checks of SURVEY.md App. A.6."""
import numpy as np
import pytest
import scipy.linalg as sla

from oracle import dcgp_oracle as O
from tests.util import golden_names, layer_from_golden, layers_from_golden, load_golden

TOL = dict(rtol=1e-11, atol=1e-12)


@pytest.mark.parametrize("name", golden_names("convlayer_"))
def test_convlayer_pieces(name):
    g = load_golden(name)
    lay = layer_from_golden(g, 0, "conv")
    jit = float(g["jitter"])
    X = g["X"]
    N = X.shape[0]
    NHWC = X.reshape(N, lay["H"], lay["W"], lay["C"])
    assert O.out_image_size(lay["H"], lay["W"], lay["f"], lay["s"]) == (int(g["out_h"]), int(g["out_w"]))
    np.testing.assert_array_equal(O.extract_patches_PNL(NHWC, lay["f"], lay["s"]), g["PNL"])
    np.testing.assert_array_equal(O.extract_patches(NHWC, lay["f"], lay["s"]), g["NPL"])
    assert g["PNL"].shape == (int(g["patch_count"]), N, int(g["patch_length"]))
    np.testing.assert_allclose(O.mo_Kuu(lay["Z"], lay["variance"], lay["lengthscale"], jit), g["Kuu"], **TOL)
    np.testing.assert_allclose(O.mo_Kuf(lay["Z"], g["PNL"], lay["variance"], lay["lengthscale"]), g["Kuf"], **TOL)
    np.testing.assert_allclose(O.mo_Kdiag(g["PNL"], lay["variance"]), g["Knn"], **TOL)
    fmean, fvar = O.conditional(g["Kuf"], g["Kuu"], g["Knn"], lay["q_mu"], lay["q_sqrt"], lay["white"])
    np.testing.assert_allclose(fmean, g["fmean"], **TOL)
    np.testing.assert_allclose(fvar, g["fvar"], **TOL)
    fmean2, fvar2 = O.conditional_single_solve(g["Kuf"], g["Kuu"], g["Knn"], lay["q_mu"], lay["q_sqrt"], lay["white"])
    np.testing.assert_allclose(fmean2, g["fmean"], rtol=1e-9, atol=1e-10)   # A.6 (4)
    np.testing.assert_allclose(fvar2, g["fvar"], rtol=1e-9, atol=1e-10)
    mean, var = O.convlayer_conditional_ND(X, lay, jit)
    np.testing.assert_allclose(mean, g["mean"], **TOL)
    np.testing.assert_allclose(var, g["var"], **TOL)
    np.testing.assert_allclose(O.convlayer_KL(lay, jit), g["KL"], rtol=1e-11)


@pytest.mark.parametrize("name", golden_names("lastlayer_"))
def test_lastlayer(name):
    g = load_golden(name)
    lay = layer_from_golden(g, 0, "svgp_conv")
    jit = float(g["jitter"])
    np.testing.assert_allclose(O.convkernel_Kzx(lay["Z"], g["X"], lay), g["Kzx"], **TOL)
    np.testing.assert_allclose(O.convkernel_Kdiag(g["X"], lay), g["Kdiag"], **TOL)
    np.testing.assert_allclose(O.convkernel_Kuu(lay["Z"], lay, 0.0), g["Kzz"], **TOL)
    mean, var = O.svgp_conditional_ND(g["X"], lay, jit)
    np.testing.assert_allclose(mean, g["mean"], rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(var, g["var"], rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(O.svgp_KL(lay, jit), g["KL"], rtol=1e-11)


@pytest.mark.parametrize("name", golden_names("dgp"))
def test_dgp_elbo(name):
    g = load_golden(name)
    layers = layers_from_golden(g)
    S, jit = int(g["S"]), float(g["jitter"])
    zs = [g["z%d" % i] for i in range(len(layers))]
    Fs, Fmeans, Fvars = O.propagate(layers, g["X"], S, zs, jit)
    for i in range(len(layers)):
        np.testing.assert_allclose(Fmeans[i], g["Fmean%d" % i], rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(Fvars[i], g["Fvar%d" % i], rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(Fs[i], g["F%d" % i], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose([O.layer_KL(l, jit) for l in layers], g["KLs"], rtol=1e-10)
    elbo = O.dgp_elbo(layers, g["X"], g["Y"], zs, float(g["num_data"]), S, jit)
    np.testing.assert_allclose(elbo, g["elbo"], rtol=1e-10)
    elbo_fast = O.dgp_elbo(layers, g["X"], g["Y"], zs, float(g["num_data"]), S, jit, fast=True)
    np.testing.assert_allclose(elbo_fast, g["elbo"], rtol=1e-9)


@pytest.mark.parametrize("name", golden_names("dgp"))
def test_dgp_prediction_path(name):
    """DS/dgp.py:116-126 predict_y / predict_density on the golden samples (BroadcastingLikelihood from the reference source)."""
    g = load_golden(name)
    layers = layers_from_golden(g)
    S, jit = int(g["S"]), float(g["jitter"])
    zs = [g["z%d" % i] for i in range(len(layers))]
    m, v = O.dgp_predict_y(layers, g["X"], zs, S, jit)
    np.testing.assert_allclose(m, g["pred_mean"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(v, g["pred_var"], rtol=1e-9, atol=1e-12)
    ld = O.dgp_predict_density(layers, g["X"], g["Y"], zs, S, jit)
    np.testing.assert_allclose(ld, g["pred_logdensity"], rtol=1e-9)
    assert np.all(m > 0) and np.all(m.sum(-1) < 1.0 + 1e-9)


# ----------------------------------------------------------------------------- App. A.6 self-checks
def _rand_layer(rng, H=7, W=7, C=2, f=3, s=1, M=9, R=3, white=False):
    L = f * f * C
    q_sqrt = np.tril(rng.standard_normal((R, M, M)) * 0.3) + 0.5 * np.eye(M)
    return dict(type="conv", H=H, W=W, C=C, f=f, s=s, M=M, R=R, white=white, variance=1.7, lengthscale=2.3,
                Z=rng.standard_normal((M, L)), q_mu=rng.standard_normal((M, R)), q_sqrt=q_sqrt)


def test_patch_order_vs_slicing():
    """A.6 (6), in the style of reference tests/test_views.py:27-29."""
    rng = np.random.RandomState(0)
    X = rng.standard_normal((2, 9, 8, 3))
    f, s = 4, 2
    PNL = O.extract_patches_PNL(X, f, s)
    OH, OW = O.out_image_size(9, 8, f, s)
    for oy in range(OH):
        for ox in range(OW):
            np.testing.assert_array_equal(PNL[oy * OW + ox, 1], X[1, oy * s:oy * s + f, ox * s:ox * s + f, :].reshape(-1))


def test_single_patch_equals_dense_svgp():
    """A.6 (1): f=H=W -> P=1 and the conv layer is a dense SVGP on flattened images."""
    rng = np.random.RandomState(1)
    lay = _rand_layer(rng, H=4, W=4, C=2, f=4, s=1, M=6, R=3)
    X = rng.standard_normal((5, 32))
    mean, var = O.convlayer_conditional_ND(X, lay)
    Kuu = O.rbf_K(lay["Z"], None, 1.7, 2.3) + O.JITTER * np.eye(6)
    Kuf = O.rbf_K(lay["Z"], X, 1.7, 2.3)
    A = np.linalg.solve(Kuu, Kuf)
    np.testing.assert_allclose(mean, A.T @ lay["q_mu"], rtol=1e-9, atol=1e-11)
    for r in range(3):
        S = lay["q_sqrt"][r] @ lay["q_sqrt"][r].T
        v = 1.7 + np.einsum("mn,mk,kn->n", A, S - Kuu, A)
        np.testing.assert_allclose(var[:, r], v, rtol=1e-8, atol=1e-10)


def test_white_nonwhite_equivalence():
    """A.6 (2)."""
    rng = np.random.RandomState(2)
    lay = _rand_layer(rng)
    X = rng.standard_normal((3, 7 * 7 * 2))
    mean, var = O.convlayer_conditional_ND(X, lay)
    Lm = np.linalg.cholesky(O.mo_Kuu(lay["Z"], lay["variance"], lay["lengthscale"]))
    layw = dict(lay, white=True, q_mu=sla.solve_triangular(Lm, lay["q_mu"], lower=True),
                q_sqrt=np.stack([sla.solve_triangular(Lm, lay["q_sqrt"][r], lower=True) for r in range(3)]))
    meanw, varw = O.convlayer_conditional_ND(X, layw)
    np.testing.assert_allclose(meanw, mean, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(varw, var, rtol=1e-8, atol=1e-10)


def test_prior_recovery():
    """A.6 (3): q = prior  =>  var == sigma^2, mean == 0, KL == 0."""
    rng = np.random.RandomState(3)
    lay = _rand_layer(rng)
    Lm = np.linalg.cholesky(O.mo_Kuu(lay["Z"], lay["variance"], lay["lengthscale"]))
    lay["q_mu"] = np.zeros_like(lay["q_mu"])
    lay["q_sqrt"] = np.tile(Lm[None], (3, 1, 1))
    X = rng.standard_normal((3, 7 * 7 * 2))
    mean, var = O.convlayer_conditional_ND(X, lay)
    np.testing.assert_allclose(mean, 0.0, atol=1e-12)
    np.testing.assert_allclose(var, lay["variance"], rtol=1e-9)
    np.testing.assert_allclose(O.convlayer_KL(lay), 0.0, atol=1e-9)


def test_convkernel_limits():
    """A.6 (5): Kzx with P=1 is the plain RBF; Kdiag is the diagonal of ConvKernel.K (kernels.py:81-104)."""
    rng = np.random.RandomState(4)
    lay = dict(H=3, W=3, C=2, f=3, s=1, variance=0.8, lengthscale=1.9, patch_weights=np.ones(1))
    X = rng.standard_normal((4, 18))
    Z = rng.standard_normal((5, 18))
    np.testing.assert_allclose(O.convkernel_Kzx(Z, X, lay), O.rbf_K(Z, X, 0.8, 1.9), rtol=1e-12)
    lay = dict(H=5, W=5, C=1, f=3, s=1, variance=0.8, lengthscale=1.9, patch_weights=0.5 + rng.random(9))
    X = rng.standard_normal((3, 25))
    pat = O.extract_patches(X.reshape(3, 5, 5, 1), 3, 1)
    K = O.rbf_K(pat.reshape(27, 9), None, 0.8, 1.9).reshape(3, 9, 3, 9)
    w = lay["patch_weights"]
    Kfull = (K * (w[None, :, None, None] * w[None, None, None, :])).sum(axis=(1, 3)) / 81.0
    np.testing.assert_allclose(O.convkernel_Kdiag(X, lay), np.diag(Kfull), rtol=1e-12)
