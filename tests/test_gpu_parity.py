"""GPU parity tests: the CUDA path (through the Python mirror -> ctypes -> C ABI of libdcgp.so) against
(1) golden vectors produced by the reference's own source and (2) the float64 oracle on seeded inputs.

Tolerances (BASELINE.json north_star / BASELINE.md section 3):
  conditional mean / var : normwise <= 1e-4 and element-wise |d| <= 1e-4*|ref| + 1e-4*sigma^2
  ELBO                   : 1e-3 relative
  integer / index work (patch extraction): bit exact
"""
import numpy as np
import pytest
import torch

from tests.util import assert_parity, golden_names, layer_from_golden, layers_from_golden, load_golden

pytestmark = pytest.mark.gpu

ALGOS = ["simt", "tc"]
# KL: the fp64 path reproduces the reference to rounding; on the tensor-core path the O(R M^3) trace term
# sum_r |Lp^-1 L_r|_F^2 runs as a split-fp16 GEMM with fp32 accumulation (the ELBO gate is 1e-3).
KL_RTOL = {"simt": 1e-9, "tc": 5e-6}


def _algo(name):
    import deepcgp_b200 as D
    return D.ALGO_SIMT if name == "simt" else D.ALGO_TC


def dev():
    return torch.device("cuda:0")


def build_conv(lay, algo):
    import deepcgp_b200 as D
    view = D.FullView((lay["H"], lay["W"]), lay["f"], lay["C"], lay["s"])
    kern = D.RBF(lay["f"] ** 2 * lay["C"], variance=lay["variance"], lengthscales=lay["lengthscale"])
    layer = D.ConvLayer(kern, D.Zero(), feature=D.PatchInducingFeatures(lay["Z"]), view=view, white=lay["white"],
                        gp_count=lay["R"], q_mu=lay["q_mu"], q_sqrt=lay["q_sqrt"], device=dev())
    layer.algo = _algo(algo)
    return layer


def build_last(lay, algo):
    import deepcgp_b200 as D
    view = D.FullView((lay["H"], lay["W"], lay["C"]), lay["f"], lay["C"], lay["s"])       # models.py:173
    kern = D.ConvKernel(D.RBF(lay["f"] ** 2 * lay["C"], variance=lay["variance"], lengthscales=lay["lengthscale"]),
                        view=view, patch_weights=lay.get("patch_weights"))
    layer = D.SVGP_Layer(kern, lay["R"], D.Zero(lay["R"]), feature=D.PatchInducingFeatures(lay["Z"]),
                         white=lay["white"], q_mu=lay["q_mu"], q_sqrt=lay["q_sqrt"], device=dev())
    layer.algo = _algo(algo)
    return layer


def build_model(layers, X, Y, S, num_data, algo):
    import deepcgp_b200 as D
    ls = [build_conv(l, algo) if l["type"] == "conv" else build_last(l, algo) for l in layers]
    return D.DGP_Base(X, Y, D.MultiClass(10), ls, num_samples=S, num_data=num_data, device=dev())


def npy(t):
    return t.detach().cpu().numpy().astype(np.float64)


# ----------------------------------------------------------------------------------------------- a1/a2
@pytest.mark.parametrize("name", golden_names("convlayer_"))
def test_patches_bit_exact(name):
    import deepcgp_b200 as D
    g = load_golden(name)
    lay = layer_from_golden(g, 0, "conv")
    X32 = g["X"].astype(np.float32)
    view = D.FullView((lay["H"], lay["W"]), lay["f"], lay["C"], lay["s"])
    assert (view.patch_count, view.patch_length) == (int(g["patch_count"]), int(g["patch_length"]))
    assert (view.out_image_height, view.out_image_width) == (int(g["out_h"]), int(g["out_w"]))
    NHWC = torch.as_tensor(X32.reshape(-1, lay["H"], lay["W"], lay["C"]), device=dev())
    np.testing.assert_array_equal(view.extract_patches_PNL(NHWC).cpu().numpy(), g["PNL"].astype(np.float32))
    np.testing.assert_array_equal(view.extract_patches(NHWC).cpu().numpy(), g["NPL"].astype(np.float32))


# ----------------------------------------------------------------------------------------------- a3/a4
@pytest.mark.parametrize("name", golden_names("convlayer_"))
def test_kuu_kuf(name):
    import deepcgp_b200 as D
    g = load_golden(name)
    lay = layer_from_golden(g, 0, "conv")
    kern = D.RBF(lay["f"] ** 2 * lay["C"], variance=lay["variance"], lengthscales=lay["lengthscale"])
    mok = D.MultiOutputConvKernel(kern, lay["H"] * lay["W"] * lay["C"], int(g["patch_count"]))
    Z = torch.as_tensor(lay["Z"], device=dev())
    np.testing.assert_allclose(npy(mok.Kuu(Z)), g["Kuu"], rtol=1e-12, atol=1e-13)
    PNL = torch.as_tensor(g["PNL"].astype(np.float32), device=dev())
    Kuf = npy(mok.Kuf(Z, PNL))
    assert Kuf.shape == g["Kuf"].shape
    np.testing.assert_allclose(Kuf, g["Kuf"], rtol=2e-5, atol=1e-6 * lay["variance"])
    NHWC = torch.as_tensor(g["X"].astype(np.float32).reshape(-1, lay["H"], lay["W"], lay["C"]), device=dev())
    Kuf2 = npy(mok.Kuf_images(Z, NHWC, lay["f"], lay["s"]))
    np.testing.assert_allclose(Kuf2, g["Kuf"], rtol=2e-5, atol=1e-6 * lay["variance"])
    np.testing.assert_allclose(npy(mok.Kdiag(PNL)), g["Knn"], rtol=1e-7)


# ----------------------------------------------------------------------------------------------- K-C
@pytest.mark.parametrize("M", [1, 17, 64, 65, 129, 200, 512, 1024])
def test_cholesky(M):
    from deepcgp_b200 import _lib
    rng = np.random.RandomState(M)
    A = rng.standard_normal((M, M + 3))
    K = A @ A.T + 0.5 * np.eye(M)
    Kd = torch.as_tensor(K, device=dev()).contiguous()
    ws = torch.empty(_lib.lib.dcgp_cholesky_workspace_bytes(M), dtype=torch.uint8, device=dev())
    info = torch.ones(1, dtype=torch.int32, device=dev())
    _lib.check(_lib.lib.dcgp_cholesky(_lib.ptr(Kd), M, _lib.ptr(ws), ws.numel(), _lib.ptr(info), _lib.stream()))
    assert int(info.item()) == 0
    np.testing.assert_allclose(npy(Kd), np.linalg.cholesky(K), rtol=1e-9, atol=1e-10)


def test_cholesky_not_pd_is_reported():
    from deepcgp_b200 import _lib
    M = 100
    K = np.eye(M)
    K[70, 70] = -1.0
    Kd = torch.as_tensor(K, device=dev()).contiguous()
    ws = torch.empty(_lib.lib.dcgp_cholesky_workspace_bytes(M), dtype=torch.uint8, device=dev())
    info = torch.zeros(1, dtype=torch.int32, device=dev())
    _lib.check(_lib.lib.dcgp_cholesky(_lib.ptr(Kd), M, _lib.ptr(ws), ws.numel(), _lib.ptr(info), _lib.stream()))
    assert int(info.item()) == 71
    with pytest.raises(FloatingPointError):
        _lib.raise_if_not_pd(info)


# ----------------------------------------------------------------------------------------------- a5
@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("name", golden_names("convlayer_"))
def test_conditional_vs_golden(name, algo):
    import deepcgp_b200 as D
    g = load_golden(name)
    lay = layer_from_golden(g, 0, "conv")
    t = lambda a, dt: torch.as_tensor(a.astype(dt), device=dev())
    fmean, fvar = D.conditional(t(g["Kuf"], np.float32), t(g["Kuu"], np.float64), t(g["Knn"], np.float32),
                                t(lay["q_mu"], np.float64), q_sqrt=t(lay["q_sqrt"], np.float64), white=lay["white"],
                                algo=_algo(algo))
    assert fmean.shape == g["fmean"].shape and fvar.shape == g["fvar"].shape
    assert_parity(npy(fmean), g["fmean"], lay["variance"], "fmean")
    assert_parity(npy(fvar), g["fvar"], lay["variance"], "fvar")


# ----------------------------------------------------------------------------------------------- a6 / a6'
@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("name", golden_names("convlayer_"))
def test_convlayer_vs_golden(name, algo):
    g = load_golden(name)
    lay = layer_from_golden(g, 0, "conv")
    layer = build_conv(lay, algo)
    X = torch.as_tensor(g["X"].astype(np.float32), device=dev())
    mean, var = layer.conditional_ND(X)
    assert mean.shape == g["mean"].shape
    assert_parity(npy(mean), g["mean"], lay["variance"], "mean")
    assert_parity(npy(var), g["var"], lay["variance"], "var")
    np.testing.assert_allclose(float(layer.KL().item()), float(g["KL"]), rtol=KL_RTOL[algo])


def test_convlayer_kl_uses_initial_Z():
    """layers.py:149-150: moving Z after construction must not move the KL prior."""
    from oracle import dcgp_oracle as O
    g = load_golden("convlayer_a")
    lay = layer_from_golden(g, 0, "conv")
    layer = build_conv(lay, "simt")
    rng = np.random.RandomState(5)
    Z_new = lay["Z"] + 0.3 * rng.standard_normal(lay["Z"].shape)
    layer.feature.Z = torch.as_tensor(Z_new, device=dev())
    ref = O.convlayer_KL(dict(lay, Z=Z_new, Z_prior=lay["Z"]), float(g["jitter"]))
    np.testing.assert_allclose(float(layer.KL().item()), ref, rtol=1e-9)
    X = torch.as_tensor(g["X"].astype(np.float32), device=dev())
    mean, var = layer.conditional_ND(X)
    mref, vref = O.convlayer_conditional_ND(g["X"].astype(np.float32).astype(np.float64), dict(lay, Z=Z_new))
    assert_parity(npy(mean), mref, lay["variance"], "mean(Z moved)")
    assert_parity(npy(var), vref, lay["variance"], "var(Z moved)")


# ----------------------------------------------------------------------------------------------- a7 / a7'
@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("name", golden_names("lastlayer_"))
def test_lastlayer_vs_golden(name, algo):
    g = load_golden(name)
    lay = layer_from_golden(g, 0, "svgp_conv")
    layer = build_last(lay, algo)
    X = torch.as_tensor(g["X"].astype(np.float32), device=dev())
    Z = torch.as_tensor(lay["Z"], device=dev())
    np.testing.assert_allclose(npy(layer.kern.Kzx(Z, X)), g["Kzx"], rtol=2e-5, atol=1e-6 * lay["variance"])
    np.testing.assert_allclose(npy(layer.kern.Kdiag(X)), g["Kdiag"], rtol=2e-5)
    np.testing.assert_allclose(npy(layer.kern.Kzz(Z)), g["Kzz"], rtol=1e-12, atol=1e-13)
    mean, var = layer.conditional_ND(X)
    assert_parity(npy(mean), g["mean"], lay["variance"], "mean")
    assert_parity(npy(var), g["var"], lay["variance"], "var")
    np.testing.assert_allclose(float(layer.KL().item()), float(g["KL"]), rtol=KL_RTOL[algo])


# ----------------------------------------------------------------------------------------------- a8 / a9
@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("name", golden_names("dgp"))
def test_dgp_elbo_vs_golden(name, algo):
    g = load_golden(name)
    layers = layers_from_golden(g)
    S = int(g["S"])
    model = build_model(layers, g["X"].astype(np.float32), g["Y"], S, float(g["num_data"]), algo)
    zs = [torch.as_tensor(g["z%d" % i].astype(np.float32), device=dev()) for i in range(len(layers))]
    Fs, Fmeans, Fvars = model.propagate(torch.as_tensor(g["X"].astype(np.float32)), S=S, zs=zs)
    # First layer sees the golden inputs exactly; deeper layers inherit fp32 rounding of their inputs, so they
    # are compared with the same metric (the reference is a float64 graph end to end).
    for i, lay in enumerate(layers):
        assert_parity(npy(Fmeans[i]), g["Fmean%d" % i], lay["variance"], "Fmean%d" % i)
        assert_parity(npy(Fvars[i]), g["Fvar%d" % i], lay["variance"], "Fvar%d" % i)
    elbo = model.compute_log_likelihood(g["X"].astype(np.float32), g["Y"], zs=zs)
    assert abs(elbo - float(g["elbo"])) <= 1e-3 * abs(float(g["elbo"])), (elbo, float(g["elbo"]))
    np.testing.assert_allclose(npy(model._kls), g["KLs"], rtol=KL_RTOL[algo])
    ve = model.likelihood.variational_expectations(Fmeans[-1], Fvars[-1], g["Y"])
    np.testing.assert_allclose(npy(ve), g["varexp"], rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("name", golden_names("dgp"))
def test_prediction_path_vs_golden(name):
    """predict_y / predict_density (DS/dgp.py:116-126) against the reference's BroadcastingLikelihood on the golden samples."""
    g = load_golden(name)
    layers = layers_from_golden(g)
    S = int(g["S"])
    X32 = g["X"].astype(np.float32)
    model = build_model(layers, X32, g["Y"], S, float(g["num_data"]), "tc")
    zs = [torch.as_tensor(g["z%d" % i].astype(np.float32), device=dev()) for i in range(len(layers))]
    pm, pv = model.predict_y(X32, S, zs=zs)
    assert tuple(pm.shape) == g["pred_mean"].shape
    np.testing.assert_allclose(npy(pm), g["pred_mean"], rtol=2e-3, atol=2e-5)
    np.testing.assert_allclose(npy(pv), g["pred_var"], rtol=2e-3, atol=2e-5)
    ld = model.predict_density(X32, g["Y"], S, zs=zs)
    np.testing.assert_allclose(npy(ld), g["pred_logdensity"], rtol=1e-3, atol=1e-4)
    # the likelihood kernel alone, on the golden (float64) inputs rounded to float32
    Fm = torch.as_tensor(g["Fmean%d" % (len(layers) - 1)].astype(np.float32), device=dev())
    Fv = torch.as_tensor(g["Fvar%d" % (len(layers) - 1)].astype(np.float32), device=dev())
    m2, v2 = model.likelihood.predict_mean_and_var(Fm, Fv)
    np.testing.assert_allclose(npy(m2), g["pred_mean"], rtol=1e-4, atol=1e-6)
    d2 = model.likelihood.predict_density(Fm, Fv, g["Y"])
    np.testing.assert_allclose(npy(d2), g["pred_density"], rtol=1e-4, atol=1e-5)


def test_varexp_matches_oracle_on_random_inputs():
    import deepcgp_b200 as D
    from oracle import dcgp_oracle as O
    rng = np.random.RandomState(0)
    S, N, K = 3, 50, 10
    Fmu = rng.standard_normal((S * N, K)).astype(np.float32) * 2
    Fvar = (rng.random((S * N, K)).astype(np.float32) * 3 + 1e-3)
    Fvar[0, :] = 0.0          # clipped to 1e-10 by the likelihood
    Y = rng.randint(0, K, size=N)
    ve, tot = D.MultiClass(K).variational_expectations(torch.as_tensor(Fmu, device=dev()), torch.as_tensor(Fvar, device=dev()), Y, S=S)
    ref = O.robustmax_varexp(Fmu.astype(np.float64), Fvar.astype(np.float64), np.tile(Y, S), K)
    np.testing.assert_allclose(npy(ve), ref, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(float(tot.item()), ref.sum(), rtol=1e-10)


# ----------------------------------------------------------------------------------------------- oracle, seeded mid-size
def _synthetic_conv(rng, H, W, C, f, s, M, R, white=False, trained=True):
    from oracle import dcgp_oracle as O
    L = f * f * C
    P = ((H - f) // s + 1) * ((W - f) // s + 1)
    Ximg = rng.standard_normal((max(16, 2 * M // P + 1), H, W, C))      # enough distinct patches for M inducing points
    pat = O.extract_patches(Ximg, f, s).reshape(-1, L)
    Z = pat[rng.choice(pat.shape[0], M, replace=False)] + 0.1 * rng.standard_normal((M, L))
    lay = dict(type="conv", H=H, W=W, C=C, f=f, s=s, M=M, R=R, white=white, variance=5.0, lengthscale=5.0, Z=Z)
    if trained:
        lay["q_mu"] = rng.standard_normal((M, R))
        lay["q_sqrt"] = np.tril(rng.standard_normal((R, M, M)) * 0.3) + 0.5 * np.eye(M)
    else:   # models.py:136-138 initial state
        lay["q_mu"] = np.zeros((M, R))
        lay["q_sqrt"] = np.tile(1e-5 * np.linalg.cholesky(O.mo_Kuu(Z, 5.0, 5.0))[None], (R, 1, 1))
    return lay


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("cfg", [
    dict(N=6, H=28, W=28, C=1, f=5, s=2, M=128, R=10, white=False, trained=True),     # cfg2 layer 1
    dict(N=3, H=32, W=32, C=3, f=5, s=2, M=512, R=10, white=False, trained=True),     # cfg3 layer 1
    dict(N=3, H=14, W=14, C=10, f=5, s=1, M=512, R=10, white=False, trained=False),   # cfg3 layer 2, init state
    dict(N=4, H=14, W=14, C=10, f=5, s=1, M=200, R=7, white=True, trained=True),      # ragged M, R; whitened
    dict(N=2, H=16, W=16, C=3, f=5, s=3, M=1024, R=4, white=False, trained=True),     # cfg4-size M
    dict(N=3, H=12, W=12, C=4, f=3, s=1, M=384, R=5, white=False, trained=True),      # M padded to 128 only (stage 1 on BN=128 maps)
    dict(N=3, H=12, W=12, C=4, f=3, s=1, M=300, R=5, white=False, trained=True),      # M padded to 320: single-accumulator stage 1
    dict(N=5, H=9, W=9, C=2, f=3, s=2, M=100, R=3, white=True, trained=True),         # odd row count (T = 80), M padded to 128
])
def test_convlayer_vs_oracle(cfg, algo):
    from oracle import dcgp_oracle as O
    cfg = dict(cfg)
    N, trained = cfg.pop("N"), cfg.pop("trained")
    rng = np.random.RandomState(1234)
    lay = _synthetic_conv(rng, trained=trained, **cfg)
    X = rng.standard_normal((N, cfg["H"] * cfg["W"] * cfg["C"])).astype(np.float32)
    mref, vref = O.convlayer_conditional_ND_fast(X.astype(np.float64), lay)
    layer = build_conv(lay, algo)
    mean, var = layer.conditional_ND(torch.as_tensor(X, device=dev()))
    assert_parity(npy(mean), mref, lay["variance"], "mean")
    assert_parity(npy(var), vref, lay["variance"], "var")
    np.testing.assert_allclose(float(layer.KL().item()), O.convlayer_KL(lay), rtol=max(1e-8, KL_RTOL[algo]), atol=1e-4)


@pytest.mark.parametrize("algo", ALGOS)
def test_ill_conditioned_kuu(algo):
    """1024 inducing patches in a 32-dimensional patch space: cond(Kuu) ~ 1e4.  The reference is float64; the T-sized
    arithmetic here is fp32-class, so the error grows with the cancellation in Lm^-1 k (|Lm^-1||k| / |a| = 14, then x 50
    through alpha^T a).  Both paths meet the 1e-4 gate: the fp32 CUDA-core path at 4.7e-5 / 5e-6, the tensor-core path at
    3.5e-5 / 8e-6 with the first stage spread over four TMEM accumulators (the default, dcgp_set_precise_stage1).  With a single
    accumulator the 3*M/16 round-toward-zero accumulations per output (tools/diag_accum.py: 0.69 ulp low per MMA) are
    amplified by the same cancellation: 1.7e-4 / 3.2e-5, kept here as the documented bound 3e-4 of that setting."""
    from oracle import dcgp_oracle as O
    from deepcgp_b200 import _lib
    from tests.util import parity_err
    rng = np.random.RandomState(1234)
    lay = _synthetic_conv(rng, 12, 12, 2, 4, 3, 1024, 4, trained=True)
    X = rng.standard_normal((2, 12 * 12 * 2)).astype(np.float32)
    mref, vref = O.convlayer_conditional_ND_fast(X.astype(np.float64), lay)
    saved = _lib.lib.dcgp_get_precise_stage1()
    try:
        for precise, bound in ((1, 1e-4),) if algo == "simt" else ((1, 1e-4), (0, 3e-4)):
            _lib.lib.dcgp_set_precise_stage1(precise)
            mean, var = build_conv(lay, algo).conditional_ND(torch.as_tensor(X, device=dev()))
            for got, ref, what in ((mean, mref, "mean"), (var, vref, "var")):
                normwise, _ = parity_err(npy(got), ref, 5.0)
                assert normwise <= bound, "%s (precise=%d): normwise %.3e > %.0e" % (what, precise, normwise, bound)
    finally:
        _lib.lib.dcgp_set_precise_stage1(saved)


@pytest.mark.parametrize("algo", ALGOS)
def test_cfg1_single_layer_model_vs_oracle(algo):
    """BASELINE config 1 shape (28x28x1, f=5, s=1, P=576, M=32): one SVGP(ConvKernel) layer, reduced batch."""
    from oracle import dcgp_oracle as O
    rng = np.random.RandomState(1235)
    H = W = 28
    M, N, S = 32, 8, 3
    Ximg = rng.standard_normal((16, H, W, 1))
    pat = O.extract_patches(Ximg, 5, 1).reshape(-1, 25)
    Z = pat[rng.choice(pat.shape[0], M, replace=False)] + 0.1 * rng.standard_normal((M, 25))
    lay = dict(type="svgp_conv", H=H, W=W, C=1, f=5, s=1, M=M, R=10, white=False, variance=5.0, lengthscale=5.0, Z=Z,
               q_mu=rng.standard_normal((M, 10)), q_sqrt=np.tril(rng.standard_normal((10, M, M)) * 0.3) + 0.5 * np.eye(M),
               patch_weights=np.ones(576))
    X = rng.standard_normal((N, H * W)).astype(np.float32)
    Y = rng.randint(0, 10, size=(N, 1))
    zs = [rng.standard_normal((S, N, 10)).astype(np.float32)]
    ref = O.dgp_elbo([lay], X.astype(np.float64), Y, [z.astype(np.float64) for z in zs], 1000.0, S)
    model = build_model([lay], X, Y, S, 1000.0, algo)
    elbo = model.compute_log_likelihood(X, Y, zs=[torch.as_tensor(z, device=dev()) for z in zs])
    assert abs(elbo - ref) <= 1e-3 * abs(ref), (elbo, ref)


# ----------------------------------------------------------------------------------------------- full-size properties
@pytest.mark.parametrize("algo", ALGOS)
def test_prior_recovery_full_size(algo):
    """SURVEY A.6 (3) at BASELINE cfg3 layer-1 size (M=512, P=196, R=10, 64 images): q(u) = p(u) => mean == 0,
    var == sigma^2 exactly in exact arithmetic, for every patch; KL == 0."""
    from oracle import dcgp_oracle as O
    rng = np.random.RandomState(7)
    lay = _synthetic_conv(rng, 32, 32, 3, 5, 2, 512, 10, trained=True)
    Lm = np.linalg.cholesky(O.mo_Kuu(lay["Z"], 5.0, 5.0))
    lay["q_mu"] = np.zeros((512, 10))
    lay["q_sqrt"] = np.tile(Lm[None], (10, 1, 1))
    layer = build_conv(lay, algo)
    X = torch.randn((64, 32 * 32 * 3), device=dev(), generator=torch.Generator(device=dev()).manual_seed(3))
    mean, var = layer.conditional_ND(X)
    assert mean.shape == (64, 196 * 10)
    assert float(mean.abs().max()) == 0.0 or float(mean.abs().max()) < 1e-6
    assert float((var - 5.0).abs().max()) <= 1e-4 * 5.0 + 1e-4 * 5.0
    assert abs(float(layer.KL().item())) < (1e-6 if algo == "simt" else 5e-3)    # terms of size R*M = 5120 cancel


@pytest.mark.parametrize("algo", ALGOS)
def test_white_nonwhite_equivalence_full_size(algo):
    """SURVEY A.6 (2) at cfg3 layer-2 size: the whitened parameterisation gives the same q(f)."""
    import scipy.linalg as sla
    from oracle import dcgp_oracle as O
    rng = np.random.RandomState(8)
    lay = _synthetic_conv(rng, 14, 14, 10, 5, 1, 512, 10, trained=True)
    Lm = np.linalg.cholesky(O.mo_Kuu(lay["Z"], 5.0, 5.0))
    layw = dict(lay, white=True, q_mu=sla.solve_triangular(Lm, lay["q_mu"], lower=True),
                q_sqrt=np.stack([sla.solve_triangular(Lm, lay["q_sqrt"][r], lower=True) for r in range(10)]))
    X = torch.randn((32, 14 * 14 * 10), device=dev(), generator=torch.Generator(device=dev()).manual_seed(4))
    m1, v1 = build_conv(lay, algo).conditional_ND(X)
    m2, v2 = build_conv(layw, algo).conditional_ND(X)
    assert_parity(npy(m2), npy(m1), 5.0, "mean white vs non-white")
    assert_parity(npy(v2), npy(v1), 5.0, "var white vs non-white")


def test_simt_and_tc_paths_agree_full_size():
    """The tensor-core product path against the fp32 CUDA-core path at cfg3 layer-2 size, 256 images."""
    rng = np.random.RandomState(9)
    lay = _synthetic_conv(rng, 14, 14, 10, 5, 1, 512, 10, trained=True)
    X = torch.randn((256, 14 * 14 * 10), device=dev(), generator=torch.Generator(device=dev()).manual_seed(5))
    m1, v1 = build_conv(lay, "simt").conditional_ND(X)
    m2, v2 = build_conv(lay, "tc").conditional_ND(X)
    assert_parity(npy(m2), npy(m1), 5.0, "mean tc vs simt")
    assert_parity(npy(v2), npy(v1), 5.0, "var tc vs simt")


def test_s_broadcast_equals_tiling():
    """DS/dgp.py:63: the first layer's S identical copies -- computing once and broadcasting must equal tiling."""
    rng = np.random.RandomState(10)
    lay = _synthetic_conv(rng, 12, 12, 1, 5, 2, 64, 3, trained=True)
    layer = build_conv(lay, "simt")
    from deepcgp_b200.layers import TiledInput
    X = torch.randn((5, 144), device=dev())
    z = torch.randn((4, 5, layer.num_outputs), device=dev())
    s1, m1, v1 = layer.sample_from_conditional(TiledInput(X, 4), z=z)
    s2, m2, v2 = layer.sample_from_conditional(X[None].repeat(4, 1, 1), z=z)
    torch.testing.assert_close(m1, m2, rtol=0, atol=0)
    torch.testing.assert_close(v1, v2, rtol=0, atol=0)
    torch.testing.assert_close(s1, s2, rtol=0, atol=0)


# ----------------------------------------------------------------------------------------------- 8e: rank-invariant noise
def test_counter_based_normals_are_shard_invariant_and_standard():
    """dcgp_randn (the stand-in for tf.random_normal, DS/layers.py:104): the draw of an image depends on its GLOBAL index
    only, so a batch sharded over ranks reproduces the unsharded draws bit for bit; moments of N(0,1)."""
    g = load_golden("dgp2_elbo")
    layers = layers_from_golden(g)
    model = build_model(layers, g["X"].astype(np.float32), g["Y"], 3, 100.0, "tc")
    full = model.draw_zs(8, 8, 0, step=5)
    a = model.draw_zs(3, 8, 0, step=5)
    b = model.draw_zs(5, 8, 3, step=5)
    for zf, za, zb in zip(full, a, b):
        assert torch.equal(zf, torch.cat([za, zb], dim=1))
    other = model.draw_zs(8, 8, 0, step=6)
    assert not torch.equal(full[0], other[0])
    from deepcgp_b200 import _lib
    z = torch.empty((4, 1000, 500), dtype=torch.float32, device=dev())
    _lib.check(_lib.lib.dcgp_randn(_lib.ptr(z), 4, 1000, 500, 1000, 0, 12345, 1, 0, _lib.stream()))
    zz = z.double()
    assert abs(float(zz.mean())) < 3e-3
    assert abs(float(zz.var()) - 1.0) < 5e-3
    assert abs(float((zz ** 4).mean()) - 3.0) < 5e-2
    assert float(zz.abs().max()) < 7.0 and bool(torch.isfinite(z).all())
    # successive samples / images / outputs are uncorrelated
    assert abs(float((zz[:, :, :-1] * zz[:, :, 1:]).mean())) < 3e-3
    assert abs(float((zz[:, :-1] * zz[:, 1:]).mean())) < 3e-3
    assert abs(float((zz[:-1] * zz[1:]).mean())) < 3e-3


def test_elbo_data_term_is_additive_over_image_shards():
    """ELBO = scale * sum_n l_n - sum KL, so the data terms of two image shards (with the draws of their GLOBAL
    image indices) add up to the unsharded one."""
    g = load_golden("dgp3_elbo")
    layers = layers_from_golden(g)
    S = int(g["S"])
    X32 = g["X"].astype(np.float32)
    N = X32.shape[0]
    Y = np.asarray(g["Y"]).reshape(-1)
    model = build_model(layers, X32, g["Y"], S, float(g["num_data"]), "tc")
    zs = model.draw_zs(N, N, 0, step=1)
    full = model.compute_log_likelihood(X32, Y, zs=zs)
    kl = float(model._kls.sum().item())
    parts = 0.0
    for lo, hi in ((0, 1), (1, N)):
        zsh = model.draw_zs(hi - lo, N, lo, step=1)
        e = float(model._build_likelihood(X32[lo:hi], Y[lo:hi], zs=zsh, n_global=N).item())
        parts += e + kl
    assert abs((full + kl) - parts) <= 1e-6 * abs(full + kl), (full, parts, kl)
