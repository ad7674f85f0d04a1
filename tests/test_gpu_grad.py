"""GPU: gradient of the ELBO (a10) -- the CUDA backward (dcgp_layer_backward + M-only chain rule) against torch.autograd
of the float64 oracle (oracle/dcgp_oracle_torch.py), on the golden models generated from the reference source.

Tolerance: gradients are sums over all patches of products of fp32-class quantities; gate = 2e-3 of the tensor's max
magnitude (the ELBO itself is gated at 1e-3)."""
import numpy as np
import pytest
import torch

from tests.test_gpu_parity import build_model, dev, npy
from tests.util import golden_names, layers_from_golden, load_golden

pytestmark = pytest.mark.gpu


def _check(name, got, ref, tol=2e-3, floor=1e-5):
    """max|got - ref| <= tol * max|ref| + floor.  The floor only matters for gradients that vanish analytically (e.g. the
    patch-weight gradient when every class has the same predictive variance: sum_r dELBO/dvar_r = 0 exactly, while the
    float32 upstream gradients, each of size ~ num_data/N, cancel only to ~1e-7 relative); callers scale it with num_data/N."""
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    scale = np.max(np.abs(ref))
    err = np.max(np.abs(got.reshape(ref.shape) - ref))
    assert err <= tol * scale + floor, "%s: max|diff| = %.3e, max|ref| = %.3e" % (name, err, scale)


@pytest.mark.parametrize("name", golden_names("dgp"))
def test_elbo_gradient_vs_autograd_oracle(name):
    import deepcgp_b200 as D
    from oracle import dcgp_oracle_torch as OT
    g = load_golden(name)
    layers = layers_from_golden(g)
    S = int(g["S"])
    X32 = g["X"].astype(np.float32)
    zs32 = [g["z%d" % i].astype(np.float32) for i in range(len(layers))]
    ref_elbo, ref_grads = OT.elbo_and_grads(layers, X32.astype(np.float64), g["Y"], [z.astype(np.float64) for z in zs32],
                                            float(g["num_data"]), S)
    model = build_model(layers, X32, g["Y"], S, float(g["num_data"]), "tc")
    eg = D.ElboGradient(model)
    elbo, grads = eg(X32, g["Y"], zs=[torch.as_tensor(z, device=dev()) for z in zs32])
    elbo = float(elbo.item())
    assert abs(elbo - ref_elbo) <= 1e-3 * abs(ref_elbo)
    for i, (got, ref) in enumerate(zip(grads, ref_grads)):
        for k, v in ref.items():
            _check("layer %d %s" % (i, k), npy(got[k]), v, floor=2e-8 * float(g["num_data"]) / X32.shape[0])


def test_varexp_grad_matches_autograd():
    from deepcgp_b200 import _lib
    from oracle import dcgp_oracle_torch as OT
    rng = np.random.RandomState(0)
    S, N, K = 2, 40, 10
    Fmu = (rng.standard_normal((S * N, K)) * 2).astype(np.float32)
    Fvar = (rng.random((S * N, K)) * 3 + 1e-2).astype(np.float32)
    Y = rng.randint(0, K, size=N)
    tm = torch.tensor(Fmu.astype(np.float64), requires_grad=True)
    tv = torch.tensor(Fvar.astype(np.float64), requires_grad=True)
    (0.37 * OT.robustmax_varexp(tm, tv, np.tile(Y, S), K).sum()).backward()
    d = dev()
    gm = torch.empty((S * N, K), dtype=torch.float32, device=d)
    gv = torch.empty_like(gm)
    Yd = torch.as_tensor(Y, device=d).to(torch.int32)
    dmu, dvar = torch.as_tensor(Fmu, device=d), torch.as_tensor(Fvar, device=d)     # keep alive across the async call
    _lib.check(_lib.lib.dcgp_multiclass_varexp_grad(_lib.ptr(dmu), _lib.ptr(dvar), _lib.ptr(Yd), S, N, K, 1e-3, 0.37,
                                                    _lib.ptr(gm), _lib.ptr(gv), _lib.stream()))
    _check("gmu", npy(gm), tm.grad.numpy(), 1e-5)
    _check("gvar", npy(gv), tv.grad.numpy(), 1e-5)


def test_adam_step_matches_reference_update_rule():
    from deepcgp_b200 import _lib
    rng = np.random.RandomState(1)
    n = 1000
    p, g = rng.standard_normal(n), rng.standard_normal(n)
    m, v = np.zeros(n), np.zeros(n)
    d = dev()
    tp, tg, tm, tv = (torch.as_tensor(a.copy(), device=d) for a in (p, g, m, v))
    lr, b1, b2, eps = 0.01, 0.9, 0.999, 1e-8
    for step in (1, 2, 3):
        _lib.check(_lib.lib.dcgp_adam(_lib.ptr(tp), _lib.ptr(tg), _lib.ptr(tm), _lib.ptr(tv), n, lr, b1, b2, eps, step, 0, _lib.stream()))
        m = b1 * m + (1 - b1) * g
        v = b2 * v + (1 - b2) * g * g
        p = p - lr * np.sqrt(1 - b2 ** step) / (1 - b1 ** step) * m / (np.sqrt(v) + eps)     # tf.train.AdamOptimizer
    np.testing.assert_allclose(npy(tp), p, rtol=1e-12, atol=1e-14)


def test_training_steps_increase_the_elbo():
    """A few Adam steps on a fixed minibatch with fixed noise must raise the ELBO (sanity of signs end to end)."""
    import deepcgp_b200 as D
    g = load_golden("dgp2_elbo")
    layers = layers_from_golden(g)
    S = int(g["S"])
    X32 = g["X"].astype(np.float32)
    model = build_model(layers, X32, g["Y"], S, float(g["num_data"]), "tc")
    zs = [torch.as_tensor(g["z%d" % i].astype(np.float32), device=dev()) for i in range(len(layers))]
    eg, opt = D.ElboGradient(model), D.Adam(model, lr=0.01)
    elbos = []
    for _ in range(6):
        elbo, grads = eg(X32, g["Y"], zs=zs)
        elbos.append(float(elbo.item()))
        opt.step(grads)
    assert elbos[-1] > elbos[0], elbos


def test_elbo_gradient_cfg2_shape_vs_autograd_oracle():
    """BASELINE config 2 shapes (MNIST 2-layer DCGP: 28x28x1, f=5 s=2 -> 12x12x10, f=5 s=1; M=128,128; R=10), reduced batch:
    ELBO and every parameter gradient against float64 autograd of the oracle."""
    import bench
    import deepcgp_b200 as D
    from oracle import dcgp_oracle_torch as OT
    cfg = dict(bench.CONFIGS["cfg2"])
    layers = bench.synth_params(cfg, seed=77)
    rng = np.random.RandomState(5)
    N, S = 4, 3
    X32 = rng.standard_normal((N, 28 * 28)).astype(np.float32)
    Y = rng.randint(0, 10, size=(N, 1))
    dims = [144 * 10, 10]
    zs32 = [rng.standard_normal((S, N, d)).astype(np.float32) for d in dims]
    ref_elbo, ref_grads = OT.elbo_and_grads(layers, X32.astype(np.float64), Y, [z.astype(np.float64) for z in zs32], 60000.0, S)
    model = build_model(layers, X32, Y, S, 60000.0, "tc")
    elbo, grads = D.ElboGradient(model)(X32, Y, zs=[torch.as_tensor(z, device=dev()) for z in zs32])
    assert abs(float(elbo.item()) - ref_elbo) <= 1e-3 * abs(ref_elbo)
    for i, (got, ref) in enumerate(zip(grads, ref_grads)):
        for k, v in ref.items():
            _check("layer %d %s" % (i, k), npy(got[k]), v, floor=2e-8 * 60000.0 / N)


@pytest.mark.parametrize("name", ["dgp2_elbo", "dgp3_elbo"])
def test_pipelined_train_step_equals_sequential_steps(name):
    """grad.TrainStep (split backward, per-layer Adam + prepare on side streams, lazy hyper-parameter read-back) must walk
    the same parameter trajectory as ElboGradient + Adam.step: same ELBOs and same parameters after 4 steps."""
    import deepcgp_b200 as D
    if name not in golden_names("dgp"):
        pytest.skip("golden model %s not present" % name)
    g = load_golden(name)
    S = int(g["S"])
    X32 = g["X"].astype(np.float32)

    def fresh():
        layers = layers_from_golden(g)
        model = build_model(layers, X32, g["Y"], S, float(g["num_data"]), "tc")
        zs = [torch.as_tensor(g["z%d" % i].astype(np.float32), device=dev()) for i in range(len(layers))]
        return model, zs

    m1, zs1 = fresh()
    eg, opt = D.ElboGradient(m1), D.Adam(m1, lr=0.01)
    e1 = []
    for _ in range(4):
        elbo, grads = eg(X32, g["Y"], zs=zs1)
        e1.append(float(elbo.item()))
        opt.step(grads)
    m2, zs2 = fresh()
    step = D.TrainStep(m2, lr=0.01)
    e2 = [float(step(X32, g["Y"], zs=zs2).item()) for _ in range(4)]
    step.finish()
    torch.cuda.synchronize()
    np.testing.assert_allclose(e2, e1, rtol=1e-7)
    np.testing.assert_allclose(npy(step.opt.flat), npy(opt.flat), rtol=1e-6, atol=1e-9)
    for l1, l2 in zip(m1.layers, m2.layers):
        assert abs(l1._base_kernel.variance - l2._base_kernel.variance) <= 1e-9 * abs(l1._base_kernel.variance)
        assert abs(l1._base_kernel.lengthscales - l2._base_kernel.lengthscales) <= 1e-9 * abs(l1._base_kernel.lengthscales)


def test_model_builder_checkpoint_roundtrip(tmp_path):
    """ModelBuilder (models.py:35-198: k-means patch init, identity_conv shape propagation, q_sqrt x 1e-5),
    a few optimisation steps, experiment.py:56-64 parameter dump, and a rebuild from it (models.py:200-233)."""
    import deepcgp_b200 as D
    from deepcgp_b200 import models as Mo
    rng = np.random.RandomState(0)
    X = rng.standard_normal((40, 12, 12, 1))
    Y = rng.randint(0, 10, size=(40, 1))
    flags = Mo.default_parser().parse_args(["-M", "8,8", "--feature-maps", "3", "--filter-sizes", "5,3", "--strides", "2,1",
                                            "--batch-size", "8", "--num-samples", "2"])
    model = Mo.ModelBuilder(flags, X, Y, device=dev()).build()
    assert [type(l).__name__ for l in model.layers] == ["ConvLayer", "SVGP_Layer"]
    l0, l1 = model.layers
    assert l0.feature.Z.shape == (8, 25) and l0.num_outputs == 16 * 3 and l1.feature.Z.shape == (8, 9 * 3)
    assert float(l0._base_kernel.variance) == 5.0 and float(l0._base_kernel.lengthscales) == 5.0        # models.py:115-116
    Ku = D.MultiOutputConvKernel(l0.base_kernel, 144, 16).Kuu(l0.feature.Z)
    np.testing.assert_allclose(npy(l0.q_sqrt[0]), 1e-5 * np.linalg.cholesky(npy(Ku)), rtol=1e-6, atol=1e-12)
    step = D.TrainStep(model, lr=0.01)
    Xb = X[:8].reshape(8, -1).astype(np.float32)
    for _ in range(3):
        step(Xb, Y[:8])
    step.finish()
    torch.cuda.synchronize()
    path = str(tmp_path / "model.npy")
    Mo.save_model_parameters(model, path, global_step=3)
    saved = np.load(path, allow_pickle=True).item()
    assert saved["global_step"] == 3 and "DGP/layers/1/kern/patch_weights" in saved and "DGP/layers/0/feature/Z" in saved
    flags2 = Mo.default_parser().parse_args(["-M", "8,8", "--feature-maps", "3", "--filter-sizes", "5,3", "--strides", "2,1",
                                             "--batch-size", "8", "--num-samples", "2", "--load-model", "model"])
    builder = Mo.ModelBuilder(flags2, X, Y, model_path=path, device=dev())
    model2 = builder.build()
    assert builder.global_step == 3
    zs = model.draw_zs(8, 8, 0, step=11)
    m1, v1 = model.predict_f(Xb, 2, zs=zs)
    m2, v2 = model2.predict_f(Xb, 2, zs=zs)
    np.testing.assert_allclose(npy(m2), npy(m1), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(npy(v2), npy(v1), rtol=1e-5, atol=1e-6)


def test_experiment_driver_runs_and_logs(tmp_path):
    """experiment.py:28-64 for --optimizer Adam: `test_every` iterations with the staircase learning rate, the accuracy
    logger (utils/log.py:55-68) and the parameter dump."""
    import argparse
    import deepcgp_b200 as D
    rng = np.random.RandomState(1)
    X = rng.standard_normal((48, 12, 12, 1))
    Y = rng.randint(0, 10, size=(48, 1))
    flags = D.models.default_parser().parse_args(["-M", "8,8", "--feature-maps", "3", "--filter-sizes", "5,3", "--strides",
                                                  "2,1", "--batch-size", "8", "--num-samples", "2", "--lr", "0.01"])
    flags.name, flags.log_dir, flags.lr_decay_steps, flags.test_every, flags.optimizer = "exp", str(tmp_path), 4, 3, "Adam"
    exp = D.Experiment(flags, X, Y, X_test=X[:40], Y_test=Y[:40], device=dev())
    e1 = exp.train_step()
    e2 = exp.train_step()
    assert (e1["global_step"], e2["global_step"]) == (3, 6)
    assert e1["lr"] == 0.01 and abs(e2["lr"] - 0.001) < 1e-12          # staircase: x0.1 after 4 steps
    assert 0.0 <= e2["test_accuracy"] <= 1.0 and np.isfinite(e2["elbo"])
    saved = np.load(str(tmp_path / "exp.npy"), allow_pickle=True).item()
    assert saved["global_step"] == 6
    np.testing.assert_allclose(saved["DGP/layers/0/q_mu"], npy(exp.model.layers[0].q_mu))
