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


def _check(name, got, ref, tol=2e-3):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    scale = max(np.max(np.abs(ref)), 1e-12)
    err = np.max(np.abs(got.reshape(ref.shape) - ref)) / scale
    assert err <= tol, "%s: max|diff|/max|ref| = %.3e (ref max %.3e)" % (name, err, scale)


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
            _check("layer %d %s" % (i, k), npy(got[k]), v)


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
