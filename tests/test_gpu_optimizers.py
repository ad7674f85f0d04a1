"""GPU: SURVEY.md 8 row f4 -- the NatGrad + Adam hybrid with its gamma schedule and Cholesky-failure back-off
(conv_gp/experiment.py:38-49,71-99) and the fixed Conv2dMean mean function (conv_gp/mean_functions.py:28-41, `--identity-mean`)."""
import numpy as np
import pytest
import torch

from tests.test_gpu_parity import build_model, dev, npy
from tests.util import layers_from_golden, load_golden

pytestmark = pytest.mark.gpu


def _flags(tmp_path, **kw):
    import deepcgp_b200 as D
    flags = D.models.default_parser().parse_args(["-M", "8,8", "--feature-maps", "3", "--filter-sizes", "5,3", "--strides", "2,1",
                                                  "--batch-size", "8", "--num-samples", "2", "--lr", "0.01"])
    flags.name, flags.log_dir, flags.lr_decay_steps, flags.test_every = "exp", str(tmp_path), 100, 3
    for k, v in kw.items():
        setattr(flags, k, v)
    return flags


def test_natural_gradient_step_with_gamma_one_is_exact_for_a_conjugate_objective():
    """For L(mu, S) = KL[N(mu, S) || N(m0, S0)] (quadratic in the natural parameters) one natural-gradient step with gamma = 1
    lands on the optimum (mu, S) = (m0, S0): checks the eta / theta algebra and the Cholesky backward rule of grad.NatGrad."""
    from deepcgp_b200.grad import NatGrad
    rng = np.random.RandomState(0)
    M, R = 6, 2
    A = rng.standard_normal((M, M))
    S0 = torch.tensor(A @ A.T + M * np.eye(M))
    m0 = torch.tensor(rng.standard_normal((M, R)))

    class L(object):
        pass
    layer = L()
    layer.q_mu = torch.tensor(rng.standard_normal((M, R)))
    layer.q_sqrt = torch.tensor(np.tril(rng.standard_normal((R, M, M)) * 0.3) + np.eye(M)).contiguous()
    layer._fresh = False
    model = L()
    model.layers = [layer]
    q_mu = layer.q_mu.clone().requires_grad_(True)
    q_sqrt = layer.q_sqrt.clone().requires_grad_(True)
    S0inv = torch.linalg.inv(S0)
    kl = 0.0
    for r in range(R):
        Lr = torch.tril(q_sqrt[r])
        S = Lr @ Lr.T
        d = q_mu[:, r] - m0[:, r]
        kl = kl + 0.5 * (torch.trace(S0inv @ S) + d @ S0inv @ d - M + torch.logdet(S0) - torch.logdet(S))
    g_mu, g_sq = torch.autograd.grad(-kl, [q_mu, q_sqrt])          # "ELBO" = -KL
    NatGrad(model).step([{"q_mu": g_mu, "q_sqrt": torch.tril(g_sq)}], gamma=1.0)
    np.testing.assert_allclose(layer.q_mu.numpy(), m0.numpy(), rtol=1e-9, atol=1e-10)
    for r in range(R):
        np.testing.assert_allclose((layer.q_sqrt[r] @ layer.q_sqrt[r].T).numpy(), S0.numpy(), rtol=1e-9, atol=1e-10)


def test_natgrad_raises_when_the_step_leaves_the_cone():
    from deepcgp_b200 import _lib
    from deepcgp_b200.grad import NatGrad

    class L(object):
        pass
    M, R = 4, 1
    layer = L()
    layer.q_mu = torch.zeros((M, R), dtype=torch.float64)
    layer.q_sqrt = torch.eye(M, dtype=torch.float64)[None].contiguous()
    layer._fresh = False
    model = L()
    model.layers = [layer]
    before = layer.q_sqrt.clone()
    g = {"q_mu": torch.zeros((M, R), dtype=torch.float64), "q_sqrt": 50.0 * torch.eye(M, dtype=torch.float64)[None]}
    with pytest.raises(_lib.NotPositiveDefiniteError):
        NatGrad(model).step([g], gamma=1.0)      # precision 1 - 2*gamma*25 < 0
    assert torch.equal(layer.q_sqrt, before)     # parameters untouched on failure


@pytest.mark.parametrize("optimizer", ["NatGrad", "SGD"])
def test_experiment_driver_other_optimizers(tmp_path, optimizer):
    """experiment.py:84-108: NatGrad on (q_mu, q_sqrt) + Adam on the rest; SGD.  A few iterations raise the ELBO on a fixed
    small problem and leave finite parameters; the variational parameters are frozen for Adam under NatGrad."""
    import deepcgp_b200 as D
    rng = np.random.RandomState(1)
    X = rng.standard_normal((48, 12, 12, 1))
    Y = rng.randint(0, 10, size=(48, 1))
    flags = _flags(tmp_path, optimizer=optimizer, gamma=0.01, lr=0.01 if optimizer == "NatGrad" else 1e-6)
    exp = D.Experiment(flags, X, Y, X_test=X[:16], Y_test=Y[:16], device=dev())
    q0 = [l.q_mu.clone() for l in exp.model.layers]
    e = exp.train_step()
    assert e["global_step"] == 3 and np.isfinite(e["elbo"])
    if optimizer == "NatGrad":
        assert abs(exp.gamma() - min((3 / 100.0 * 1e-3 + 0.01), 1.0)) < 1e-12
        assert any(float((l.q_mu - q).abs().max()) > 0 for l, q in zip(exp.model.layers, q0))   # moved by the natural gradient
        assert float(exp.opt.m[exp.opt.layer_range(0)[0]:exp.opt.layer_range(0)[1]].abs().max()) > 0   # Adam moved the rest
    for l in exp.model.layers:
        assert bool(torch.isfinite(l.q_mu).all()) and bool(torch.isfinite(l.q_sqrt).all())


def test_natgrad_backoff_shrinks_gamma_and_retries(tmp_path, monkeypatch):
    """experiment.py:38-49: a natural-gradient step that fails is retried with gamma * 0.2 (steps_back += 1), at most five times;
    the failure itself (Cholesky of the new precision) is covered by test_natgrad_raises_when_the_step_leaves_the_cone."""
    import deepcgp_b200 as D
    from deepcgp_b200 import _lib
    from deepcgp_b200.grad import NatGrad
    rng = np.random.RandomState(2)
    X = rng.standard_normal((48, 12, 12, 1))
    Y = rng.randint(0, 10, size=(48, 1))
    flags = _flags(tmp_path, optimizer="NatGrad", gamma=1.0)
    flags.test_every = 1
    exp = D.Experiment(flags, X, Y, device=dev())
    real_step, seen = NatGrad.step, []

    def step(self, grads, gamma):
        seen.append(gamma)
        if gamma > 0.1:              # stands for tf.errors.InvalidArgumentError from the failed Cholesky
            raise _lib.NotPositiveDefiniteError("step too long")
        return real_step(self, grads, gamma)

    monkeypatch.setattr(NatGrad, "step", step)
    e = exp.train_step()
    assert seen == pytest.approx([1.0, 0.2, 0.04]) and exp.steps_back == 2 and e["global_step"] == 1
    assert all(bool(torch.isfinite(l.q_sqrt).all()) for l in exp.model.layers)
    monkeypatch.setattr(NatGrad, "step", lambda self, grads, gamma: (_ for _ in ()).throw(_lib.NotPositiveDefiniteError("always")))
    with pytest.raises(_lib.NotPositiveDefiniteError):
        exp.train_step()             # five retries, then the error propagates (experiment.py:41-42)


def test_conv2dmean_layer_forward_and_input_gradient():
    """ConvLayer with Conv2dMean: mean/sample shifted by the centre tap of input map 0 (mean_functions.py:28-41), and the input
    gradient gains the matching scatter."""
    import deepcgp_b200 as D
    from deepcgp_b200.grad import LayerBackward
    g = load_golden("convlayer_a")
    from tests.util import layer_from_golden
    lay = layer_from_golden(g, 0, "conv")
    from tests.test_gpu_parity import build_conv
    base = build_conv(lay, "tc")
    view = D.FullView((lay["H"], lay["W"]), lay["f"], lay["C"], lay["s"])
    kern = D.RBF(lay["f"] ** 2 * lay["C"], variance=lay["variance"], lengthscales=lay["lengthscale"])
    mf = D.Conv2dMean(lay["f"], lay["C"], lay["R"], stride=lay["s"])
    layer = D.ConvLayer(kern, mf, feature=D.PatchInducingFeatures(lay["Z"]), view=view, white=lay["white"], gp_count=lay["R"],
                        q_mu=lay["q_mu"], q_sqrt=lay["q_sqrt"], device=dev())
    X = torch.as_tensor(g["X"].astype(np.float32), device=dev())
    m0, v0 = base.conditional_ND(X)
    m1, v1 = layer.conditional_ND(X)
    N = X.shape[0]
    img = X.reshape(N, lay["H"], lay["W"], lay["C"])
    OH, OW, c, s = view.out_image_height, view.out_image_width, lay["f"] // 2, lay["s"]
    expect = torch.zeros((N, OH * OW, lay["R"]), device=dev())
    expect[:, :, 0] = img[:, c:c + (OH - 1) * s + 1:s, c:c + (OW - 1) * s + 1:s, 0].reshape(N, -1)
    torch.testing.assert_close(m1 - m0, expect.reshape(N, -1), rtol=0, atol=1e-6)
    torch.testing.assert_close(v1, v0, rtol=0, atol=0)
    gm = torch.randn_like(m1)
    gv = torch.randn_like(v1)
    outs = []
    for lyr in (base, layer):
        lyr.prepare(); lyr._hold = True; lyr._conditional(X)
        outs.append(LayerBackward(lyr).t_sized(X, 1, gm, gv, True).clone())
        lyr._hold = False
    torch.testing.assert_close(outs[1] - outs[0], mf.backward(gm, lay["H"], lay["W"]), rtol=0, atol=1e-5)
