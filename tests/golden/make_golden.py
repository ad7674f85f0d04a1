#!/usr/bin/env python
"""Generate tests/golden/*.npz by executing the reference's OWN source files.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

What runs: conv_gp/{views,layers,conditionals,kernels}.py and
submodules/Doubly-Stochastic-DGP/doubly_stochastic_dgp/{layers,dgp,utils}.py, imported unmodified
from /root/reference.  TensorFlow and GPflow are not installable here, so the two packages those files
import are provided by oracle/refshim (numpy/float64 eager restatements of the individual TF ops and
GPflow classes, from their published semantics).  The reference's own call sequence, axis orders,
reshapes, jitter placement and KL/ELBO assembly are therefore exactly the reference's.

Every file stores the inputs, the hyper-parameters and the reference outputs; tests compare
(1) oracle/dcgp_oracle.py and (2) the CUDA path against them.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("DCGP_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))
sys.path.insert(0, os.path.join(REF, "submodules", "Doubly-Stochastic-DGP"))
sys.path.insert(0, os.path.join(REF, "conv_gp"))

import numpy as np  # noqa: E402
import tensorflow as tf  # noqa: E402  (the shim)
import gpflow  # noqa: E402  (the shim)

assert "refshim" in tf.__file__ and "refshim" in gpflow.__file__

from views import FullView  # noqa: E402  reference conv_gp/views.py
from layers import ConvLayer, MultiOutputConvKernel  # noqa: E402  reference conv_gp/layers.py
from conditionals import conditional  # noqa: E402  reference conv_gp/conditionals.py
from kernels import ConvKernel, PatchInducingFeatures  # noqa: E402  reference conv_gp/kernels.py
from doubly_stochastic_dgp.layers import SVGP_Layer  # noqa: E402
from doubly_stochastic_dgp.dgp import DGP_Base  # noqa: E402

for mod in (sys.modules["views"], sys.modules["layers"], sys.modules["conditionals"], sys.modules["kernels"],
            sys.modules["doubly_stochastic_dgp.layers"], sys.modules["doubly_stochastic_dgp.dgp"]):
    assert mod.__file__.startswith(REF), mod.__file__


def rand_q(rng, M, R, scale=0.3):
    q_mu = rng.standard_normal((M, R))
    q_sqrt = np.tril(rng.standard_normal((R, M, M)) * scale) + 0.5 * np.eye(M)[None]
    return q_mu, q_sqrt


def make_conv_layer(rng, H, W, C, f, s, M, R, white, variance, lengthscale, X_for_Z):
    view = FullView(input_size=(H, W), filter_size=f, feature_maps=C, stride=s)
    L = f * f * C
    # inducing patches = seeded sample of real patches + noise (stand-in for k-means, kernels.py:147-164)
    pat = np.asarray(view.extract_patches(tf.constant(X_for_Z.reshape(-1, H, W, C)))).reshape(-1, L)
    Z = pat[rng.choice(pat.shape[0], M, replace=pat.shape[0] < M)] + 0.1 * rng.standard_normal((M, L))
    q_mu, q_sqrt = rand_q(rng, M, R)
    kern = gpflow.kernels.RBF(L, variance=variance, lengthscales=lengthscale)
    layer = ConvLayer(base_kernel=kern, mean_function=gpflow.mean_functions.Zero(),
                      feature=PatchInducingFeatures(Z), view=view, white=white, gp_count=R,
                      q_mu=q_mu, q_sqrt=q_sqrt)
    meta = dict(H=H, W=W, C=C, f=f, s=s, M=M, R=R, white=int(white), variance=variance,
                lengthscale=lengthscale, Z=Z, q_mu=q_mu, q_sqrt=np.tril(q_sqrt))
    return layer, meta


def make_last_layer(rng, H, W, C, f, s, M, R, white, variance, lengthscale, X_for_Z, weights=True):
    view = FullView(input_size=(H, W, C), filter_size=f, feature_maps=C, stride=s)   # models.py:173
    L = f * f * C
    pat = np.asarray(view.extract_patches(tf.constant(X_for_Z.reshape(-1, H, W, C)))).reshape(-1, L)
    Z = pat[rng.choice(pat.shape[0], M, replace=pat.shape[0] < M)] + 0.1 * rng.standard_normal((M, L))
    q_mu, q_sqrt = rand_q(rng, M, R)
    w = 0.5 + rng.random(view.patch_count) if weights else None
    kern = ConvKernel(base_kernel=gpflow.kernels.RBF(L, variance=variance, lengthscales=lengthscale),
                      view=view, patch_weights=w)
    layer = SVGP_Layer(kern=kern, num_outputs=R, feature=PatchInducingFeatures(Z),
                       mean_function=gpflow.mean_functions.Zero(output_dim=R), white=white,
                       q_mu=q_mu, q_sqrt=q_sqrt)
    meta = dict(H=H, W=W, C=C, f=f, s=s, M=M, R=R, white=int(white), variance=variance,
                lengthscale=lengthscale, Z=Z, q_mu=q_mu, q_sqrt=np.tril(q_sqrt),
                patch_weights=np.ones(view.patch_count) if w is None else w)
    return layer, meta


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrs.items()})
    print("wrote", path, os.path.getsize(path), "bytes")


def flat(prefix, meta):
    return {prefix + k: v for k, v in meta.items()}


def case_convlayer(name, seed, N, H, W, C, f, s, M, R, white, variance=1.3, lengthscale=2.1):
    rng = np.random.RandomState(seed)
    X = rng.standard_normal((N, H * W * C))
    layer, meta = make_conv_layer(rng, H, W, C, f, s, M, R, white, variance, lengthscale, X)
    NHWC = tf.constant(X.reshape(N, H, W, C))
    PNL = layer.view.extract_patches_PNL(NHWC)
    NPL = layer.view.extract_patches(NHWC)
    Kuu = layer.conv_kernel.Kuu(layer.feature.Z)
    Kuf = layer.conv_kernel.Kuf(layer.feature.Z, PNL)
    Knn = layer.conv_kernel.Kdiag(PNL)
    fmean, fvar = conditional(Kuf, Kuu, Knn, layer.q_mu, full_cov=False, q_sqrt=layer.q_sqrt, white=white)
    mean, var = layer.conditional_ND(tf.constant(X))
    # perturb Z after construction so that KL's frozen prior (layers.py:149-150) differs from Kuu(Z)
    KL_same = layer.KL()
    save(name, X=X, PNL=PNL, NPL=NPL, Kuu=Kuu, Kuf=Kuf, Knn=Knn, fmean=fmean, fvar=fvar, mean=mean, var=var,
         KL=KL_same, patch_count=layer.view.patch_count, patch_length=layer.view.patch_length,
         out_h=layer.view.out_image_height, out_w=layer.view.out_image_width, jitter=gpflow.settings.jitter,
         **flat("l0_", meta))


def case_lastlayer(name, seed, N, H, W, C, f, s, M, R, white, variance=0.9, lengthscale=1.7):
    rng = np.random.RandomState(seed)
    X = rng.standard_normal((N, H * W * C))
    layer, meta = make_last_layer(rng, H, W, C, f, s, M, R, white, variance, lengthscale, X)
    Xt = tf.constant(X)
    Kzx = layer.kern.Kzx(layer.feature.Z, Xt)
    Kdiag = layer.kern.Kdiag(Xt)
    Kzz = layer.kern.Kzz(layer.feature.Z)
    mean, var = layer.conditional_ND(Xt)
    KL = layer.KL()
    save(name, X=X, Kzx=Kzx, Kdiag=Kdiag, Kzz=Kzz, mean=mean, var=var, KL=KL, jitter=gpflow.settings.jitter,
         **flat("l0_", meta))


def case_dgp(name, seed, N, S, num_data, H, W, C, conv_specs, last_spec, white=False):
    """conv_specs: list of (f, s, M, R); last_spec: (f, s, M). 10 classes."""
    rng = np.random.RandomState(seed)
    X = rng.standard_normal((N, H * W * C))
    Y = rng.randint(0, 10, size=(N, 1)).astype(np.float64)
    layers, metas = [], []
    h, w, c = H, W, C
    Xz = X
    for (f, s, M, R) in conv_specs:
        layer, meta = make_conv_layer(rng, h, w, c, f, s, M, R, white, 5.0, 5.0, Xz)
        layers.append(layer), metas.append(meta)
        h, w, c = layer.view.out_image_height, layer.view.out_image_width, R
        Xz = rng.standard_normal((N, h * w * c))
    f, s, M = last_spec
    layer, meta = make_last_layer(rng, h, w, c, f, s, M, 10, white, 5.0, 5.0, Xz)
    layers.append(layer), metas.append(meta)

    model = DGP_Base(X, Y, likelihood=gpflow.likelihoods.MultiClass(10), layers=layers,
                     num_samples=S, num_data=num_data, minibatch_size=None, name="DGP")
    tf.set_random_seed(seed + 1)
    elbo = model._build_likelihood()            # draws z via tf.random_normal (recorded by the shim)
    zs = [np.array(z) for z in tf.random_state.draws]
    assert len(zs) == len(layers)
    Fs, Fmeans, Fvars = model.propagate(tf.constant(X), full_cov=False, S=S, zs=[tf.constant(z) for z in zs])
    L = tf.reduce_sum(model.likelihood.variational_expectations(Fmeans[-1], Fvars[-1], model.Y))
    KLs = [lay.KL() for lay in layers]
    # prediction path (DS/dgp.py:116-126) on the same samples: BroadcastingLikelihood.predict_mean_and_var / predict_density
    pmean, pvar = model.likelihood.predict_mean_and_var(Fmeans[-1], Fvars[-1])
    pdens = model.likelihood.predict_density(Fmeans[-1], Fvars[-1], model.Y)
    logdens = tf.reduce_logsumexp(pdens - np.log(float(S)), axis=0)
    arrs = dict(X=X, Y=Y, S=S, num_data=num_data, n_layers=len(layers), elbo=elbo, KLs=np.array(KLs),
                varexp=model.likelihood.variational_expectations(Fmeans[-1], Fvars[-1], model.Y),
                pred_mean=pmean, pred_var=pvar, pred_density=pdens, pred_logdensity=logdens,
                jitter=gpflow.settings.jitter)
    for i, (m, z) in enumerate(zip(metas, zs)):
        arrs.update(flat("l%d_" % i, m))
        arrs["z%d" % i] = z
        arrs["F%d" % i] = Fs[i]
        arrs["Fmean%d" % i] = Fmeans[i]
        arrs["Fvar%d" % i] = Fvars[i]
    save(name, **arrs)


if __name__ == "__main__":
    np.random.seed(0)
    case_convlayer("convlayer_a", 11, N=3, H=8, W=8, C=2, f=3, s=1, M=6, R=2, white=False)
    case_convlayer("convlayer_b_white_stride2", 12, N=2, H=9, W=11, C=1, f=4, s=2, M=5, R=3, white=True)
    case_convlayer("convlayer_c_m64", 13, N=2, H=10, W=10, C=3, f=5, s=2, M=64, R=4, white=False,
                   variance=5.0, lengthscale=5.0)
    case_lastlayer("lastlayer_a", 21, N=5, H=6, W=6, C=3, f=3, s=1, M=7, R=10, white=False)
    case_lastlayer("lastlayer_b_white", 22, N=4, H=7, W=5, C=2, f=3, s=2, M=6, R=10, white=True)
    # cfg1-like: a single SVGP(ConvKernel) layer
    case_dgp("dgp1_elbo", 31, N=6, S=3, num_data=1000, H=10, W=10, C=1, conv_specs=[], last_spec=(5, 1, 8))
    # cfg2-like: ConvLayer -> SVGP(ConvKernel)
    case_dgp("dgp2_elbo", 32, N=4, S=3, num_data=100, H=12, W=12, C=1, conv_specs=[(5, 2, 8, 3)],
             last_spec=(3, 1, 8))
    # cfg3-like: ConvLayer(s=2) -> ConvLayer(s=1) -> SVGP(ConvKernel)
    case_dgp("dgp3_elbo", 33, N=3, S=2, num_data=500, H=14, W=14, C=2,
             conv_specs=[(3, 2, 10, 3), (3, 1, 9, 2)], last_spec=(3, 1, 8))
    case_dgp("dgp3_elbo_white", 34, N=3, S=2, num_data=500, H=14, W=14, C=2,
             conv_specs=[(3, 2, 10, 3), (3, 1, 9, 2)], last_spec=(3, 1, 8), white=True)
