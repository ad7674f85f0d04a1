"""Model assembly, inducing-patch initialisation and parameter checkpoints -- host-side mirror of conv_gp/models.py
(ModelBuilder :35-240, identity_conv :29-33), conv_gp/kernels.py:139-170 (k-means patch init) and
conv_gp/experiment.py:56-64 (.npy parameter dump).  SURVEY.md 8 row f3: runs once, on the host; the layers it builds
are the CUDA-backed ones of this package.

The flags object is the reference's argparse namespace (conv_gp/arguments.py:9-43); built: RBF base kernels,
`--last-kernel conv`, Zero mean functions or `--identity-mean` (Conv2dMean).
"""
import argparse

import numpy as np
import torch

from .dgp import DGP_Base
from .kernels import RBF, ConvKernel, PatchInducingFeatures
from .layers import Conv2dMean, ConvLayer, SVGP_Layer, Zero
from .likelihoods import MultiClass
from .views import FullView


def parse_ints(int_string):
    """models.py:14-18"""
    if int_string == '':
        return []
    return [int(i) for i in int_string.split(',')]


def image_HW(patch_count):
    """models.py:20-22"""
    image_height = int(np.sqrt(patch_count))
    return [image_height, image_height]


def default_parser():
    """The model-shaping flags of conv_gp/arguments.py:9-43 (same names and defaults)."""
    p = argparse.ArgumentParser()
    p.add_argument('--num-samples', type=int, default=10)
    p.add_argument('--lr', type=float, default=0.01)
    p.add_argument('--batch-size', type=int, default=32)
    p.add_argument('-M', type=str, default='384,384')
    p.add_argument('--feature-maps', type=str, default='10')
    p.add_argument('--filter-sizes', type=str, default='5,5')
    p.add_argument('--strides', type=str, default='2,1')
    p.add_argument('--base-kernel', type=str, default='rbf')
    p.add_argument('--white', action='store_true', default=False)
    p.add_argument('--last-kernel', type=str, default='conv')
    p.add_argument('--identity-mean', action='store_true')
    p.add_argument('--optimizer', type=str, default='Adam')
    p.add_argument('--gamma', type=float, default=0.001)
    p.add_argument('--load-model', type=str, default=None)
    return p


def identity_conv(NHWC_X, filter_size, feature_maps_in, feature_maps_out, stride, rng=None, samples=1000):
    """models.py:29-33 + mean_functions.py:6-26: push `samples` random images through the fixed centre-tap filter
    (identity_filter[f//2, f//2, :, :] = 1, VALID, stride) to get the nominal input of the next layer.  Every output map
    is the SUM over the input maps of the centre pixel."""
    rng = np.random if rng is None else rng
    X = np.asarray(NHWC_X)
    idx = rng.choice(np.arange(X.shape[0]), size=samples)
    X = X[idx]
    H, W = X.shape[1], X.shape[2]
    oh, ow = (H - filter_size) // stride + 1, (W - filter_size) // stride + 1
    c = filter_size // 2
    centre = X[:, c:c + (oh - 1) * stride + 1:stride, c:c + (ow - 1) * stride + 1:stride, :]
    assert centre.shape[-1] == feature_maps_in
    return np.repeat(centre.sum(axis=3, keepdims=True), feature_maps_out, axis=3)


def _sample_patches(HW_image, N, patch_size, patch_length, rng):
    """kernels.py:139-145 (offsets drawn from randint(0, H - f), as there)"""
    out = np.zeros((N, patch_length))
    for i in range(N):
        y = rng.randint(0, HW_image.shape[0] - patch_size)
        x = rng.randint(0, HW_image.shape[1] - patch_size)
        out[i] = HW_image[y:y + patch_size, x:x + patch_size].reshape(patch_length)
    return out


def cluster_patches(NHWC_X, M, patch_size, rng=None, samples_per_inducing_point=100):
    """kernels.py:147-164: k-means (M clusters, init='random') over M*100 random patches of random images."""
    from sklearn import cluster
    rng = np.random.RandomState(0) if rng is None else rng
    X = np.asarray(NHWC_X)
    L = patch_size ** 2 * X.shape[3]
    n = M * samples_per_inducing_point
    patches = np.zeros((n, L))
    for i in range(n):
        image = X[rng.randint(0, X.shape[0])]
        patches[i] = _sample_patches(image, 1, patch_size, L, rng)[0]
    km = cluster.KMeans(n_clusters=M, init='random', n_init=1, random_state=rng.randint(0, 2 ** 31 - 1))
    km.fit(patches)
    return km.cluster_centers_


def patch_features_from_images(NHWC_X, M, patch_size, rng=None, samples_per_inducing_point=100):
    """PatchInducingFeatures.from_images (kernels.py:166-170) with the reference's k-means initialisation."""
    return PatchInducingFeatures(cluster_patches(NHWC_X, M, patch_size, rng, samples_per_inducing_point))


# --------------------------------------------------------------------------------------------- checkpoints
def model_parameters(model, global_step=0):
    """experiment.py:56-64: {GPflow pathname: constrained value} + 'global_step'.  Key scheme DGP/layers/<i>/..."""
    params = {}
    for i, layer in enumerate(model.layers):
        base = "DGP/layers/%d/" % i
        last = isinstance(layer, SVGP_Layer)
        kbase = base + ("kern/base_kernel/" if last else "base_kernel/")
        params[kbase + "variance"] = np.float64(layer._base_kernel.variance)
        params[kbase + "lengthscales"] = np.float64(layer._base_kernel.lengthscales)
        params[base + "feature/Z"] = layer.feature.Z.detach().cpu().numpy().copy()
        params[base + "q_mu"] = layer.q_mu.detach().cpu().numpy().copy()
        params[base + "q_sqrt"] = torch.tril(layer.q_sqrt).detach().cpu().numpy().copy()
        if last:
            params[base + "kern/patch_weights"] = np.asarray(
                layer.kern.patch_weights.detach().cpu() if isinstance(layer.kern.patch_weights, torch.Tensor)
                else layer.kern.patch_weights, dtype=np.float64).copy()
    params["global_step"] = int(global_step)
    return params


def save_model_parameters(model, path, global_step=0):
    np.save(path, model_parameters(model, global_step))


def load_layer_parameters(path_or_dict, n_model_layers):
    """models.py:200-233: group a saved dict by layer; when the checkpoint has fewer layers than the model, its last
    layer's parameters move to the model's last layer."""
    parameters = path_or_dict if isinstance(path_or_dict, dict) else np.load(path_or_dict, allow_pickle=True).item()
    parameters = dict(parameters)
    global_step = parameters.pop('global_step')
    layer_params = {}
    for key, value in parameters.items():
        if 'layers' not in key:
            continue
        parts = key.split('/')
        layer, path = int(parts[2]), "/".join(parts[3:])
        vals = layer_params.setdefault(layer, {})
        for name in ('q_mu', 'q_sqrt', 'Z', 'base_kernel/variance', 'base_kernel/lengthscales', 'patch_weights'):
            if name in path:
                vals[name] = value
                break
    stored_layers = max(layer_params.keys()) + 1
    assert stored_layers <= n_model_layers, "Can't load model if the checkpoint has more layers than the model"
    if stored_layers != n_model_layers:
        layer_params[n_model_layers - 1] = layer_params.pop(stored_layers - 1)
    return global_step, layer_params


# --------------------------------------------------------------------------------------------- ModelBuilder
class ModelBuilder(object):
    """models.py:35-198 for the default flags (rbf base kernel, conv last kernel, Zero means)."""

    def __init__(self, flags, NHWC_X_train, Y_train, model_path=None, device="cuda", seed=0, num_data=None):
        self.flags = flags
        self.X_train = np.asarray(NHWC_X_train)
        self.Y_train = Y_train
        self.model_path = model_path
        self.global_step = None
        self.device = device
        self.rng = np.random.RandomState(seed)
        self.num_data = num_data

    def build(self):
        f = self.flags
        if getattr(f, "base_kernel", "rbf") != "rbf" or getattr(f, "last_kernel", "conv") != "conv":
            raise NotImplementedError("only --base-kernel rbf / --last-kernel conv (the defaults) are on the CUDA path")
        Ms = parse_ints(f.M)
        feature_maps = parse_ints(f.feature_maps)
        strides = parse_ints(f.strides)
        filter_sizes = parse_ints(f.filter_sizes)
        loaded = {}
        if getattr(f, "load_model", None) is not None:                                           # models.py:50-53
            self.global_step, loaded = load_layer_parameters(self.model_path, len(Ms))
        assert len(strides) == len(filter_sizes)
        assert len(feature_maps) == (len(Ms) - 1)
        layers, H_X = [], self.X_train
        for i in range(len(feature_maps)):
            layer, H_X = self._conv_layer(H_X, Ms[i], feature_maps[i], filter_sizes[i], strides[i], loaded.get(i))
            layers.append(layer)
        last_params = loaded[max(loaded.keys())] if len(loaded) > 0 else None
        layers.append(self._last_layer(H_X, Ms[-1], filter_sizes[-1], strides[-1], last_params))
        X = self.X_train.reshape(-1, int(np.prod(self.X_train.shape[1:])))
        return DGP_Base(X, self.Y_train, likelihood=MultiClass(10), num_samples=f.num_samples, layers=layers,
                        minibatch_size=f.batch_size, num_data=self.num_data or X.shape[0], device=self.device)

    def _conv_layer(self, NHWC_X, M, feature_map, filter_size, stride, layer_params=None):
        layer_params = layer_params or {}
        NHWC = NHWC_X.shape
        view = FullView(input_size=NHWC[1:3], filter_size=filter_size, feature_maps=NHWC[3], stride=stride)
        H_X = identity_conv(NHWC_X, filter_size, NHWC[3], feature_map, stride, rng=self.rng)
        if 'Z' in layer_params:
            feat = PatchInducingFeatures(layer_params['Z'])
        else:
            feat = patch_features_from_images(NHWC_X, M, filter_size, rng=self.rng)
        L = filter_size ** 2 * NHWC[3]
        kern = RBF(L, variance=float(layer_params.get('base_kernel/variance', 5.0)),
                   lengthscales=float(layer_params.get('base_kernel/lengthscales', 5.0)))       # models.py:113-117
        q_mu, q_sqrt = layer_params.get('q_mu'), layer_params.get('q_sqrt')
        if getattr(self.flags, "identity_mean", False):                                          # models.py:95-100 (not trainable)
            conv_mean = Conv2dMean(filter_size, NHWC[3], feature_map, stride=stride)
        else:
            conv_mean = Zero()
        layer = ConvLayer(base_kernel=kern, mean_function=conv_mean, feature=feat, view=view, white=self.flags.white,
                          gp_count=feature_map, q_mu=q_mu, q_sqrt=q_sqrt, device=self.device)
        if q_sqrt is None:
            layer.q_sqrt = (layer.q_sqrt * 1e-5).contiguous()                                    # models.py:136-138
        return layer, H_X

    def _last_layer(self, H_X, M, filter_size, stride, layer_params=None):
        layer_params = layer_params or {}
        NHWC = H_X.shape
        Z, q_mu, q_sqrt = layer_params.get('Z'), layer_params.get('q_mu'), layer_params.get('q_sqrt')
        if Z is not None:
            saved_filter_size = int(np.sqrt(np.shape(Z)[1] / NHWC[3]))
            if filter_size != saved_filter_size:                                                 # models.py:153-159
                print("filter_size {} != {} for last layer. Resetting parameters.".format(filter_size, saved_filter_size))
                Z = q_mu = q_sqrt = None
        view = FullView(input_size=NHWC[1:], filter_size=filter_size, feature_maps=NHWC[3], stride=stride)
        feat = PatchInducingFeatures(Z) if Z is not None else patch_features_from_images(H_X, M, filter_size, rng=self.rng)
        kern = ConvKernel(base_kernel=RBF(filter_size ** 2 * NHWC[3],
                                          variance=float(layer_params.get('base_kernel/variance', 5.0)),
                                          lengthscales=float(layer_params.get('base_kernel/lengthscales', 5.0))),
                          view=view, patch_weights=layer_params.get('patch_weights'))
        return SVGP_Layer(kern=kern, num_outputs=10, feature=feat, mean_function=Zero(10), white=self.flags.white,
                          q_mu=q_mu, q_sqrt=q_sqrt, device=self.device)
