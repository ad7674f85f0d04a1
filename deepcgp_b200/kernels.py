"""Kernels -- host-side mirror of conv_gp/kernels.py (ConvKernel, PatchInducingFeatures, Kuu/Kuf) plus the
minimal stand-in for gpflow.kernels.RBF that carries the two hyper-parameters.
"""
import numpy as np
import torch

from . import _lib
from .views import FullView

JITTER = 1e-3  # reference gpflowrc:11


class RBF(object):
    """gpflow.kernels.RBF as used at models.py:113-117,184: constrained variance / lengthscales (scalars)."""

    def __init__(self, input_dim, variance=1.0, lengthscales=1.0):
        self.input_dim = int(input_dim)
        self.variance = float(variance)
        self.lengthscales = float(lengthscales)

    def K(self, X, X2=None):
        """RBF.K(Z) for inducing patches (float64, M x M); X2 is not needed on the hot path."""
        if X2 is not None:
            raise NotImplementedError("RBF.K(X, X2) is reached through MultiOutputConvKernel.Kuf / ConvKernel.Kzx")
        Z = _lib.f64(X)
        M, L = Z.shape
        out = torch.empty((M, M), dtype=torch.float64, device=Z.device)
        _lib.check(_lib.lib.dcgp_kuu(_lib.ptr(Z), M, L, self.variance, self.lengthscales, 0.0, _lib.ptr(out), _lib.stream()))
        return out

    def Kdiag(self, X):
        return torch.full((X.shape[0],), self.variance, dtype=torch.float32, device=X.device)


class PatchInducingFeatures(object):
    """conv_gp/kernels.py:166-170 (InducingPointsBase with Z [M, L], float64 like the reference)."""

    def __init__(self, Z):
        self.Z = Z if isinstance(Z, torch.Tensor) else torch.as_tensor(np.asarray(Z))
        self.Z = self.Z.to(torch.float64).contiguous()

    def __len__(self):
        return self.Z.shape[0]

    @classmethod
    def from_images(cls, NHWC_X, M, patch_size, seed=0):
        """kernels.py:147-164 draws M*100 random patches and k-means them (init only, host side).
        Here: a seeded sample of M patches (same offsets rule: randint(0, H - f), kernels.py:142-143)."""
        rng = np.random.RandomState(seed)
        X = np.asarray(NHWC_X.cpu() if isinstance(NHWC_X, torch.Tensor) else NHWC_X)
        N, H, W, Cc = X.shape
        out = np.zeros((M, patch_size * patch_size * Cc))
        for i in range(M):
            n = rng.randint(0, N)
            y = rng.randint(0, H - patch_size)
            x = rng.randint(0, W - patch_size)
            out[i] = X[n, y:y + patch_size, x:x + patch_size].reshape(-1)
        return cls(out)


class AdditivePatchKernel(object):
    """conv_gp/kernels.py:15-32 constructor state shared with ConvKernel."""

    def __init__(self, base_kernel, view, patch_weights=None):
        self.base_kernel = base_kernel
        self.view = view
        self.patch_length = view.patch_length
        self.patch_count = view.patch_count
        self.image_size = view.input_size
        self.input_dim = int(np.prod(view.input_size))
        if patch_weights is None or int(np.size(patch_weights)) != self.patch_count:   # kernels.py:26-27
            patch_weights = np.ones(self.patch_count)
        self.patch_weights = patch_weights


class ConvKernel(AdditivePatchKernel):
    """conv_gp/kernels.py:79-136: image-level kernel of the final SVGP layer."""

    def _desc(self, M=1, R=1):
        H, W = int(self.view.input_size[0]), int(self.view.input_size[1])
        return _lib.LayerDesc(_lib.LAYER_SVGP_CONV, H, W, self.view.feature_maps, self.view.filter_size,
                              self.view.stride, M, R, 0, self.base_kernel.variance, self.base_kernel.lengthscales, JITTER)

    def _w(self, device):
        return _lib.f64(self.patch_weights, device=device)

    def Kzx(self, Z, ND_X):
        """kernels.py:117-133 -> [M, N]"""
        X = _lib.f32(ND_X)
        Z = _lib.f64(Z, device=X.device)
        N, M = X.shape[0], Z.shape[0]
        d = self._desc(M)
        ws = torch.empty(_lib.lib.dcgp_convkernel_kzx_workspace_bytes(d, N), dtype=torch.uint8, device=X.device)
        out = torch.empty((M, N), dtype=torch.float32, device=X.device)
        w = self._w(X.device)
        _lib.check(_lib.lib.dcgp_convkernel_kzx(d, _lib.ptr(Z), _lib.ptr(w), _lib.ptr(X), N, _lib.ptr(out),
                                                _lib.ptr(ws), ws.numel(), _lib.stream()))
        return out

    def Kdiag(self, ND_X):
        """kernels.py:106-115 -> [N]"""
        X = _lib.f32(ND_X)
        out = torch.empty((X.shape[0],), dtype=torch.float32, device=X.device)
        w = self._w(X.device)
        _lib.check(_lib.lib.dcgp_convkernel_kdiag(self._desc(), _lib.ptr(w), _lib.ptr(X), X.shape[0], _lib.ptr(out),
                                                  _lib.stream()))
        return out

    def Kzz(self, Z):
        """kernels.py:135-136"""
        return self.base_kernel.K(Z)


def Kuu(feature, kern, jitter=0.0):
    """kernels.py:172-174 (the GPflow dispatch target for PatchInducingFeatures x AdditivePatchKernel)."""
    K = kern.Kzz(feature.Z)
    return K + torch.eye(len(feature), dtype=K.dtype, device=K.device) * jitter


def Kuf(feature, kern, Xnew):
    """kernels.py:176-178"""
    return kern.Kzx(feature.Z, Xnew)
