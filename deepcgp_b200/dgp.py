"""DGP_Base -- host-side mirror of submodules/Doubly-Stochastic-DGP/doubly_stochastic_dgp/dgp.py:35-126:
propagate (:61-76), E_log_p_Y (:83-90), _build_likelihood (:92-98).  TensorFlow's graph/session is
replaced by eager calls into libdcgp.so; the loop structure is the reference's.
"""
import numpy as np
import torch

from . import _lib
from .dist import allreduce_sum_
from .layers import TiledInput
from .likelihoods import BroadcastingLikelihood


class Minibatch(object):
    """gpflow.params.Minibatch(X, batch, seed=0) stand-in: seeded shuffle, repeat, batch (host side)."""

    def __init__(self, X, batch_size, seed=0):
        self.X, self.batch_size = X, int(batch_size)
        self._rng = np.random.RandomState(seed)
        self._perm, self._pos = self._rng.permutation(len(X)), 0

    def next_indices(self):
        if self._pos + self.batch_size > len(self._perm):
            self._perm, self._pos = self._rng.permutation(len(self.X)), 0
        idx = self._perm[self._pos:self._pos + self.batch_size]
        self._pos += self.batch_size
        return idx


class DGP_Base(object):
    def __init__(self, X, Y, likelihood, layers, minibatch_size=None, num_samples=1, num_data=None, device="cuda",
                 seed=0, **kwargs):
        self.num_samples = int(num_samples)
        self.num_data = num_data or X.shape[0]                              # DS/dgp.py:49
        self.device = torch.device(device)
        self.X_all, self.Y_all = X, Y
        self.minibatch_size = minibatch_size
        self._mb = Minibatch(X, minibatch_size, seed=0) if minibatch_size else None   # DS/dgp.py:50-52
        self.likelihood = BroadcastingLikelihood(likelihood)
        self.layers = list(layers)
        self._kls = torch.zeros(len(self.layers), dtype=torch.float64, device=self.device)
        self._sum = torch.zeros(1, dtype=torch.float64, device=self.device)
        self._elbo = torch.zeros(1, dtype=torch.float64, device=self.device)
        self._side = None
        self.seed = int(seed)
        self._z_step = 0

    def draw_zs(self, n_local, n_global=None, n0=0, step=None):
        """The N(0,1) draws of one step (tf.random_normal at DS/layers.py:104), one [S, n_local, D_l] tensor per layer, from
        the counter-based generator dcgp_randn: the draw for (layer, sample, GLOBAL image n0 + n, output) depends only on
        (seed, step) -- not on the rank that holds the image, so image-sharded runs on any number of GPUs see the same
        noise.  `step` defaults to an internal counter that advances by one per call."""
        if step is None:
            self._z_step += 1
            step = self._z_step
        S = self.num_samples
        n_global = int(n_global or n_local)
        zs = []
        for li, layer in enumerate(self.layers):
            z = torch.empty((S, int(n_local), layer.num_outputs), dtype=torch.float32, device=self.device)
            _lib.check(_lib.lib.dcgp_randn(_lib.ptr(z), S, int(n_local), layer.num_outputs, n_global, int(n0), self.seed,
                                           int(step), li, _lib.stream()))
            zs.append(z)
        return zs

    # ------------------------------------------------------------------ DS/dgp.py:61-76
    def propagate(self, X, full_cov=False, S=1, zs=None):
        X = _lib.f32(X, self.device)
        F = TiledInput(X, S)                                                # sX = tile(X[None], [S,1,1]) (:63)
        Fs, Fmeans, Fvars = [], [], []
        zs = zs or [None, ] * len(self.layers)
        for layer, z in zip(self.layers, zs):
            F, Fmean, Fvar = layer.sample_from_conditional(F, z=z, full_cov=full_cov)
            Fs.append(F), Fmeans.append(Fmean), Fvars.append(Fvar)
        return Fs, Fmeans, Fvars

    def _build_predict(self, X, full_cov=False, S=1, zs=None):
        Fs, Fmeans, Fvars = self.propagate(X, full_cov=full_cov, S=S, zs=zs)
        return Fmeans[-1], Fvars[-1]

    def E_log_p_Y(self, X, Y, zs=None):
        """DS/dgp.py:83-90 -> [N, 1]"""
        Fmean, Fvar = self._build_predict(X, full_cov=False, S=self.num_samples, zs=zs)
        var_exp = self.likelihood.variational_expectations(Fmean, Fvar, Y)
        return var_exp.mean(dim=0)

    def _next_batch(self):
        if self._mb is None:
            return self.X_all, self.Y_all
        idx = self._mb.next_indices()
        return self.X_all[idx], self.Y_all[idx]

    def _build_likelihood(self, X=None, Y=None, zs=None, n_global=None, keep=False, defer_sum=False):
        """DS/dgp.py:92-98: ELBO = sum_n E_q[log p(y_n|f_n)] * num_data/batch - sum_l KL_l, as a device scalar.
        The whole step stays on the stream; nothing is read back here.
        defer_sum (grad.TrainStep, image-sharded): leave this rank's data term in self._sum and skip the ELBO kernel -- the
        caller folds the sum into its gradient all-reduce and calls _finish_elbo() afterwards (no collective mid-step)."""
        if X is None:
            X, Y = self._next_batch()
        X = _lib.f32(X, self.device)
        N = X.shape[0]
        S = self.num_samples
        # Minibatch-independent ("M-only") work once per step.  It is latency-bound (Cholesky chains on a handful of
        # CTAs) and independent across layers, so each layer's chain runs on its own side stream, concurrently with
        # the other layers' chains and with the minibatch-sized kernels of earlier layers on the main stream.
        main = torch.cuda.current_stream(self.device)
        if self._side is None:
            self._side = [torch.cuda.Stream(device=self.device) for _ in self.layers]
        for layer, side in zip(self.layers, self._side):
            layer._hold = True
            if layer._pending is not None:      # grad.TrainStep pipelines this layer's update + prepare(); it is
                continue                        # completed lazily, right before the layer is applied
            if layer._fresh and layer._ready is not None:
                continue                        # TrainStep.finish() already queued the prepare for these parameters
            side.wait_stream(main)
            with torch.cuda.stream(side):
                layer.prepare()
                layer._ready = torch.cuda.Event()
                layer._ready.record(side)
        try:
            Fs, Fmeans, Fvars = self.propagate(X, full_cov=False, S=S, zs=zs)
            Fmean, Fvar = Fmeans[-1], Fvars[-1]
            self._fwd = (Fs, Fmeans, Fvars) if keep else None     # the backward pass re-uses the layer outputs
            K = Fmean.shape[2]
            lik = self.likelihood.likelihood
            lik.variational_expectations(Fmean.reshape(S * N, K), Fvar.reshape(S * N, K), Y, S=S, out_sum=self._sum)
            for i, layer in enumerate(self.layers):
                main.wait_event(layer._ready)
                self._kls[i:i + 1].copy_(layer._kl)
        finally:
            for layer in self.layers:
                layer._hold = False
                layer._ready = None
                layer._fresh = False
        self._elbo_args = (S, float(self.num_data), float(n_global or N))
        if defer_sum:
            return self._elbo[0]
        if n_global is not None and n_global != N:
            allreduce_sum_(self._sum)        # image-sharded step: the data term is a sum over all ranks' images
        return self._finish_elbo()

    def _finish_elbo(self):
        """ELBO scalar from self._sum (already summed over ranks) and the KLs, on the current stream."""
        S, num_data, n = self._elbo_args
        _lib.check(_lib.lib.dcgp_elbo(_lib.ptr(self._sum), S, num_data, n, _lib.ptr(self._kls), len(self.layers),
                                      _lib.ptr(self._elbo), _lib.stream()))
        return self._elbo[0]

    # ------------------------------------------------------------------ DS/dgp.py:100-126 (autoflow methods)
    def predict_f(self, Xnew, num_samples, zs=None):
        """(Fmean, Fvar) of the last layer, each [S, N, K]."""
        return self._build_predict(Xnew, full_cov=False, S=int(num_samples), zs=zs)

    def predict_all_layers(self, Xnew, num_samples, zs=None):
        return self.propagate(Xnew, full_cov=False, S=int(num_samples), zs=zs)

    def predict_y(self, Xnew, num_samples, zs=None):
        """DS/dgp.py:116-119: class probabilities and their variance per sample, each [S, N, K] (float64)."""
        Fmean, Fvar = self._build_predict(Xnew, full_cov=False, S=int(num_samples), zs=zs)
        return self.likelihood.predict_mean_and_var(Fmean, Fvar)

    def predict_density(self, Xnew, Ynew, num_samples, zs=None):
        """DS/dgp.py:121-126: log (1/S) sum_s p(y | f_s) -> [N, 1]."""
        S = int(num_samples)
        Fmean, Fvar = self._build_predict(Xnew, full_cov=False, S=S, zs=zs)
        l = self.likelihood.predict_density(Fmean, Fvar, Ynew)
        return torch.logsumexp(l - float(np.log(S)), dim=0)

    def compute_log_likelihood(self, X=None, Y=None, zs=None):
        """GPflow Model.compute_log_likelihood: the ELBO as a Python float (synchronises; raises on a failed Cholesky)."""
        elbo = self._build_likelihood(X, Y, zs)
        for layer in self.layers:
            _lib.raise_if_not_pd(layer._info)
        return float(elbo.item())
