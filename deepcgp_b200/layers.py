"""The conv-GP layer and the DS-DGP layer API -- host-side mirror of conv_gp/layers.py
(MultiOutputConvKernel :12-50, ConvLayer :52-161) and of the plugin host in
submodules/Doubly-Stochastic-DGP/doubly_stochastic_dgp/layers.py (Layer :37-121, SVGP_Layer :124-256).

Same class names, constructor keywords and method signatures, so the doubly-stochastic training loop
(`propagate` -> `sample_from_conditional` -> `conditional_SND` -> `conditional_ND`, plus `KL()`) runs
unchanged on top.  Values are eager CUDA tensors (float32 activations, float64 parameters); all
arithmetic happens in libdcgp.so.
"""
import numpy as np
import torch

from . import _lib
from .kernels import JITTER, ConvKernel, PatchInducingFeatures


class Zero(object):
    """gpflow.mean_functions.Zero -- the default mean of every layer (models.py:95-100,192)."""

    def __init__(self, output_dim=1):
        self.output_dim = output_dim

    def __call__(self, X):
        return 0.0


class Conv2dMean(object):
    """conv_gp/mean_functions.py:28-41 (Conv2dMean over IdentityConv2dMean :6-26): a FIXED (non-trainable, models.py:100)
    VALID convolution whose only non-zero tap is filter[f//2, f//2, 0, 0] = 1 -- output map 0 copies the centre pixel of
    input map 0 of every patch, the other maps have zero mean; flattened to [N, P*R] (p-major, r fastest: the layer's
    output layout, layers.py:128-134).  A strided slice, no arithmetic."""

    def __init__(self, filter_size, feature_maps_in, feature_maps_out=1, stride=1):
        self.filter_size, self.feature_maps_in = int(filter_size), int(feature_maps_in)
        self.feature_maps_out, self.stride = int(feature_maps_out), int(stride)

    def _taps(self, H, W):
        f, s = self.filter_size, self.stride
        OH, OW = (H - f) // s + 1, (W - f) // s + 1
        c = f // 2
        return slice(c, c + (OH - 1) * s + 1, s), slice(c, c + (OW - 1) * s + 1, s), OH, OW

    def __call__(self, NHWC_X):
        N, H, W, _ = NHWC_X.shape
        ys, xs, OH, OW = self._taps(H, W)
        out = torch.zeros((N, OH * OW, self.feature_maps_out), dtype=NHWC_X.dtype, device=NHWC_X.device)
        out[:, :, 0] = NHWC_X[:, ys, xs, 0].reshape(N, OH * OW)
        return out.reshape(N, OH * OW * self.feature_maps_out)

    def backward(self, g_out, H, W):
        """d/dX of sum(g_out * self(X)): [N, P*R] -> [N, H*W*C] (zeros except the tapped pixels of input map 0)."""
        N = g_out.shape[0]
        ys, xs, OH, OW = self._taps(H, W)
        gX = torch.zeros((N, H, W, self.feature_maps_in), dtype=g_out.dtype, device=g_out.device)
        gX[:, ys, xs, 0] = g_out.reshape(N, OH * OW, self.feature_maps_out)[:, :, 0].reshape(N, OH, OW)
        return gX.reshape(N, H * W * self.feature_maps_in)


class TiledInput(object):
    """`tile(X[None], [S,1,1])` of DS/dgp.py:63 without materialising the S identical copies."""

    def __init__(self, X, S):
        self.X, self.S = X, int(S)

    @property
    def shape(self):
        return (self.S,) + tuple(self.X.shape)


class MultiOutputConvKernel(object):
    """conv_gp/layers.py:12-50."""

    def __init__(self, base_kernel, input_dim, patch_count):
        self.base_kernel = base_kernel
        self.input_dim = int(input_dim)
        self.patch_count = int(patch_count)
        self._ws = _lib.Workspace()

    def Kuu(self, ML_Z):
        """layers.py:18-21: base_kernel.K(Z) + jitter*I, float64 [M,M]."""
        Z = _lib.f64(ML_Z)
        M, L = Z.shape
        out = torch.empty((M, M), dtype=torch.float64, device=Z.device)
        _lib.check(_lib.lib.dcgp_kuu(_lib.ptr(Z), M, L, self.base_kernel.variance, self.base_kernel.lengthscales, JITTER,
                                     _lib.ptr(out), _lib.stream()))
        return out

    def Kuf_images(self, ML_Z, NHWC_X, filter_size, stride):
        """Fused form: Kuf straight from the images (patches never materialised) -> [P,M,N]."""
        X = _lib.f32(NHWC_X)
        Z = _lib.f64(ML_Z, X.device)
        N, H, W, Cc = X.shape
        M, L = Z.shape
        P = ((H - filter_size) // stride + 1) * ((W - filter_size) // stride + 1)
        out = torch.empty((P, M, N), dtype=torch.float32, device=X.device)
        ws = self._ws.get("kuf", _lib.lib.dcgp_kuf_workspace_bytes(M, L), X.device)
        _lib.check(_lib.lib.dcgp_kuf(_lib.ptr(X), N, H, W, Cc, filter_size, stride, _lib.ptr(Z), M,
                                     self.base_kernel.variance, self.base_kernel.lengthscales, 0, 0, _lib.ptr(out),
                                     _lib.ptr(ws), ws.numel(), _lib.stream()))
        return out

    def Kuf(self, ML_Z, PNL_patches):
        """layers.py:23-32: [P,M,N].  A [P,N,L] patch tensor is a stack of P 'images' of size 1x1xL with a 1x1
        filter, so the same fused kernel serves the reference signature."""
        pnl = _lib.f32(PNL_patches)
        P, N, L = pnl.shape
        K = self.Kuf_images(ML_Z, pnl.reshape(P * N, 1, 1, L), 1, 1)      # [1, M, P*N]
        M = K.shape[1]
        return K.reshape(M, P, N).permute(1, 0, 2).contiguous()

    def Kdiag(self, PNL_patches):
        """layers.py:43-50: RBF diagonal = variance, [P,N]."""
        P, N = PNL_patches.shape[:2]
        return torch.full((P, N), self.base_kernel.variance, dtype=torch.float32, device=PNL_patches.device)

    def Kff(self, PNL_patches):
        raise NotImplementedError("Kff is only reached with full_cov=True (layers.py:115-116), outside the hot path")


class Layer(object):
    """DS/layers.py:37-121: multi-sample conditional + reparameterised sampling around `conditional_ND`."""

    input_prop_dim = None

    def conditional_ND(self, X, full_cov=False):
        raise NotImplementedError

    def KL(self):
        return torch.zeros((), dtype=torch.float64)

    def conditional_SND(self, X, full_cov=False):
        """DS/layers.py:53-76: flatten [S,N,D] -> [S*N,D], ONE conditional_ND call, reshape back."""
        if full_cov:
            raise NotImplementedError("full_cov=True is outside the ELBO-step hot path (SURVEY.md 8 f1)")
        if isinstance(X, TiledInput):
            mean, var = self._conditional(X.X, n_rep=X.S)
            S, N = X.S, X.X.shape[0]
        else:
            S, N, D = X.shape
            mean, var = self.conditional_ND(X.reshape(S * N, D))
        return [m.reshape(S, N, self.num_outputs) for m in (mean, var)]

    def sample_from_conditional(self, X, z=None, full_cov=False):
        """DS/layers.py:78-121 (no input propagation on the conv path): returns samples, mean, var, each [S,N,D].
        z=None draws N(0,1) on the device (tf.random_normal at :104)."""
        if full_cov:
            raise NotImplementedError("full_cov=True is outside the ELBO-step hot path (SURVEY.md 8 f1)")
        if isinstance(X, TiledInput):
            S, N, X2, n_rep = X.S, X.X.shape[0], X.X, X.S
        else:
            S, N, D = X.shape
            X2, n_rep = X.reshape(S * N, D), 1
        dev = X2.device
        if z is None:
            z = torch.randn((S, N, self.num_outputs), dtype=torch.float32, device=dev)
        z = _lib.f32(z, dev).reshape(S * N, self.num_outputs)
        mean, var, samples = self._conditional(X2, n_rep=n_rep, z=z)
        shp = (S, N, self.num_outputs)
        return samples.reshape(shp), mean.reshape(shp), var.reshape(shp)


class _PatchGPLayer(Layer):
    """State and library plumbing shared by ConvLayer and SVGP_Layer(ConvKernel)."""

    _kind = None

    def _init_common(self, view, base_kernel, feature, white, R, q_mu, q_sqrt, device):
        self._view, self._base_kernel = view, base_kernel
        self.white = bool(white)
        self.feature = feature
        self.num_inducing = len(feature)
        self._R = int(R)
        self.device = torch.device(device if device is not None else feature.Z.device if feature.Z.is_cuda else "cuda")
        feature.Z = feature.Z.to(self.device)
        M = self.num_inducing
        if q_mu is None:
            q_mu = np.zeros((M, self._R))                                   # layers.py:160-161 / DS/layers.py:164-166
        self.q_mu = _lib.f64(q_mu, self.device)
        self._ws = _lib.Workspace()
        self._prep = None
        self._kl = torch.zeros(1, dtype=torch.float64, device=self.device)
        self._info = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._hold = False
        self._ready = None      # CUDA event recorded after prepare() when it ran on a side stream
        self._ready_fwd = None  # ... and the earlier point from which the forward operands are complete (set by prepare())
        self._ev_fwd = None
        self._pending = None    # set by grad.TrainStep: finishes this layer's pipelined update + prepare() (lazy, see there)
        self._fresh = False     # prepare() for the CURRENT parameters is already queued (its completion event is _ready)
        self._prep_done = None  # event recorded after the most recent prepare(), on whichever stream it ran
        self.algo = None
        if q_sqrt is None:
            if not self.white:                                              # layers.py:154-158 / DS/layers.py:168-174
                Lu = torch.linalg.cholesky(self._Kuu_init().cpu()).to(self.device)   # init-time only, host LAPACK
                q_sqrt = Lu[None].repeat(self._R, 1, 1)
            else:
                q_sqrt = torch.eye(M, dtype=torch.float64)[None].repeat(self._R, 1, 1)
        # gpflow.transforms.LowerTriangular: only the lower triangle is representable
        self.q_sqrt = torch.tril(_lib.f64(q_sqrt, self.device)).contiguous()

    # -- library calls -------------------------------------------------------------------------------
    def _desc(self):
        v, k = self._view, self._base_kernel
        return _lib.LayerDesc(self._kind, int(v.input_size[0]), int(v.input_size[1]), v.feature_maps, v.filter_size,
                              v.stride, self.num_inducing, self._R, int(self.white), float(k.variance),
                              float(k.lengthscales), JITTER)

    def _algo(self):
        from . import default_algo
        return default_algo() if self.algo is None else self.algo

    def _patch_weights(self):
        return None

    def _Z_prior(self):
        return None

    def prepare(self, check=False, hyp=None):
        """Minibatch-independent work of the step (Kuu, Cholesky, L^-1, stacked operand W, KL).
        hyp: device tensor [variance, lengthscale] (float64) that overrides the host's copy of the kernel hyper-parameters --
        grad.TrainStep queues the next step's prepare behind the optimiser update before it has read them back."""
        d = self._desc()
        dev = self.device
        # Two prepares of one layer share the Kuu / factor / operand buffers: whatever streams they are queued on, the
        # later one starts only after the earlier one has drained (TrainStep queues prepares on per-layer side streams).
        cur = torch.cuda.current_stream(dev)
        if self._prep_done is not None:
            cur.wait_event(self._prep_done)
        nprep = _lib.lib.dcgp_prepare_bytes(d)
        if self._prep is None or self._prep.numel() < nprep:
            self._prep = torch.empty(nprep, dtype=torch.uint8, device=dev)
        ws = self._ws.get("prep", _lib.lib.dcgp_prepare_workspace_bytes(d), dev)
        Z = _lib.f64(self.feature.Z, dev)
        Zp = self._Z_prior()
        self._keep = (Z, Zp, _lib.f64(self.q_mu, dev), torch.tril(_lib.f64(self.q_sqrt, dev)).contiguous())
        if self._ev_fwd is None:
            self._ev_fwd = torch.cuda.Event()
            self._ev_fwd.record()               # forces the underlying cudaEvent_t into existence
        if hyp is not None:
            assert hyp.dtype == torch.float64 and hyp.numel() == 2 and hyp.is_cuda
            self._keep = self._keep + (hyp,)
        _lib.check(_lib.lib.dcgp_layer_prepare_hyp(d, _lib.ptr(Z), _lib.ptr(Zp), _lib.ptr(self._keep[2]),
                                                   _lib.ptr(self._keep[3]), self._algo(), _lib.ptr(self._prep),
                                                   _lib.ptr(self._kl), _lib.ptr(ws), ws.numel(), _lib.ptr(self._info),
                                                   self._ev_fwd.cuda_event, _lib.ptr(hyp), _lib.stream()))
        self._ready_fwd = self._ev_fwd          # recorded inside the call, once the forward operands were queued
        self._prep_done = torch.cuda.Event()
        self._prep_done.record(cur)
        if check:
            _lib.raise_if_not_pd(self._info)

    def _run_pending(self):
        if self._pending is not None:
            fn, self._pending = self._pending, None
            fn()

    def _conditional(self, X, n_rep=1, z=None):
        self._run_pending()
        if not self._hold:
            self.prepare()
        elif self._ready is not None:           # prepare() ran on a side stream: wait for its forward operands only
            torch.cuda.current_stream(self.device).wait_event(self._ready_fwd or self._ready)
        d = self._desc()
        X = _lib.f32(X, self.device)
        n_rows = X.shape[0]
        D = self.num_outputs
        rows = n_rows * n_rep
        mean = torch.empty((rows, D), dtype=torch.float32, device=self.device)
        var = torch.empty((rows, D), dtype=torch.float32, device=self.device)
        sample = torch.empty((rows, D), dtype=torch.float32, device=self.device) if z is not None else None
        ws = self._ws.get("apply", _lib.lib.dcgp_apply_workspace_bytes(d, n_rows, n_rep), self.device)
        w = self._patch_weights()
        _lib.check(_lib.lib.dcgp_layer_apply(d, _lib.ptr(self._prep), _lib.ptr(w), _lib.ptr(X), n_rows, n_rep,
                                             _lib.ptr(z), self._algo(), _lib.ptr(mean), _lib.ptr(var), _lib.ptr(sample),
                                             _lib.ptr(ws), ws.numel(), _lib.stream()))
        mf = self._mean_function_value(X, n_rep)
        if mf is not None:                      # layers.py:133-134; the sample is linear in the mean (DS/utils.py:41)
            mean += mf
            if sample is not None:
                sample += mf
        return (mean, var) if z is None else (mean, var, sample)

    def _mean_function_value(self, X, n_rep):
        return None

    def conditional_ND(self, ND_X, full_cov=False):
        if full_cov:
            raise NotImplementedError("full_cov=True is outside the ELBO-step hot path (SURVEY.md 8 f1)")
        mean, var = self._conditional(ND_X)
        if not self._hold:
            _lib.raise_if_not_pd(self._info)
        return mean, var

    def KL(self):
        self._run_pending()
        if not self._hold:
            self.prepare(check=True)
        return self._kl[0]


class ConvLayer(_PatchGPLayer):
    """conv_gp/layers.py:52-161."""

    _kind = _lib.LAYER_CONV

    def __init__(self, base_kernel, mean_function=None, feature=None, view=None, white=False, gp_count=1, q_mu=None,
                 q_sqrt=None, device=None, **kwargs):
        if mean_function is not None and not isinstance(mean_function, (Zero, Conv2dMean)):
            raise NotImplementedError("mean functions of the conv path: Zero (models.py:99) or Conv2dMean (--identity-mean)")
        self.base_kernel = base_kernel
        self.view = view
        self.feature_maps_in = view.feature_maps
        self.gp_count = int(gp_count)
        self.patch_count = view.patch_count
        self.patch_length = view.patch_length
        self.num_outputs = self.patch_count * self.gp_count                 # layers.py:66
        self.conv_kernel = MultiOutputConvKernel(base_kernel, int(np.prod(view.input_size)) * view.feature_maps,
                                                 patch_count=self.patch_count)
        self.mean_function = mean_function or Zero()
        self._init_common(view, base_kernel, feature, white, gp_count, q_mu, q_sqrt, device)
        self._build_prior_cholesky()

    def _Kuu_init(self):
        return self.conv_kernel.Kuu(self.feature.Z)

    def _build_prior_cholesky(self):
        """layers.py:149-152: the KL prior is Kuu at the Z the layer was CONSTRUCTED with."""
        self.Z_prior = self.feature.Z.clone()

    def _Z_prior(self):
        return None if self.white else _lib.f64(self.Z_prior, self.device)

    def _mean_function_value(self, X, n_rep):
        if not isinstance(self.mean_function, Conv2dMean):
            return None
        v = self._view
        mf = self.mean_function(X.reshape(X.shape[0], int(v.input_size[0]), int(v.input_size[1]), v.feature_maps))
        return mf.repeat(n_rep, 1) if n_rep > 1 else mf


class SVGP_Layer(_PatchGPLayer):
    """DS/layers.py:124-256 specialised to kern = ConvKernel with PatchInducingFeatures (models.py:173-198)."""

    _kind = _lib.LAYER_SVGP_CONV

    def __init__(self, kern, num_outputs, mean_function=None, Z=None, feature=None, white=False, input_prop_dim=None,
                 q_mu=None, q_sqrt=None, device=None, **kwargs):
        if not isinstance(kern, ConvKernel):
            raise NotImplementedError("the conv path builds its last layer with ConvKernel (models.py:176-178)")
        if mean_function is not None and not isinstance(mean_function, Zero):
            raise NotImplementedError("only the Zero mean is on the hot path (models.py:192)")
        if feature is None:
            feature = PatchInducingFeatures(Z)
        self.kern = kern
        self.num_outputs = int(num_outputs)
        self.mean_function = mean_function or Zero(num_outputs)
        self._init_common(kern.view, kern.base_kernel, feature, white, num_outputs, q_mu, q_sqrt, device)

    def _Kuu_init(self):
        from .kernels import Kuu
        return Kuu(self.feature, self.kern, jitter=JITTER)                  # DS/layers.py:170

    def _patch_weights(self):
        return _lib.f64(self.kern.patch_weights, self.device)
