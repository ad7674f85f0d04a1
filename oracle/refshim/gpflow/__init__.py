"""TEST INFRASTRUCTURE ONLY -- a numpy/float64 eager stand-in for the GPflow-1.2.0 symbols the
reference's hot path imports (SURVEY.md section 8c lists them).  GPflow 1.2.0 is a third-party
dependency that is absent from /root/reference (requirements.txt:2) and cannot be installed here, so
its *published* formulas are restated below (RBF.K, gauss_kl, MultiClass/RobustMax variational
expectations, LowerTriangular, Zero mean) -- App. A.5 of SURVEY.md.  Those restatements are
"parity unpinned" (nothing under /root/reference pins them); everything the reference itself owns is
executed from its own unmodified source on top of this shim by tests/golden/make_golden.py.

There is no graph: a Param is an ndarray (constrained value), `params_as_tensors` is the identity.
"""
import configparser as _configparser
import logging as _logging
import os as _os
import sys as _sys
import types as _types

import numpy as _np
import scipy.linalg as _sla
import scipy.special as _ssp
import tensorflow as _tf  # the sibling shim

_t = _tf._t


def _module(name):
    m = _types.ModuleType(name)
    _sys.modules[name] = m
    return m


# ----------------------------------------------------------------------------- settings
class _Settings:
    """float_type / jitter as GPflow would read them from the reference's gpflowrc
    (/root/reference/gpflowrc:6-11 -> float64, jitter 1e-3)."""

    def __init__(self):
        self.float_type = _np.float64
        self.int_type = _np.int32
        self.jitter = 1e-6  # GPflow's packaged default; overridden by gpflowrc below
        rc = _os.environ.get("DCGP_REF_GPFLOWRC", "/root/reference/gpflowrc")
        if _os.path.exists(rc):
            cp = _configparser.ConfigParser()
            cp.read(rc)
            self.jitter = float(cp["numerics"]["jitter_level"])
            assert cp["dtypes"]["float_type"] == "float64"
        self.dtypes = _types.SimpleNamespace(float_type=self.float_type, int_type=self.int_type)
        self.numerics = _types.SimpleNamespace(jitter_level=self.jitter)

    def logger(self):
        return _logging.getLogger("gpflow-shim")


settings = _Settings()
_sys.modules["gpflow.settings"] = settings


# ----------------------------------------------------------------------------- params
class Parameterized:
    def __init__(self, name=None, **kwargs):
        self.name = name

    def __setattr__(self, k, v):
        if isinstance(v, _np.ndarray) and not isinstance(v, _tf.Tensor):
            v = _t(v)
        object.__setattr__(self, k, v)

    def enquire_session(self, session=None):
        return _tf.Session()

    def set_trainable(self, flag):
        pass

    def compile(self):
        pass


class _Transform:
    def forward(self, x):
        return x


def Parameter(value, transform=None, prior=None, trainable=True, dtype=None, fix_shape=True, name=None):
    """The constrained value, as an ndarray.  A LowerTriangular transform keeps only the lower
    triangle (that is what its packed unconstrained vector can represent)."""
    v = _np.array(value, dtype=_np.float64)
    if isinstance(transform, transforms.LowerTriangular):
        v = _np.tril(v)
    return _t(v)


Param = Parameter


class ParamList(list, Parameterized):
    def __init__(self, items, **kw):
        list.__init__(self, items)


def DataHolder(x, **kw):
    return _t(x)


def Minibatch(x, batch_size=None, seed=0, **kw):
    """tf.data minibatching is outside the hot path; the golden script feeds one explicit batch, so
    the holder simply carries whatever array it is given."""
    return _t(x)


def params_as_tensors(f):
    return f


class params_as_tensors_for:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def autoflow(*specs):
    def deco(f):
        return f
    return deco


params = _module("gpflow.params")
params.Parameterized = Parameterized
params.Parameter = Parameter
params.DataHolder = DataHolder
params.Minibatch = Minibatch
params.ParamList = ParamList
decors = _module("gpflow.decors")
decors.params_as_tensors = params_as_tensors
decors.autoflow = autoflow


# ----------------------------------------------------------------------------- transforms
transforms = _module("gpflow.transforms")


class _LowerTriangular(_Transform):
    def __init__(self, N, num_matrices=1, squeeze=False):
        self.N, self.num_matrices = N, num_matrices


class _Logistic(_Transform):
    def __init__(self, a=0.0, b=1.0):
        self.a, self.b = a, b


transforms.LowerTriangular = _LowerTriangular
transforms.Logistic = _Logistic
transforms.positive = _Transform()
transforms.Identity = _Transform


# ----------------------------------------------------------------------------- kernels
kernels = _module("gpflow.kernels")


class Kernel(Parameterized):
    def __init__(self, input_dim, active_dims=None, name=None):
        super().__init__(name=name)
        self.input_dim = int(input_dim)


class RBF(Kernel):
    """GPflow 1.2.0 `Stationary.square_dist` + `RBF.K`: inputs divided by the lengthscale, squared
    distance by the expansion |x|^2 + |y|^2 - 2 x.y (no clamping), K = variance * exp(-d / 2)."""

    def __init__(self, input_dim, variance=1.0, lengthscales=1.0, active_dims=None, ARD=None, name=None):
        super().__init__(input_dim, active_dims, name=name)
        self.variance = Parameter(variance)
        self.lengthscales = Parameter(lengthscales)

    def square_dist(self, X, X2):
        X = _np.asarray(X) / self.lengthscales
        Xs = _np.sum(_np.square(X), axis=1)
        if X2 is None:
            dist = -2 * X @ X.T
            dist += Xs.reshape(-1, 1) + Xs.reshape(1, -1)
            return dist
        X2 = _np.asarray(X2) / self.lengthscales
        X2s = _np.sum(_np.square(X2), axis=1)
        dist = -2 * X @ X2.T
        dist += Xs.reshape(-1, 1) + X2s.reshape(1, -1)
        return dist

    def K(self, X, X2=None, presliced=False):
        return _t(_np.asarray(self.variance) * _np.exp(-self.square_dist(X, X2) / 2))

    def Kdiag(self, X, presliced=False):
        return _t(_np.full((_np.shape(X)[0],), float(self.variance)))

    def compute_K_symm(self, X):
        return _np.asarray(self.K(X))


class ArcCosine(Kernel):
    def __init__(self, input_dim, order=0, **kw):
        super().__init__(input_dim)


kernels.Kernel = Kernel
kernels.RBF = RBF
kernels.ArcCosine = ArcCosine


# ----------------------------------------------------------------------------- features / dispatch
features = _module("gpflow.features")


class InducingPointsBase(Parameterized):
    def __init__(self, Z):
        super().__init__()
        self.Z = Parameter(Z)

    def __len__(self):
        return self.Z.shape[0]


class InducingPoints(InducingPointsBase):
    pass


features.InducingPointsBase = InducingPointsBase
features.InducingPoints = InducingPoints
_mo = _module("gpflow.multioutput")
_mof = _module("gpflow.multioutput.features")
_mof.SeparateIndependentMof = type("SeparateIndependentMof", (Parameterized,), {})
_mo.features = _mof

dispatch_mod = _module("gpflow.dispatch")
_REGISTRY = {}


class _Dispatcher:
    """multipledispatch-style lookup by the types of the leading positional arguments."""

    def __init__(self, name):
        self.name = name
        self.impls = []

    def __call__(self, *args, **kwargs):
        best = None
        for types, fn in self.impls:
            if len(types) <= len(args) and all(isinstance(a, t) for a, t in zip(args, types)):
                if best is None or len(types) >= len(best[0]):
                    best = (types, fn)
        if best is None:
            raise NotImplementedError("%s%s" % (self.name, tuple(type(a).__name__ for a in args)))
        return best[1](*args, **kwargs)


def dispatch(*types):
    def deco(fn):
        d = _REGISTRY.setdefault(fn.__name__, _Dispatcher(fn.__name__))
        d.impls.append((types, fn))
        return d
    return deco


dispatch_mod.dispatch = dispatch


@dispatch(InducingPoints, Kernel)
def Kuu(feat, kern, jitter=0.0):
    return _t(_np.asarray(kern.K(feat.Z)) + jitter * _np.eye(len(feat)))


@dispatch(InducingPoints, Kernel, object)
def Kuf(feat, kern, Xnew):
    return kern.K(feat.Z, Xnew)


conditionals = _module("gpflow.conditionals")
conditionals.Kuu = _REGISTRY["Kuu"]
conditionals.Kuf = _REGISTRY["Kuf"]
conditionals.conditional = None  # imported by DS/layers.py:20, never called on the conv path
features.Kuu = conditionals.Kuu
features.Kuf = conditionals.Kuf


# ----------------------------------------------------------------------------- KL
kullback_leiblers = _module("gpflow.kullback_leiblers")


def gauss_kl(q_mu, q_sqrt, K=None):
    """GPflow 1.2.0 gauss_kl for q_mu [M,L], q_sqrt [L,M,M] (lower-tri), K [M,M] or None (white):
    0.5 * ( mahalanobis - M*L - sum_l log|S_l| + trace [+ L * log|K|] )."""
    q_mu = _np.asarray(q_mu)
    q_sqrt = _np.asarray(q_sqrt)
    white = K is None
    M, B = q_mu.shape
    if white:
        alpha = q_mu
    else:
        Lp = _np.linalg.cholesky(_np.asarray(K))
        alpha = _sla.solve_triangular(Lp, q_mu, lower=True)
    assert q_sqrt.ndim == 3
    Lq = _np.tril(q_sqrt)
    Lq_diag = _np.diagonal(Lq, axis1=-2, axis2=-1)
    mahalanobis = _np.sum(_np.square(alpha))
    constant = -float(q_mu.size)
    logdet_qcov = _np.sum(_np.log(_np.square(Lq_diag)))
    if white:
        trace = _np.sum(_np.square(Lq))
    else:
        LpiLq = _np.stack([_sla.solve_triangular(Lp, Lq[b], lower=True) for b in range(B)])
        trace = _np.sum(_np.square(LpiLq))
    twoKL = mahalanobis + constant - logdet_qcov + trace
    if not white:
        twoKL += B * _np.sum(_np.log(_np.square(_np.diagonal(Lp))))
    return _t(0.5 * twoKL)


kullback_leiblers.gauss_kl = gauss_kl


# ----------------------------------------------------------------------------- likelihoods
likelihoods = _module("gpflow.likelihoods")


class Likelihood(Parameterized):
    def __init__(self, name=None):
        super().__init__(name)
        self.num_gauss_hermite_points = 20


class Gaussian(Likelihood):
    def __init__(self, variance=1.0, **kw):
        super().__init__()
        self.variance = Parameter(variance)


class RobustMax(Parameterized):
    """GPflow 1.2.0 RobustMax(num_classes, epsilon=1e-3)."""

    def __init__(self, num_classes, epsilon=1e-3):
        super().__init__()
        self.epsilon = Parameter(epsilon)
        self.num_classes = num_classes
        self._eps_K1 = float(epsilon) / (num_classes - 1.0)

    def prob_is_largest(self, Y, mu, var, gh_x, gh_w):
        Y = _np.asarray(Y).astype(_np.int64).reshape(-1)
        mu = _np.asarray(mu)
        var = _np.asarray(var)
        oh_on = _np.asarray(_tf.one_hot(Y, self.num_classes, 1.0, 0.0))
        mu_selected = _np.sum(oh_on * mu, 1)
        var_selected = _np.sum(oh_on * var, 1)
        X = mu_selected.reshape(-1, 1) + gh_x * _np.sqrt(_np.clip(2.0 * var_selected, 1e-10, _np.inf)).reshape(-1, 1)
        dist = (X[:, None, :] - mu[:, :, None]) / _np.sqrt(_np.clip(var, 1e-10, _np.inf))[:, :, None]
        cdfs = 0.5 * (1.0 + _ssp.erf(dist / _np.sqrt(2.0)))
        cdfs = cdfs * (1 - 2e-4) + 1e-4
        oh_off = _np.asarray(_tf.one_hot(Y, self.num_classes, 0.0, 1.0))
        cdfs = cdfs * oh_off[:, :, None] + oh_on[:, :, None]
        return _np.prod(cdfs, axis=1) @ (gh_w / _np.sqrt(_np.pi)).reshape(-1, 1)


class MultiClass(Likelihood):
    def __init__(self, num_classes, invlink=None, **kw):
        super().__init__()
        self.num_classes = num_classes
        self.invlink = invlink if invlink is not None else RobustMax(num_classes)

    def variational_expectations(self, Fmu, Fvar, Y):
        gh_x, gh_w = _np.polynomial.hermite.hermgauss(self.num_gauss_hermite_points)
        p = self.invlink.prob_is_largest(Y, Fmu, Fvar, gh_x, gh_w)
        eps = float(self.invlink.epsilon)
        return _t(p * _np.log(1.0 - eps) + (1.0 - p) * _np.log(self.invlink._eps_K1))

    def predict_mean_and_var(self, Fmu, Fvar):
        gh_x, gh_w = _np.polynomial.hermite.hermgauss(self.num_gauss_hermite_points)
        n = _np.shape(Fmu)[0]
        ps = []
        for k in range(self.num_classes):
            p = self.invlink.prob_is_largest(_np.full((n,), k), Fmu, Fvar, gh_x, gh_w)
            eps = float(self.invlink.epsilon)
            ps.append(p * (1.0 - eps) + (1.0 - p) * self.invlink._eps_K1)
        ps = _np.concatenate(ps, axis=1)
        return _t(ps), _t(ps - _np.square(ps))

    def predict_density(self, Fmu, Fvar, Y):
        gh_x, gh_w = _np.polynomial.hermite.hermgauss(self.num_gauss_hermite_points)
        p = self.invlink.prob_is_largest(Y, Fmu, Fvar, gh_x, gh_w)
        eps = float(self.invlink.epsilon)
        return _t(_np.log(p * (1.0 - eps) + (1.0 - p) * self.invlink._eps_K1))


likelihoods.Likelihood = Likelihood
likelihoods.Gaussian = Gaussian
likelihoods.MultiClass = MultiClass
likelihoods.RobustMax = RobustMax
likelihoods.Bernoulli = type("Bernoulli", (Likelihood,), {})


# ----------------------------------------------------------------------------- mean functions
mean_functions = _module("gpflow.mean_functions")


class MeanFunction(Parameterized):
    def __call__(self, X):
        raise NotImplementedError


class Zero(MeanFunction):
    def __init__(self, output_dim=1):
        super().__init__()
        self.output_dim = output_dim

    def __call__(self, X):
        return _t(_np.zeros((_np.shape(X)[0], self.output_dim)))


class Identity(MeanFunction):
    def __call__(self, X):
        return X


class Linear(MeanFunction):
    def __init__(self, A=None, b=None):
        super().__init__()
        self.A, self.b = Parameter(A), Parameter(b)

    def __call__(self, X):
        return _t(_np.asarray(X) @ self.A + self.b)


mean_functions.MeanFunction = MeanFunction
mean_functions.Zero = Zero
mean_functions.Identity = Identity
mean_functions.Linear = Linear


# ----------------------------------------------------------------------------- models & inert imports
models = _module("gpflow.models")
_model = _module("gpflow.models.model")


class Model(Parameterized):
    def compute_log_likelihood(self):
        return self._build_likelihood()


_model.Model = Model
models.model = _model
models.Model = Model
_gplvm = _module("gpflow.models.gplvm")
_gplvm.BayesianGPLVM = type("BayesianGPLVM", (Model,), {})
models.gplvm = _gplvm

for _name, _attrs in [
    ("gpflow.expectations", {"expectation": None}),
    ("gpflow.probability_distributions", {"DiagonalGaussian": None}),
    ("gpflow.logdensities", {"multivariate_normal": None}),
    ("gpflow.priors", {"Gaussian": None, "Beta": None}),
    ("gpflow.quadrature", {"mvhermgauss": None}),
    ("gpflow.actions", {"Loop": None}),
    ("gpflow.train", {}),
    ("gpflow.training", {}),
]:
    _m = _module(_name)
    for _k, _v in _attrs.items():
        setattr(_m, _k, _v)
    globals()[_name.split(".")[1]] = _m
