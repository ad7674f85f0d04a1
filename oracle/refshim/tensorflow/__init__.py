"""TEST INFRASTRUCTURE ONLY -- a numpy/float64 eager stand-in for the handful of TensorFlow-1.x
ops the reference's hot path emits.

Purpose: TensorFlow >= 1.11 cannot be installed in this image (no wheels for Python 3.12, no
network), so the reference's *own, unmodified* source files under /root/reference are imported on top
of this shim by ``tests/golden/make_golden.py`` to generate the committed golden vectors.  Every op
below is restated from the TF-1.x API documentation; nothing here is shipped or measured.

Each function evaluates immediately on ``numpy`` arrays (there is no graph / session).
"""
import numpy as _np
import scipy.linalg as _sla
import scipy.special as _ssp

float64 = _np.float64
float32 = _np.float32
int32 = _np.int32
int64 = _np.int64


class Tensor(_np.ndarray):
    """ndarray with the few tf.Tensor / gpflow.Param accessors the reference touches."""

    def get_shape(self):
        class _S(tuple):
            @property
            def ndims(s):
                return len(s)

            def as_list(s):
                return list(s)
        return _S(self.shape)

    def read_value(self):
        return self

    @property
    def value(self):
        return _np.array(self)

    def set_trainable(self, flag):
        pass

    # tf tensors are immutable: `a += b` builds a new (broadcast) tensor, never writes in place
    def __iadd__(self, o):
        return _np.add(_np.asarray(self), _np.asarray(o)).view(Tensor)

    def __isub__(self, o):
        return _np.subtract(_np.asarray(self), _np.asarray(o)).view(Tensor)

    def __imul__(self, o):
        return _np.multiply(_np.asarray(self), _np.asarray(o)).view(Tensor)

    def __itruediv__(self, o):
        return _np.true_divide(_np.asarray(self), _np.asarray(o)).view(Tensor)


def _t(x, dtype=None):
    a = _np.asarray(x, dtype=dtype)
    if a.dtype == _np.float32 and dtype is None:
        a = a.astype(_np.float64)
    return a.view(Tensor)


def constant(x, dtype=None):
    return _t(x, dtype)


convert_to_tensor = constant


def cast(x, dtype=None):
    return _t(_np.asarray(x).astype(dtype))


def shape(x, out_type=None):
    return tuple(int(d) for d in _np.shape(x))


def size(x, out_type=None):
    return int(_np.size(x))


def reshape(x, shp):
    shp = [int(s) for s in (shp if not isinstance(shp, (int, _np.integer)) else [shp])]
    return _t(_np.reshape(_np.asarray(x), shp))


def transpose(x, perm=None):
    return _t(_np.transpose(_np.asarray(x), perm))


def tile(x, multiples):
    return _t(_np.tile(_np.asarray(x), [int(m) for m in multiples]))


def expand_dims(x, axis):
    return _t(_np.expand_dims(_np.asarray(x), axis))


def stack(xs, axis=0):
    return _t(_np.stack([_np.asarray(x) for x in xs], axis=axis))


def concat(xs, axis):
    return _t(_np.concatenate([_np.asarray(x) for x in xs], axis=axis))


def zeros(shp, dtype=float64):
    return _t(_np.zeros([int(s) for s in shp], dtype=dtype))


def ones(shp, dtype=float64):
    return _t(_np.ones([int(s) for s in shp], dtype=dtype))


def zeros_like(x):
    return _t(_np.zeros_like(_np.asarray(x)))


def fill(dims, value):
    return _t(_np.full([int(d) for d in dims], value, dtype=_np.float64))


def eye(n, dtype=float64):
    return _t(_np.eye(int(n), dtype=dtype))


def _axis(axis, reduction_indices):
    a = axis if axis is not None else reduction_indices
    if isinstance(a, list):
        a = tuple(a)
    return a


def reduce_sum(x, axis=None, reduction_indices=None, keepdims=False):
    if isinstance(x, (list, tuple)):
        x = _np.stack([_np.asarray(v) for v in x])
    return _t(_np.sum(_np.asarray(x), axis=_axis(axis, reduction_indices), keepdims=keepdims))


def reduce_mean(x, axis=None, reduction_indices=None):
    return _t(_np.mean(_np.asarray(x), axis=_axis(axis, reduction_indices)))


def reduce_prod(x, axis=None, reduction_indices=None):
    return _t(_np.prod(_np.asarray(x), axis=_axis(axis, reduction_indices)))


def reduce_logsumexp(x, axis=None):
    return _t(_ssp.logsumexp(_np.asarray(x), axis=axis))


def square(x):
    return _t(_np.square(_np.asarray(x)))


def sqrt(x):
    return _t(_np.sqrt(_np.asarray(x)))


def exp(x):
    return _t(_np.exp(_np.asarray(x)))


def log(x):
    return _t(_np.log(_np.asarray(x)))


def erf(x):
    return _t(_ssp.erf(_np.asarray(x)))


def maximum(a, b):
    return _t(_np.maximum(a, b))


def clip_by_value(x, lo, hi):
    return _t(_np.clip(_np.asarray(x), lo, hi))


def one_hot(indices, depth, on_value=1.0, off_value=0.0):
    idx = _np.asarray(indices).astype(_np.int64)
    out = _np.full(idx.shape + (int(depth),), off_value, dtype=_np.float64)
    _np.put_along_axis(out, idx[..., None], on_value, axis=-1)
    return _t(out)


def matmul(a, b, transpose_a=False, transpose_b=False):
    a = _np.asarray(a)
    b = _np.asarray(b)
    if transpose_a:
        a = _np.swapaxes(a, -1, -2)
    if transpose_b:
        b = _np.swapaxes(b, -1, -2)
    return _t(_np.matmul(a, b))


def tensordot(a, b, axes):
    return _t(_np.tensordot(_np.asarray(a), _np.asarray(b), axes=axes))


def cholesky(a):
    """tf.cholesky: lower factor, batched over leading dims; raises on non-PD like
    tf.errors.InvalidArgumentError."""
    try:
        return _t(_np.linalg.cholesky(_np.asarray(a)))
    except _np.linalg.LinAlgError as e:  # pragma: no cover
        raise errors.InvalidArgumentError(str(e))


def matrix_triangular_solve(matrix, rhs, lower=True, adjoint=False):
    m = _np.asarray(matrix)
    r = _np.asarray(rhs)
    if m.ndim == 2:
        return _t(_sla.solve_triangular(m, r, lower=lower, trans=1 if adjoint else 0))
    out = _np.empty(_np.broadcast_shapes(m.shape[:-2], r.shape[:-2]) + r.shape[-2:])
    for idx in _np.ndindex(*out.shape[:-2]):
        out[idx] = _sla.solve_triangular(m[idx], r[idx], lower=lower, trans=1 if adjoint else 0)
    return _t(out)


def cholesky_solve(chol, rhs):
    y = _sla.solve_triangular(_np.asarray(chol), _np.asarray(rhs), lower=True)
    return _t(_sla.solve_triangular(_np.asarray(chol).T, y, lower=False))


def matrix_band_part(x, num_lower, num_upper):
    x = _np.asarray(x)
    assert num_lower == -1 and num_upper == 0, "shim only restates the lower-triangular use"
    return _t(_np.tril(x))


def matrix_diag_part(x):
    return _t(_np.diagonal(_np.asarray(x), axis1=-2, axis2=-1))


def map_fn(fn, elems, dtype=None, parallel_iterations=None):
    """tf.map_fn: apply fn to slices along axis 0 (elems may be a tuple of tensors), stack results."""
    if isinstance(elems, (tuple, list)):
        n = _np.shape(elems[0])[0]
        outs = [fn(tuple(_t(e[i]) for e in elems)) for i in range(n)]
    else:
        outs = [fn(_t(elems[i])) for i in range(_np.shape(elems)[0])]
    if isinstance(outs[0], (tuple, list)):
        return tuple(_t(_np.stack([_np.asarray(o[k]) for o in outs])) for k in range(len(outs[0])))
    return _t(_np.stack([_np.asarray(o) for o in outs]))


def extract_image_patches(images, ksizes, strides, rates, padding):
    """tf.extract_image_patches, VALID padding only.  Output [N, OH, OW, kh*kw*C]; the depth axis is
    ordered (row offset, col offset, channel) with channel fastest, as documented for TF 1.x."""
    assert padding == "VALID"
    x = _np.asarray(images)
    n, h, w, c = x.shape
    kh, kw = int(ksizes[1]), int(ksizes[2])
    sh, sw = int(strides[1]), int(strides[2])
    rh, rw = int(rates[1]), int(rates[2])
    eh, ew = (kh - 1) * rh + 1, (kw - 1) * rw + 1
    oh, ow = (h - eh) // sh + 1, (w - ew) // sw + 1
    out = _np.empty((n, oh, ow, kh * kw * c), dtype=x.dtype)
    for dy in range(kh):
        for dx in range(kw):
            blk = x[:, dy * rh: dy * rh + (oh - 1) * sh + 1: sh, dx * rw: dx * rw + (ow - 1) * sw + 1: sw, :]
            out[..., (dy * kw + dx) * c:(dy * kw + dx + 1) * c] = blk
    return _t(out)


class _RandomState:
    """tf.random_normal stand-in: seeded numpy stream; every draw is recorded so a golden file can
    carry the z noise the reference consumed."""

    def __init__(self):
        self.rng = _np.random.RandomState(0)
        self.draws = []

    def seed(self, s):
        self.rng = _np.random.RandomState(s)
        self.draws = []


random_state = _RandomState()


def random_normal(shp, dtype=float64, seed=None):
    z = random_state.rng.standard_normal([int(s) for s in shp]).astype(_np.float64)
    random_state.draws.append(z)
    return _t(z)


def set_random_seed(s):
    random_state.seed(s)


class _NN:
    @staticmethod
    def conv2d(x, filt, strides, padding, data_format="NHWC"):
        assert padding == "VALID" and data_format == "NHWC"
        kh, kw, ci, co = _np.shape(filt)
        p = _np.asarray(extract_image_patches(x, [1, kh, kw, 1], strides, [1, 1, 1, 1], "VALID"))
        return _t(p @ _np.asarray(filt).reshape(kh * kw * ci, co))


nn = _NN()


class _Linalg:
    cholesky = staticmethod(cholesky)


linalg = _Linalg()


class _Errors:
    class InvalidArgumentError(Exception):
        pass


errors = _Errors()


class Session:
    def run(self, x, feed_dict=None):
        return x


class _Train:
    @staticmethod
    def exponential_decay(lr, step, decay_steps, decay_rate, staircase=False):
        e = step / decay_steps
        if staircase:
            e = _np.floor(e)
        return lr * decay_rate ** e


train = _Train()
