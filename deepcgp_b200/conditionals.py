"""Host-side mirror of conv_gp/conditionals.py:6-67 `conditional` (the P-batched sparse-GP conditional)."""
import torch

from . import _lib

_ws = _lib.Workspace()


def conditional(Kmn, Kmm, Knn, f, *, full_cov=False, q_sqrt=None, white=False, algo=None):
    """Same arguments and return layouts as the reference:
      Kmn [P,M,N], Kmm [M,M], Knn [P,N], f [M,R], q_sqrt [R,M,M]  ->  fmean [N,P,R], fvar [R,P,N].
    full_cov=True (conditionals.py:36-38,62-63) is outside the hot path."""
    if full_cov:
        raise NotImplementedError("full_cov=True is not part of the ELBO-step hot path (SURVEY.md 8 f1)")
    if q_sqrt is None:
        raise NotImplementedError("q_sqrt=None is never used by ConvLayer (layers.py:119-120)")
    from . import default_algo
    algo = default_algo() if algo is None else algo
    Kmn = _lib.f32(Kmn)
    dev = Kmn.device
    Kmm, f, q_sqrt = _lib.f64(Kmm, dev), _lib.f64(f, dev), _lib.f64(q_sqrt, dev)
    Knn = _lib.f32(Knn, dev)
    P, M, N = Kmn.shape
    R = f.shape[1]
    if q_sqrt.dim() != 3:
        raise ValueError("Bad dimension for q_sqrt: %s" % q_sqrt.dim())      # conditionals.py:59-61
    nbytes = _lib.lib.dcgp_conditional_workspace_bytes(P, M, N, R)
    ws = _ws.get("cond", nbytes, dev)
    fmean = torch.empty((N, P, R), dtype=torch.float32, device=dev)
    fvar = torch.empty((R, P, N), dtype=torch.float32, device=dev)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(_lib.lib.dcgp_conditional(_lib.ptr(Kmn), _lib.ptr(Kmm), _lib.ptr(Knn), _lib.ptr(f), _lib.ptr(q_sqrt),
                                         int(bool(white)), P, M, N, R, algo, _lib.ptr(fmean), _lib.ptr(fvar),
                                         _lib.ptr(ws), ws.numel(), _lib.ptr(info), _lib.stream()))
    _lib.raise_if_not_pd(info)
    return fmean, fvar
