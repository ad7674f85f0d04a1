"""ctypes binding of libdcgp.so (include/dcgp.h).  PyTorch tensors are used only as device buffers:
every call passes raw device pointers, sizes and the current CUDA stream.

There is NO fallback: if the shared library is missing or a symbol is absent the import fails loudly,
and every entry point raises on a non-zero status.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdcgp.so")

DCGP_OK, DCGP_ERR_ARG, DCGP_ERR_WORKSPACE, DCGP_ERR_NOT_PD, DCGP_ERR_CUDA = 0, 1, 2, 3, 4
LAYER_CONV, LAYER_SVGP_CONV = 0, 1
ALGO_SIMT, ALGO_TC = 0, 1


class NotPositiveDefiniteError(FloatingPointError):
    """Kuu's Cholesky hit a non-positive pivot (the reference raises tf.errors.InvalidArgumentError at
    conditionals.py:29; experiment.py:45-49 catches it for NatGrad)."""


class LayerDesc(C.Structure):
    """dcgp_layer_desc"""
    _fields_ = [("kind", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32), ("f", C.c_int32),
                ("s", C.c_int32), ("M", C.c_int32), ("R", C.c_int32), ("white", C.c_int32),
                ("variance", C.c_double), ("lengthscale", C.c_double), ("jitter", C.c_double)]


_vp, _i, _d, _sz = C.c_void_p, C.c_int, C.c_double, C.c_size_t
_pd = C.POINTER(LayerDesc)

# name -> (restype, argtypes); must list every symbol include/dcgp.h declares (tests/test_abi.py checks).
SIGNATURES = {
    "dcgp_last_error": (C.c_char_p, []),
    "dcgp_version": (_i, []),
    "dcgp_launch_count": (C.c_longlong, []),
    "dcgp_set_kernel_timing": (None, [_i]),
    "dcgp_set_reserved_sms": (None, [_i]),
    "dcgp_kernel_ms": (C.c_double, [_i]),
    "dcgp_kernel_tensor_flops": (C.c_double, [_i]),
    "dcgp_set_products": (None, [_i, _i, _i]),
    "dcgp_get_products": (None, [C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "dcgp_set_precise_stage1": (None, [_i]),
    "dcgp_get_precise_stage1": (_i, []),
    "dcgp_view_geometry": (_i, [_i, _i, _i, _i, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "dcgp_extract_patches": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "dcgp_kuu": (_i, [_vp, _i, _i, _d, _d, _d, _vp, _vp]),
    "dcgp_kuf_workspace_bytes": (_sz, [_i, _i]),
    "dcgp_kuf": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _i, _d, _d, _i, _i, _vp, _vp, _sz, _vp]),
    "dcgp_cholesky_workspace_bytes": (_sz, [_i]),
    "dcgp_cholesky": (_i, [_vp, _i, _vp, _sz, _vp, _vp]),
    "dcgp_conditional_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "dcgp_conditional": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp, _vp]),
    "dcgp_prepare_bytes": (_sz, [_pd]),
    "dcgp_prepare_workspace_bytes": (_sz, [_pd]),
    "dcgp_prepare_workspace_layout": (_i, [_pd, C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_i)]),
    "dcgp_prepare_layout": (_i, [_pd, C.POINTER(_sz), C.POINTER(_i)]),
    "dcgp_prepare_layout2": (_i, [_pd, C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_i)]),
    "dcgp_prepare_workspace_layout2": (_i, [_pd, C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_i)]),
    "dcgp_layer_prepare": (_i, [_pd, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _sz, _vp, _vp]),
    "dcgp_layer_prepare_ev": (_i, [_pd, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _sz, _vp, _vp, _vp]),
    "dcgp_layer_prepare_hyp": (_i, [_pd, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _sz, _vp, _vp, _vp, _vp]),
    "dcgp_apply_workspace_bytes": (_sz, [_pd, _i, _i]),
    "dcgp_layer_apply": (_i, [_pd, _vp, _vp, _vp, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "dcgp_backward_workspace_bytes": (_sz, [_pd, _i, _i]),
    "dcgp_layer_backward": (_i, [_pd, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "dcgp_layer_backward_phases": (_i, [_pd, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _vp]),
    "dcgp_chain_rule_workspace_bytes": (_sz, [_pd]),
    "dcgp_layer_chain_rule": (_i, [_pd, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _d, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "dcgp_bgemm_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "dcgp_bgemm_nt": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, C.c_longlong, C.c_longlong, _vp, _sz, _vp]),
    "dcgp_multiclass_varexp_grad": (_i, [_vp, _vp, _vp, _i, _i, _i, _d, _d, _vp, _vp, _vp]),
    "dcgp_sample_backward": (_i, [_vp, _vp, _vp, _sz, _d, _vp, _vp, _vp]),
    "dcgp_adam": (_i, [_vp, _vp, _vp, _vp, _sz, _d, _d, _d, _d, _i, _i, _vp]),
    "dcgp_convkernel_kzx_workspace_bytes": (_sz, [_pd, _i]),
    "dcgp_convkernel_kzx": (_i, [_pd, _vp, _vp, _vp, _i, _vp, _vp, _sz, _vp]),
    "dcgp_convkernel_kdiag": (_i, [_pd, _vp, _vp, _i, _vp, _vp]),
    "dcgp_randn": (_i, [_vp, _i, _i, _i, C.c_longlong, C.c_longlong, C.c_ulonglong, C.c_ulonglong, _i, _vp]),
    "dcgp_reparameterize": (_i, [_vp, _vp, _vp, _sz, _d, _vp, _vp]),
    "dcgp_multiclass_varexp": (_i, [_vp, _vp, _vp, _i, _i, _i, _d, _vp, _vp, _vp]),
    "dcgp_multiclass_predict": (_i, [_vp, _vp, _vp, _i, _i, _i, _d, _vp, _vp, _vp, _vp]),
    "dcgp_elbo": (_i, [_vp, _i, _d, _d, _vp, _i, _vp, _vp]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "deepcgp_b200: %s not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(status):
    if status == DCGP_OK:
        return
    msg = (lib.dcgp_last_error() or b"").decode()
    if status == DCGP_ERR_ARG:
        raise ValueError("libdcgp: " + msg)
    if status == DCGP_ERR_WORKSPACE:
        raise MemoryError("libdcgp: " + msg)
    if status == DCGP_ERR_NOT_PD:
        raise NotPositiveDefiniteError("libdcgp: " + msg)
    raise RuntimeError("libdcgp (status %d): %s" % (status, msg))


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise ValueError("deepcgp_b200 operates on CUDA tensors only (got a CPU tensor); there is no CPU path")
    if not t.is_contiguous():
        raise ValueError("tensor must be contiguous")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def f32(t, device=None):
    return torch.as_tensor(t, device=device).to(torch.float32).contiguous()


def f64(t, device=None):
    return torch.as_tensor(t, device=device).to(torch.float64).contiguous()


class Workspace:
    """Grow-only device scratch buffers keyed by name (the library never allocates)."""

    def __init__(self):
        self._bufs = {}

    def get(self, key, nbytes, device):
        nbytes = max(int(nbytes), 256)
        b = self._bufs.get(key)
        if b is None or b.numel() < nbytes or b.device != torch.device(device):
            b = torch.empty(nbytes, dtype=torch.uint8, device=device)
            self._bufs[key] = b
        return b


def products():
    """(cond, dk, dq): split products per k-step of the three big T-sized GEMM families (dcgp_get_products)."""
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    lib.dcgp_get_products(C.byref(a), C.byref(b), C.byref(c))
    return a.value, b.value, c.value


def precision_string():
    """The arithmetic the tensor-core path computes in, for bench.py's `dtype` field."""
    cond, dk, dq = products()
    return ("split-fp16 (hi+lo) operands on tcgen05, fp32 accumulate; products per k-step: a=Lm^-1k 3, G_r=C_r^T a %d, dK %d, dQ %d; "
            "f64 M-only" % (cond, dk, dq))


def raise_if_not_pd(info):
    """`info` is the device int written by the Cholesky; reading it synchronises the stream."""
    k = int(info.item())
    if k != 0:
        raise NotPositiveDefiniteError("Kuu is not positive definite: leading minor %d (cf. conditionals.py:29)" % k)
