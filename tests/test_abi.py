"""CPU: the C-ABI library loads, exports every symbol include/dcgp.h declares, and validates arguments
without touching a GPU; host-side mirror logic (geometry, minibatching)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "dcgp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dcgp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from deepcgp_b200 import _lib
    lib = C.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libdcgp.so does not export %s" % n
        assert n in _lib.SIGNATURES, "ctypes binding lacks a signature for %s" % n
    assert sorted(_lib.SIGNATURES) == names
    assert _lib.lib.dcgp_version() >= 100


@pytest.mark.parametrize("H,W,C,f,s", [(28, 28, 1, 5, 1), (28, 28, 1, 5, 2), (32, 32, 3, 5, 2), (14, 14, 10, 5, 1),
                                       (10, 10, 10, 5, 1), (9, 11, 1, 4, 2), (32, 32, 3, 4, 2), (5, 5, 2, 5, 3)])
def test_view_geometry_matches_reference_formula(H, W, C, f, s):
    """views.py:56-68"""
    from deepcgp_b200 import FullView
    v = FullView((H, W), f, C, s)
    oh, ow = (H - f) // s + 1, (W - f) // s + 1
    assert (v.out_image_height, v.out_image_width) == (oh, ow)
    assert v.patch_count == oh * ow and v.patch_length == f * f * C
    assert v.patch_shape == [f, f] and v.dilation == 1 and v.feature_maps == C


def test_bad_arguments_return_error_codes_not_crashes():
    from deepcgp_b200 import _lib
    lib = _lib.lib
    assert lib.dcgp_view_geometry(4, 4, 1, 5, 1, None, None, None, None) == _lib.DCGP_ERR_ARG
    assert b"geometry" in lib.dcgp_last_error()
    assert lib.dcgp_kuu(None, 4, 4, 1.0, 1.0, 0.0, None, None) == _lib.DCGP_ERR_ARG
    assert lib.dcgp_cholesky(None, 4, None, 0, None, None) == _lib.DCGP_ERR_ARG
    d = _lib.LayerDesc(_lib.LAYER_CONV, 8, 8, 1, 3, 1, 0, 1, 0, 1.0, 1.0, 1e-3)   # M = 0
    assert lib.dcgp_prepare_bytes(d) == 0
    assert lib.dcgp_layer_prepare(d, None, None, None, None, 0, None, None, None, 0, None, None) == _lib.DCGP_ERR_ARG
    with pytest.raises(ValueError):
        _lib.check(_lib.DCGP_ERR_ARG)
    with pytest.raises(MemoryError):
        _lib.check(_lib.DCGP_ERR_WORKSPACE)


def test_workspace_queries_are_host_only_and_monotone():
    from deepcgp_b200 import _lib
    lib = _lib.lib
    d1 = _lib.LayerDesc(_lib.LAYER_CONV, 32, 32, 3, 5, 2, 512, 10, 0, 5.0, 5.0, 1e-3)
    d2 = _lib.LayerDesc(_lib.LAYER_CONV, 32, 32, 3, 5, 2, 1024, 10, 0, 5.0, 5.0, 1e-3)
    assert 0 < lib.dcgp_prepare_bytes(d1) < lib.dcgp_prepare_bytes(d2)
    assert 0 < lib.dcgp_prepare_workspace_bytes(d1) < lib.dcgp_prepare_workspace_bytes(d2)
    assert 0 < lib.dcgp_apply_workspace_bytes(d1, 256, 10) < lib.dcgp_apply_workspace_bytes(d1, 2560, 1)
    assert lib.dcgp_conditional_workspace_bytes(196, 512, 64, 10) > 196 * 64 * 512 * 4
    assert lib.dcgp_cholesky_workspace_bytes(512) >= 8 * 64 * 64 * 8


def test_cpu_tensors_are_rejected_loudly():
    import torch
    from deepcgp_b200 import _lib
    with pytest.raises(ValueError, match="no CPU path"):
        _lib.ptr(torch.zeros(3))


def test_minibatch_is_seeded_and_covers_the_data():
    from deepcgp_b200.dgp import Minibatch
    X = np.arange(10)
    a, b = Minibatch(X, 3, seed=0), Minibatch(X, 3, seed=0)
    ia = [a.next_indices() for _ in range(6)]
    ib = [b.next_indices() for _ in range(6)]
    assert all((x == y).all() for x, y in zip(ia, ib))
    assert sorted(np.concatenate(ia[:3]).tolist()) == sorted(set(np.concatenate(ia[:3]).tolist()))


def test_new_entry_points_validate_arguments_on_the_host():
    """Entry points added for the backward split, the prediction path, the counter-based noise and the batched GEMM reject
    bad arguments before touching the device."""
    from deepcgp_b200 import _lib
    lib = _lib.lib
    d = _lib.LayerDesc(_lib.LAYER_CONV, 8, 8, 1, 3, 1, 4, 2, 0, 1.0, 1.0, 1e-3)
    # phases outside 1..3 / null buffers
    assert lib.dcgp_layer_backward_phases(d, None, None, None, None, None, 1, 1, None, None, None, None, None, None, None,
                                          None, 0, 3, None) == _lib.DCGP_ERR_ARG
    assert lib.dcgp_multiclass_predict(None, None, None, 1, 1, 10, 1e-3, None, None, None, None) == _lib.DCGP_ERR_ARG
    assert lib.dcgp_randn(None, 1, 1, 1, 1, 0, 0, 0, 0, None) == _lib.DCGP_ERR_ARG
    assert lib.dcgp_bgemm_nt(None, None, None, 1, 4, 4, 4, 0, 0, None, 0, None) == _lib.DCGP_ERR_ARG
    assert lib.dcgp_bgemm_workspace_bytes(10, 512, 512, 512) > 4 * 10 * 512 * 512 * 2
    assert lib.dcgp_bgemm_workspace_bytes(0, 1, 1, 1) == 0
    off, ld = C.c_size_t(), C.c_int()
    assert lib.dcgp_prepare_layout(d, C.byref(off), C.byref(ld)) == _lib.DCGP_OK
    assert ld.value == 64 and 0 < off.value < lib.dcgp_prepare_bytes(d)
