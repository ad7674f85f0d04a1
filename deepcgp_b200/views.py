"""Patch "views" -- host-side mirror of the reference's conv_gp/views.py:18-68 (FullView).

Same constructor, attributes and method names; arithmetic runs in libdcgp.so.  The layer path never
calls extract_patches*: patches are gathered inside the fused Kuf kernel.  They exist for API parity
(and tests, in the style of the reference's tests/test_views.py:27-29).
"""
import ctypes as C

import torch

from . import _lib


class View(object):
    """conv_gp/views.py:6-16."""

    def extract_patches_PNL(self, NHWC_X):
        raise NotImplementedError()

    def mean_view(self, NHWC_X, PNL_patches):
        return NHWC_X


class FullView(View):
    """conv_gp/views.py:18-68: all f x f patches at stride s, VALID, dilation 1."""

    def __init__(self, input_size, filter_size, feature_maps, stride=1):
        self.input_size = list(input_size)
        self.stride = int(stride)
        self.dilation = 1
        self.filter_size = int(filter_size)
        self.feature_maps = int(feature_maps)
        self.patch_shape = [self.filter_size, self.filter_size]
        oh, ow, p, l = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _lib.check(_lib.lib.dcgp_view_geometry(int(self.input_size[0]), int(self.input_size[1]), self.feature_maps,
                                               self.filter_size, self.stride, oh, ow, p, l))
        self.out_image_height, self.out_image_width = oh.value, ow.value
        self.patch_count = p.value      # views.py:60-63
        self.patch_length = l.value     # views.py:56-58

    def _extract(self, NHWC_X, layout):
        X = _lib.f32(NHWC_X)
        N, H, W, Cc = X.shape
        assert [H, W] == [int(v) for v in self.input_size[:2]] and Cc == self.feature_maps
        shape = (self.patch_count, N, self.patch_length) if layout == 0 else (N, self.patch_count, self.patch_length)
        out = torch.empty(shape, dtype=torch.float32, device=X.device)
        _lib.check(_lib.lib.dcgp_extract_patches(_lib.ptr(X), N, H, W, Cc, self.filter_size, self.stride, layout,
                                                 _lib.ptr(out), _lib.stream()))
        return out

    def extract_patches_PNL(self, NHWC_X):
        """views.py:40-44 -> [P, N, L]"""
        return self._extract(NHWC_X, 0)

    def extract_patches(self, NHWC_X):
        """views.py:46-54 -> [N, P, L]"""
        return self._extract(NHWC_X, 1)
