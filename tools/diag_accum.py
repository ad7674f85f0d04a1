"""Diagnostics for the tensor-core accumulation error (GPU).

1. `test_ill_conditioned_kuu`'s layer through both paths: normwise error of mean / var against the float64 oracle.
2. The rounding of the fp32 TMEM accumulator: C = A B^T through dcgp_bgemm_nt with positive operands whose split-fp16 planes
   are exact (values on a 2^-10 grid), so every deviation from the float64 product is accumulation rounding.  A negative
   signed mean of about -n_acc/2 ulp says round-toward-zero, a zero mean says round-to-nearest.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from deepcgp_b200 import _lib  # noqa: E402


def illcond():
    from oracle import dcgp_oracle as O
    from tests.test_gpu_parity import _synthetic_conv, build_conv
    from tests.util import parity_err
    rng = np.random.RandomState(1234)
    lay = _synthetic_conv(rng, 12, 12, 2, 4, 3, 1024, 4, trained=True)
    X = rng.standard_normal((2, 12 * 12 * 2)).astype(np.float32)
    mref, vref = O.convlayer_conditional_ND_fast(X.astype(np.float64), lay)
    for algo, precise in (("simt", 0), ("tc", 0), ("tc", 1)):
        _lib.lib.dcgp_set_precise_stage1(precise)
        mean, var = build_conv(lay, algo).conditional_ND(torch.as_tensor(X, device="cuda"))
        em = parity_err(mean.cpu().numpy(), mref, 5.0)[0]
        ev = parity_err(var.cpu().numpy(), vref, 5.0)[0]
        d = mean.cpu().numpy().astype(np.float64) - mref
        print("illcond %-4s precise=%d mean %.3e var %.3e | signed mean of (mean - ref) %.3e, rms %.3e" % (
            algo, precise, em, ev, d.mean(), d.std()))
    _lib.lib.dcgp_set_precise_stage1(-1)


def stage1_time():
    """Conditional (stage 1 + stage 2) of a conv layer at M = 512 / 1024, T = 256 000, with one and four accumulators."""
    import bench
    from tests.test_gpu_parity import _synthetic_conv, build_conv
    for M in (512, 1024):
        rng = np.random.RandomState(5)
        lay = _synthetic_conv(rng, 14, 14, 10, 5, 1, M, 10, trained=True)
        X = torch.as_tensor(rng.standard_normal((2560, 14 * 14 * 10)).astype(np.float32), device="cuda")
        layer = build_conv(lay, "tc")
        outs = {}
        for precise in (0, 1):
            _lib.lib.dcgp_set_precise_stage1(precise)
            _lib.lib.dcgp_set_kernel_timing(1)
            for _ in range(3):
                mean, var = layer.conditional_ND(X)
            torch.cuda.synchronize()
            ms = _lib.lib.dcgp_kernel_ms(0)
            outs[precise] = (mean.double().cpu(), var.double().cpu())
            print("M=%d precise=%d: conditional (both stages) %.3f ms" % (M, precise, ms))
        _lib.lib.dcgp_set_kernel_timing(0)
        dm = (outs[0][0] - outs[1][0]).abs().max() / outs[1][0].abs().max()
        dv = (outs[0][1] - outs[1][1]).abs().max() / outs[1][1].abs().max()
        print("M=%d: precise vs plain mean %.2e var %.2e" % (M, dm, dv))
    _lib.lib.dcgp_set_precise_stage1(-1)


def gemm_nt(A, B):
    batch, m, k = 1, A.shape[0], A.shape[1]
    n = B.shape[0]
    Cm = torch.empty((1, m, n), dtype=torch.float32, device="cuda")
    nbytes = _lib.lib.dcgp_bgemm_workspace_bytes(batch, m, n, k)
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    _lib.check(_lib.lib.dcgp_bgemm_nt(_lib.ptr(A), _lib.ptr(B), _lib.ptr(Cm), batch, m, n, k, 0, 0, _lib.ptr(ws), ws.numel(),
                                      _lib.stream()))
    return Cm[0]


def accum():
    rng = np.random.RandomState(0)
    for k in (64, 256, 1024):
        for signed in (False, True):
            # 11-bit values in [0.5, 1): hi plane exact, lo plane zero -> one exact fp16 product per term
            A = (512 + rng.randint(0, 512, size=(256, k))) / 1024.0
            B = (512 + rng.randint(0, 512, size=(256, k))) / 1024.0
            if signed:
                B = B * rng.choice([-1.0, 1.0], size=B.shape)
            A[0, 0] = 1.0
            B[0, 0] = 1.0          # pins the power-of-two scale of both operands
            ref = A @ B.T
            got = gemm_nt(torch.as_tensor(A, dtype=torch.float32, device="cuda"),
                          torch.as_tensor(B, dtype=torch.float32, device="cuda")).cpu().numpy().astype(np.float64)
            ulp = np.spacing(np.abs(ref).astype(np.float32)).astype(np.float64)
            e = (got - ref) / ulp * np.sign(ref)
            print("k=%4d signed=%d: error toward +|ref| in ulp(ref): mean %.2f  rms %.2f  max %.1f | rel rms %.2e" % (
                k, signed, e.mean(), e.std(), np.abs(e).max(), np.std((got - ref) / np.abs(ref).max())))


if __name__ == "__main__":
    torch.zeros(1, device="cuda")
    if "accum" in sys.argv[1:] or len(sys.argv) == 1:
        accum()
    illcond()
    if "time" in sys.argv[1:]:
        stage1_time()
