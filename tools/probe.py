"""GPU probe (dev tool): accuracy of SIMT vs TC paths against the float64 oracle + kernel timings."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deepcgp_b200 as D
from oracle import dcgp_oracle as O
from tests.test_gpu_parity import _synthetic_conv, build_conv, npy
from tests.util import parity_err

dev = torch.device("cuda:0")

def acc(cfg, N, trained=True):
    rng = np.random.RandomState(1234)
    lay = _synthetic_conv(rng, trained=trained, **cfg)
    X = rng.standard_normal((N, cfg["H"] * cfg["W"] * cfg["C"])).astype(np.float32)
    mref, vref = O.convlayer_conditional_ND_fast(X.astype(np.float64), lay)
    K = O.mo_Kuu(lay["Z"], 5.0, 5.0)
    out = {"cond": np.linalg.cond(K)}
    for algo in ("simt", "tc"):
        layer = build_conv(lay, algo)
        m, v = layer.conditional_ND(torch.as_tensor(X, device=dev))
        out[algo] = (parity_err(npy(m), mref, 5.0), parity_err(npy(v), vref, 5.0))
    return out

def timeit(fn, iters=5):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(iters):
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return min(ts), float(np.median(ts))

def perf(cfg, n_rows, algo):
    rng = np.random.RandomState(1)
    lay = _synthetic_conv(rng, trained=True, **cfg)
    layer = build_conv(lay, algo)
    X = torch.randn((n_rows, cfg["H"] * cfg["W"] * cfg["C"]), device=dev)
    layer.prepare(); torch.cuda.synchronize()
    tp = timeit(lambda: layer.prepare())
    layer._hold = True
    ta = timeit(lambda: layer._conditional(X))
    layer._hold = False
    return tp, ta

if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("acc", "all"):
        for name, cfg, N in [
            ("cfg2-L1 M128", dict(H=28, W=28, C=1, f=5, s=2, M=128, R=10), 6),
            ("cfg3-L1 M512", dict(H=32, W=32, C=3, f=5, s=2, M=512, R=10), 3),
            ("cfg3-L2 M512", dict(H=14, W=14, C=10, f=5, s=1, M=512, R=10), 3),
            ("M1024 small", dict(H=12, W=12, C=2, f=4, s=3, M=1024, R=4), 2),
            ("cfg4-L2 M1024", dict(H=14, W=14, C=10, f=5, s=1, M=1024, R=10), 2),
        ]:
            r = acc(cfg, N)
            print("%-14s cond(Kuu)=%.2e" % (name, r["cond"]))
            for a in ("simt", "tc"):
                (mn, me), (vn, ve) = r[a]
                print("   %-4s mean: normwise %.2e elem %.3f | var: normwise %.2e elem %.3f" % (a, mn, me, vn, ve))
    if what in ("perf", "all"):
        for name, cfg, n_rows in [
            ("cfg3-L1 (N=256 dedup)", dict(H=32, W=32, C=3, f=5, s=2, M=512, R=10), 256),
            ("cfg3-L2 (N'=2560)", dict(H=14, W=14, C=10, f=5, s=1, M=512, R=10), 2560),
        ]:
            for algo in ("tc", "simt"):
                tp, ta = perf(cfg, n_rows, algo)
                print("%-22s %-4s prepare %.3f ms (med %.3f) | apply %.3f ms (med %.3f)" % (name, algo, tp[0], tp[1], ta[0], ta[1]))
