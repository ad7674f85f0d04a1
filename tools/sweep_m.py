"""BASELINE config 5: inducing sweep M in {64..1024} on synthetic 32x32x3, f=5, 1 GPU: time of the M-only work
(Kuu + Cholesky + inverse + operand build = dcgp_layer_prepare) and Kuf throughput (kuf_tc_kernel, live CUDA events)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deepcgp_b200 as D
from deepcgp_b200 import _lib
from tests.test_gpu_parity import _synthetic_conv, build_conv

dev = torch.device("cuda:0")
rows = []
for stride, P in ((2, 196), (1, 784)):
    for M in (64, 128, 256, 512, 1024):
        rng = np.random.RandomState(M)
        lay = _synthetic_conv(rng, 32, 32, 3, 5, stride, M, 10, trained=True)
        layer = build_conv(lay, "tc")
        X = torch.randn((256, 3072), device=dev)
        for _ in range(3):
            layer.prepare()
        torch.cuda.synchronize()
        tp = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); layer.prepare(); b.record(); torch.cuda.synchronize(); tp.append(a.elapsed_time(b))
        layer._hold = True
        _lib.lib.dcgp_set_kernel_timing(1)
        for _ in range(3):
            layer._conditional(X)
        tk, tc = [], []
        for _ in range(5):
            layer._conditional(X); tk.append(_lib.lib.dcgp_kernel_ms(1)); tc.append(_lib.lib.dcgp_kernel_ms(0))
        _lib.lib.dcgp_set_kernel_timing(0)
        layer._hold = False
        T = 256 * P
        Mp = (M + 63) // 64 * 64
        kuf_bytes = 4.0 * T * Mp + 4.0 * 256 * 3072
        rows.append(dict(stride=stride, P=P, M=M, prepare_ms=float(np.median(tp)), kuf_ms=float(np.median(tk)),
                         kuf_gbs=kuf_bytes / (np.median(tk) * 1e-3) / 1e9, kuf_gflops=2.0 * T * M * 75 / (np.median(tk) * 1e-3) / 1e9,
                         cond_ms=float(np.median(tc))))
        print(json.dumps(rows[-1]))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/sweep_m.json", "w"), indent=1)
