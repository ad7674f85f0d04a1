"""kuf_tc_kernel of conv layer 2 of cfg3 (T = 256 000, M = 512, L = 250) in the benchmark state, live CUDA events
(dcgp_kernel_ms(1)): on the layer's actual input (samples of layer 1) and on N(0,1) noise of the same shape."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from deepcgp_b200 import _lib

cfg = bench.CONFIGS["cfg3"]; dev = torch.device("cuda:0")
layers = bench.synth_params(cfg); model = bench.build_model(layers, cfg["S"], dev)
layer = model.layers[1]
n_rows = cfg["S"] * cfg["batch"]
rng = np.random.RandomState(11)
X0 = torch.as_tensor(rng.standard_normal((cfg["batch"], 3072)).astype(np.float32), device=dev)
Fs, _, _ = model.propagate(X0, S=cfg["S"])
Xa = Fs[0].reshape(n_rows, -1).contiguous()
Xn = torch.randn_like(Xa)
print("actual input: mean %.3f std %.3f ptr %% 32 = %d" % (float(Xa.mean()), float(Xa.std()), Xa.data_ptr() % 32))
layer.prepare(); layer._hold = True
_lib.lib.dcgp_set_kernel_timing(1)
for name, X in (("actual", Xa), ("noise", Xn)):
    for _ in range(3):
        layer._conditional(X)
    tk, tc = [], []
    for _ in range(7):
        layer._conditional(X)
        tk.append(_lib.lib.dcgp_kernel_ms(1)); tc.append(_lib.lib.dcgp_kernel_ms(0))
    print("%-6s kuf %.4f ms (min %.4f)  cond %.4f ms" % (name, float(np.median(tk)), min(tk), float(np.median(tc))))
# the same shapes with an untrained-like state (lengthscale 5 on 250-dimensional N(0,1) patches: K ~ 1e-5 sigma^2)
from tests.test_gpu_parity import _synthetic_conv, build_conv
lay = _synthetic_conv(np.random.RandomState(5), 14, 14, 10, 5, 1, 512, 10, trained=True)
layer2 = build_conv(lay, "tc")
layer2.prepare(); layer2._hold = True
for name, X in (("tinyK/noise", Xn), ("tinyK/actual", Xa)):
    for _ in range(3):
        layer2._conditional(X)
    tk = []
    for _ in range(7):
        layer2._conditional(X)
        tk.append(_lib.lib.dcgp_kernel_ms(1))
    print("%-12s kuf %.4f ms (min %.4f)" % (name, float(np.median(tk)), min(tk)))
_lib.lib.dcgp_set_kernel_timing(0)
