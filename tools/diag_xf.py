"""dS kernel (xf_gemm_kernel) of conv layer 2 of cfg3 in the benchmark state, live CUDA events (dcgp_kernel_ms(3))."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from deepcgp_b200 import _lib
from deepcgp_b200.grad import LayerBackward
cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "cfg3"]; dev = torch.device("cuda:0")
layers = bench.synth_params(cfg); model = bench.build_model(layers, cfg["S"], dev)
layer = model.layers[1]
n_rows = cfg["S"] * cfg["batch"]
rng = np.random.RandomState(11)
X0 = torch.as_tensor(rng.standard_normal((n_rows // cfg["S"], 3072)).astype(np.float32), device=dev)
Fs, _, _ = model.propagate(X0, S=cfg["S"])
X = Fs[0].reshape(n_rows, -1).contiguous()
z = torch.randn((n_rows, layer.num_outputs), device=dev)
lb = LayerBackward(layer)
gm = torch.randn((n_rows, layer.num_outputs), device=dev) * 1e-3
gv = torch.randn((n_rows, layer.num_outputs), device=dev) * 1e-3
_lib.lib.dcgp_set_kernel_timing(1)
t = {0: [], 2: [], 3: []}
for it in range(6):
    layer.prepare(); layer._hold = True
    layer._conditional(X, z=z)
    lb.t_sized(X, 1, gm, gv, True)
    layer._hold = False
    torch.cuda.synchronize()
    if it >= 2:
        for k in t:
            t[k].append(_lib.lib.dcgp_kernel_ms(k))
_lib.lib.dcgp_set_kernel_timing(0)
print("cond %.3f  dk %.3f  dq %.3f ms" % tuple(float(np.median(t[k])) for k in (0, 2, 3)))
