"""GPU dev tool for ncu: cfg3 conv layer 2 (14x14x10 -> 10x10x10, M=512, R=10, 2560 rows) in the benchmark state, on its
ACTUAL input (samples of layer 1): prepare, apply, backward.   python tools/profile_layer2.py [iterations]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, deepcgp_b200 as D
from deepcgp_b200.grad import LayerBackward
cfg = bench.CONFIGS["cfg3"]; dev = torch.device("cuda:0")
layers = bench.synth_params(cfg); model = bench.build_model(layers, cfg["S"], dev)
layer = model.layers[1]
n_rows = cfg["S"] * cfg["batch"]
rng = np.random.RandomState(11)
X0 = torch.as_tensor(rng.standard_normal((cfg["batch"], 3072)).astype(np.float32), device=dev)
Fs, _, _ = model.propagate(X0, S=cfg["S"])
X = Fs[0].reshape(n_rows, -1).contiguous()
z = torch.randn((n_rows, layer.num_outputs), device=dev)
lb = LayerBackward(layer)
gm = torch.randn((n_rows, layer.num_outputs), device=dev) * 1e-3
gv = torch.randn((n_rows, layer.num_outputs), device=dev) * 1e-3
torch.cuda.synchronize()
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    layer.prepare(); layer._hold = True
    layer._conditional(X, z=z)
    lb.t_sized(X, 1, gm, gv, True)
    layer._hold = False
torch.cuda.synchronize()
print("done")
