"""GPU dev tool for ncu: cfg3 layer 2 (14x14x10 -> 10x10x10, M=512, R=10, 2560 rows): prepare, apply, backward."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, deepcgp_b200 as D
from deepcgp_b200.grad import LayerBackward
cfg = bench.CONFIGS["cfg3"]; dev = torch.device("cuda:0")
layers = bench.synth_params(cfg); model = bench.build_model(layers, cfg["S"], dev)
layer = model.layers[1]
n_rows = cfg["S"] * cfg["batch"]
X = torch.randn((n_rows, 14 * 14 * 10), device=dev)
z = torch.randn((n_rows, layer.num_outputs), device=dev)
lb = LayerBackward(layer)
gm = torch.randn((n_rows, layer.num_outputs), device=dev) * 1e-3
gv = torch.randn((n_rows, layer.num_outputs), device=dev) * 1e-3
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    layer.prepare(); layer._hold = True
    layer._conditional(X, z=z)
    lb.t_sized(X, 1, gm, gv, True)
    layer._hold = False
torch.cuda.synchronize()
print("done")
