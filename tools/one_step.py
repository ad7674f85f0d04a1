"""GPU dev tool for ncu: N full ELBO steps (forward + backward + Adam) of cfg3."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, deepcgp_b200 as D
cfg = bench.CONFIGS[os.environ.get("CFG", "cfg3")]; dev = torch.device("cuda:0")
layers = bench.synth_params(cfg); model = bench.build_model(layers, cfg["S"], dev)
S, B = cfg["S"], cfg["batch"]
X = torch.randn((B, cfg["H"] * cfg["W"] * cfg["C"]), device=dev); Y = torch.randint(0, 10, (B,), device=dev, dtype=torch.int32)
zs = [torch.randn((S, B, l.num_outputs), device=dev) for l in model.layers]
eg = D.ElboGradient(model); opt = D.Adam(model, lr=1e-3)
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    e, g = eg(X, Y, zs=zs); opt.step(g)
torch.cuda.synchronize()
print("done")
