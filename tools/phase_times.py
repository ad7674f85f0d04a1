"""GPU dev tool: where does the ELBO step spend its time (device time via events, host time via perf_counter)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, deepcgp_b200 as D
from deepcgp_b200 import _lib
cfg = bench.CONFIGS["cfg3"]; dev = torch.device("cuda:0")
layers = bench.synth_params(cfg); model = bench.build_model(layers, cfg["S"], dev)
S, B = cfg["S"], cfg["batch"]
X = torch.randn((B, 3072), device=dev); Y = torch.randint(0, 10, (B,), device=dev, dtype=torch.int32)
zs = [torch.randn((S, B, l.num_outputs), device=dev) for l in model.layers]
eg = D.ElboGradient(model); opt = D.Adam(model, lr=1e-3)

def timeit(fn, n=5):
    fn(); fn(); torch.cuda.synchronize()
    ts, hs = [], []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); a.record(); fn(); b.record(); h = time.perf_counter() - t0
        torch.cuda.synchronize(); ts.append(a.elapsed_time(b)); hs.append(h * 1e3)
    return np.median(ts), np.median(hs)

print("forward            dev %.2f ms  host %.2f ms" % timeit(lambda: model._build_likelihood(X, Y, zs=zs)))
def fb():
    e, g = eg(X, Y, zs=zs); return g
print("forward+backward   dev %.2f ms  host %.2f ms" % timeit(fb))
g = fb()
print("adam               dev %.2f ms  host %.2f ms" % timeit(lambda: opt.step(g)))
# T-sized backward only of each layer
Fs, Fm, Fv = model._fwd
for i in (2, 1, 0):
    lb = eg.bwd[i]; first = i == 0
    Xin = X if first else Fs[i-1].reshape(S*B, -1)
    D_out = model.layers[i].num_outputs
    gm = torch.randn((S*B, D_out), device=dev) * 1e-3; gv = torch.randn((S*B, D_out), device=dev) * 1e-3
    print("layer %d t_sized bwd dev %.2f ms  host %.2f ms" % ((i,) + timeit(lambda: lb.t_sized(Xin, S if first else 1, gm, gv, not first))))
    print("layer %d m_only  bwd dev %.2f ms  host %.2f ms" % ((i,) + timeit(lambda: lb.m_only())))
