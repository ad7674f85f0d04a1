"""GPU diagnostic: Kdiag backward alone (q(u) = p(u) makes the Kzx path vanish)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from tests.test_gpu_parity import build_last, dev, npy
from tests.test_gpu_backward_pieces import _rbf64, _patches64
from deepcgp_b200.grad import LayerBackward
from oracle import dcgp_oracle as O

cfg = bench.CONFIGS["cfg3"]; layers = bench.synth_params(cfg)
rng = np.random.RandomState(31); N = 6
F = rng.standard_normal((N, 3072))
for lay in layers[:2]:
    m, v = O.convlayer_conditional_ND_fast(F, lay); F = m + rng.standard_normal(m.shape) * np.sqrt(v + 1e-3)
lay = dict(layers[2]); X32 = F.astype(np.float32)
M, R = lay["M"], lay["R"]
Lm = np.linalg.cholesky(O.mo_Kuu(lay["Z"], lay["variance"], lay["lengthscale"]))
lay["q_mu"] = np.zeros((M, R)); lay["q_sqrt"] = np.tile(Lm[None], (R, 1, 1))
g_mean = np.zeros((N, R), np.float32)
g_var = (rng.standard_normal((N, R)) * 2.0).astype(np.float32)
t = lambda a, rg=True: torch.tensor(np.asarray(a, dtype=np.float64), requires_grad=rg)
X, var, ls, w = t(X32), t(lay["variance"]), t(lay["lengthscale"]), t(lay["patch_weights"])
pat = _patches64(X, lay); P, L = pat.shape[1:]
Kpp = torch.stack([_rbf64(pat[n], pat[n], var, ls) for n in range(N)])
kdiag = (Kpp * (w[None, :] * w[:, None])[None]).sum((1, 2)) / (P * P)
obj = (t(g_var, False).sum(1) * kdiag).sum()
gx, gv, gl, gw = torch.autograd.grad(obj, [X, var, ls, w])
layer = build_last(lay, "tc")
Xd = torch.as_tensor(X32, device=dev())
layer.prepare(); layer._hold = True; layer._conditional(Xd)
lb = LayerBackward(layer)
gX = lb.t_sized(Xd, 1, torch.as_tensor(g_mean, device=dev()), torch.as_tensor(g_var, device=dev()), True)
torch.cuda.synchronize()
nw = lambda a, b: float(np.abs(np.asarray(a, dtype=np.float64).reshape(b.shape) - b).max() / np.abs(b).max())
print("variance got %.8e ref %.8e rel %.2e" % (float(lb.gscal[0]), float(gv), abs(float(lb.gscal[0]) - float(gv)) / abs(float(gv))))
print("lengthscale got %.8e ref %.8e rel %.2e" % (float(lb.gscal[1]), float(gl), abs(float(lb.gscal[1]) - float(gl)) / abs(float(gl))))
print("X normwise %.2e  w normwise %.2e" % (nw(npy(gX), gx.numpy()), nw(npy(lb.gw), gw.numpy())))
print("gQ max", float(lb.gQB.abs().max()))
