"""GPU diagnostic: last layer (SVGP ConvKernel) direct variance / lengthscale gradient, Kzx path vs Kdiag path."""
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
lay = layers[2]; X32 = F.astype(np.float32)
M, R = lay["M"], lay["R"]
for mode in ("kzx_only", "kdiag_only", "both"):
    g_mean = (rng.standard_normal((N, R)) * 3.0).astype(np.float32)
    g_var = (rng.standard_normal((N, R)) * 2.0).astype(np.float32)
    if mode == "kzx_only":
        g_var -= g_var.mean(1, keepdims=True); g_var[:, -1] -= g_var.sum(1)           # rows sum to ~0 -> no Kdiag path
    if mode == "kdiag_only":
        g_mean[:] = 0; g_var[:] = 0
    t = lambda a, rg=True: torch.tensor(np.asarray(a, dtype=np.float64), requires_grad=rg)
    X, Z, var, ls, w = t(X32, False), t(lay["Z"], False), t(lay["variance"]), t(lay["lengthscale"]), t(lay["patch_weights"], False)
    Kuu = _rbf64(Z, Z, lay["variance"], lay["lengthscale"]) + 1e-3 * torch.eye(M, dtype=torch.float64)
    Kinv = torch.linalg.inv(Kuu); B = Kinv @ torch.tril(t(lay["q_sqrt"], False))
    Q = torch.cat([Kinv[None], B @ B.transpose(1, 2)]); beta = Kinv @ t(lay["q_mu"], False)
    pat = _patches64(X, lay); P, L = pat.shape[1:]
    K = _rbf64(pat.reshape(N * P, L), Z, var, ls)
    Kzx = (K.reshape(N, P, M) * w[None, :, None]).sum(1) / P
    Kpp = torch.stack([_rbf64(pat[n], pat[n], var, ls) for n in range(N)])
    kdiag = (Kpp * (w[None, :] * w[:, None])[None]).sum((1, 2)) / (P * P)
    quad = torch.einsum("tm,bmn,tn->tb", Kzx, Q, Kzx)
    gk = np.ones(N) if mode == "kdiag_only" else None
    if mode == "kdiag_only":
        g_var[:, 0] = 1.0
        obj = kdiag.sum() - quad[:, 0].sum() + quad[:, 1].sum()
    else:
        obj = (t(g_mean, False) * (Kzx @ beta)).sum() + (t(g_var, False) * (kdiag[:, None] - quad[:, :1] + quad[:, 1:])).sum()
    gv, gl = torch.autograd.grad(obj, [var, ls])
    # pieces of the reference by path
    obj_kd = (t(g_var, False).sum(1) * kdiag).sum()
    gv_kd, gl_kd = torch.autograd.grad(obj_kd, [var, ls], retain_graph=False) if False else (None, None)
    layer = build_last(lay, "tc")
    Xd = torch.as_tensor(X32, device=dev())
    layer.prepare(); layer._hold = True; layer._conditional(Xd)
    lb = LayerBackward(layer)
    lb.t_sized(Xd, 1, torch.as_tensor(g_mean, device=dev()), torch.as_tensor(g_var, device=dev()), True)
    torch.cuda.synchronize()
    print(mode, "variance: got %.8e ref %.8e rel %.2e | lengthscale: got %.8e ref %.8e rel %.2e" % (
        float(lb.gscal[0]), float(gv), abs(float(lb.gscal[0]) - float(gv)) / abs(float(gv)),
        float(lb.gscal[1]), float(gl), abs(float(lb.gscal[1]) - float(gl)) / abs(float(gl))))
