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
lay = layers[2]; M, R = lay["M"], lay["R"]
for case in ("real", "const"):
    X32 = F.astype(np.float32) if case == "real" else np.ones((N, 1000), np.float32) * 0.3
    g_mean = np.zeros((N, R), np.float32); g_var = np.zeros((N, R), np.float32); g_var[:, 0] = 1.0
    t = lambda a, rg=True: torch.tensor(np.asarray(a, dtype=np.float64), requires_grad=rg)
    X, Z, var, ls, w = t(X32, False), t(lay["Z"], False), t(lay["variance"]), t(lay["lengthscale"]), t(lay["patch_weights"], False)
    Kuu = _rbf64(Z, Z, lay["variance"], lay["lengthscale"]) + 1e-3 * torch.eye(M, dtype=torch.float64)
    Kinv = torch.linalg.inv(Kuu); B = Kinv @ torch.tril(t(lay["q_sqrt"], False))
    Q = torch.cat([Kinv[None], B @ B.transpose(1, 2)])
    pat = _patches64(X, lay); P, L = pat.shape[1:]
    K = _rbf64(pat.reshape(N * P, L), Z, var, ls)
    Kzx = (K.reshape(N, P, M) * w[None, :, None]).sum(1) / P
    Kpp = torch.stack([_rbf64(pat[n], pat[n], var, ls) for n in range(N)])
    kdiag = (Kpp * (w[None, :] * w[:, None])[None]).sum((1, 2)) / (P * P)
    quad = torch.einsum("tm,bmn,tn->tb", Kzx, Q, Kzx)
    o_kd = kdiag.sum(); o_q = (-quad[:, 0] + quad[:, 1]).sum()
    gv_kd, gl_kd = torch.autograd.grad(o_kd, [var, ls], retain_graph=True)
    gv_q, gl_q = torch.autograd.grad(o_q, [var, ls])
    layer = build_last(lay, "tc")
    Xd = torch.as_tensor(X32, device=dev())
    layer.prepare(); layer._hold = True; layer._conditional(Xd)
    lb = LayerBackward(layer)
    lb.t_sized(Xd, 1, torch.as_tensor(g_mean, device=dev()), torch.as_tensor(g_var, device=dev()), True)
    torch.cuda.synchronize()
    lv, ll = float(lb.gscal[0]), float(lb.gscal[1])
    print(case, "ref kd var %.8e ls %.8e | ref quad var %.8e ls %.8e" % (float(gv_kd), float(gl_kd), float(gv_q), float(gl_q)))
    print(case, "lib total var %.8e ls %.8e -> lib kd piece (total - ref quad) var %.8e (rel err %.2e) ls %.8e (rel err %.2e)" % (
        lv, ll, lv - float(gv_q), (lv - float(gv_q) - float(gv_kd)) / float(gv_kd), ll - float(gl_q),
        (ll - float(gl_q) - float(gl_kd)) / (abs(float(gl_kd)) + 1e-30)))
    kd_lib = layer.kern.Kdiag(Xd)
    print(case, "forward kdiag rel err", float(np.abs(npy(kd_lib) - kdiag.detach().numpy()).max() / kdiag.detach().numpy().max()))
