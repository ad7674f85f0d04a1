"""Emulates (NumPy, CPU) the split-fp16 tensor-core arithmetic of the conditional GEMMs to choose how many of the three
split products each GEMM needs.  x = hi + lo with hi = fp16(x*s), lo = fp16(x*s - hi); a product variant keeps a subset of
{hi*hi, hi*lo, lo*hi}; accumulation is exact here (fp32 TMEM accumulation adds ~1e-7 relative, measured on the GPU).

    python tools/precision_study.py [cfg3|cfg4] [layer]

Prints, for the chained conditional (a = Lm^-1 k;  G_r = C_r^T a;  var_r = sigma^2 - |a|^2 + |G_r|^2;  mean = alpha^T a), the
normwise error max|x - ref| / max|ref| of mean and var against float64 for each variant of stage 1 / stage 2."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def split(x, bound):
    e = np.frexp(bound)[1]
    s = 2.0 ** (14 - e)
    hi = (x * s).astype(np.float16).astype(np.float64)
    lo = (x * s - hi).astype(np.float16).astype(np.float64)
    return hi / s, lo / s


def prod(A, B, variant):
    """A [m,k], B [k,n] -> A @ B with the chosen split products."""
    ah, al = split(A, np.abs(A).max())
    bh, bl = split(B, np.abs(B).max())
    out = ah @ bh
    if variant in ("3", "a_full"):
        out = out + al @ bh
    if variant in ("3", "b_full"):
        out = out + ah @ bl
    return out


def study(layers, li, n_img=2, seed=3, state="bench"):
    rng = np.random.RandomState(seed)
    lay = layers[li]
    # input of layer li: probe propagation with the float64 oracle
    from oracle import dcgp_oracle as O
    X = rng.standard_normal((n_img, layers[0]["H"] * layers[0]["W"] * layers[0]["C"]))
    F = X
    for l in layers[:li]:
        m, v = O.convlayer_conditional_ND_fast(F, l)
        F = m + rng.standard_normal(m.shape) * np.sqrt(v + 1e-3)
    M, R = lay["M"], lay["R"]
    pat = bench._np_patches(F.reshape(-1, lay["H"], lay["W"], lay["C"]), lay["f"], lay["s"])
    K = lay["variance"] * np.exp(-0.5 * bench._np_sqdist(pat, lay["Z"]) / lay["lengthscale"] ** 2)       # [T, M]
    Kuu = O.mo_Kuu(lay["Z"], lay["variance"], lay["lengthscale"])
    Lm = np.linalg.cholesky(Kuu)
    Linv = np.linalg.inv(Lm)
    alpha = Linv @ lay["q_mu"]
    C = np.stack([Linv @ np.tril(lay["q_sqrt"][r]) for r in range(R)])                                  # [R, M, M]
    a_ref = K @ Linv.T                                                                                   # [T, M]
    G_ref = np.stack([a_ref @ C[r] for r in range(R)])                                                   # [R, T, M]
    var_ref = lay["variance"] - (a_ref ** 2).sum(1)[None] + (G_ref ** 2).sum(2)
    mean_ref = a_ref @ alpha
    print("layer %d: T=%d M=%d cond(Kuu)=%.2e median K/s2=%.3f |var| max %.3f min %.3f, |mean| max %.3f" % (
        li, K.shape[0], M, np.linalg.cond(Kuu), np.median(K) / lay["variance"], var_ref.max(), var_ref.min(), np.abs(mean_ref).max()))
    nw = lambda x, r: np.abs(x - r).max() / np.abs(r).max()
    for v1 in ("3",):
        a = prod(K, Linv.T, v1)
        print("  stage1 %-7s: a normwise %.2e, |a|^2 err (rel to s2) %.2e" % (v1, nw(a, a_ref), np.abs((a ** 2).sum(1) - (a_ref ** 2).sum(1)).max() / lay["variance"]))
        for v2 in ("3", "a_full", "b_full", "1"):
            Ccat = np.concatenate([C[r] for r in range(R)], axis=1)                                      # [M, R*M], common scale
            G = prod(a, Ccat, v2).reshape(a.shape[0], R, M).transpose(1, 0, 2)
            var = lay["variance"] - (a ** 2).sum(1)[None] + (G ** 2).sum(2)
            mean = prod(a, alpha, "3")
            print("    stage2 %-7s: var normwise %.2e   mean normwise %.2e" % (v2, nw(var, var_ref), nw(mean, mean_ref)))
    return K, Linv, C, a_ref


if __name__ == "__main__":
    cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
    layers = bench.synth_params(bench.CONFIGS[cfg])
    for li in ([int(sys.argv[2])] if len(sys.argv) > 2 else [0, 1]):
        study(layers, li)
