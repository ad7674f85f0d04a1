"""Shared helpers for the test-suite: golden-file loading and the parity metric."""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

LAYER_KEYS = ("H", "W", "C", "f", "s", "M", "R", "white", "variance", "lengthscale", "Z", "q_mu", "q_sqrt",
              "patch_weights")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def golden_names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def layer_from_golden(g, i, kind):
    """oracle-style layer dict from the l<i>_ entries of a golden file."""
    lay = {"type": kind}
    for k in LAYER_KEYS:
        key = "l%d_%s" % (i, k)
        if key in g:
            v = g[key]
            lay[k] = v if v.ndim else v.item()
    for k in ("H", "W", "C", "f", "s", "M", "R"):
        lay[k] = int(lay[k])
    lay["white"] = bool(lay["white"])
    lay["variance"] = float(lay["variance"])
    lay["lengthscale"] = float(lay["lengthscale"])
    return lay


def layers_from_golden(g):
    n = int(g["n_layers"])
    return [layer_from_golden(g, i, "conv" if i < n - 1 else "svgp_conv") for i in range(n)]


def parity_err(x, ref, sigma2):
    """BASELINE.md section 3 parity metric: normwise relative error, and the worst element-wise excess
    over the allowance 1e-4*|ref| + 1e-4*sigma^2 (<= 1 passes)."""
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    normwise = np.max(np.abs(x - ref)) / max(np.max(np.abs(ref)), 1e-300)
    elem = np.max(np.abs(x - ref) / (1e-4 * np.abs(ref) + 1e-4 * sigma2))
    return normwise, elem


def assert_parity(x, ref, sigma2, what=""):
    """conditional mean/var gate: 1e-4 relative (north_star), metric per BASELINE.md section 3."""
    normwise, elem = parity_err(x, ref, sigma2)
    assert normwise <= 1e-4 and elem <= 1.0, "%s: normwise %.3e (<=1e-4), elementwise ratio %.3f (<=1)" % (
        what, normwise, elem)
