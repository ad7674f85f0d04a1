#!/usr/bin/env python
"""bench.py -- ELBO-step images/sec for the 3-layer CIFAR-10 DCGP (M=512, 5x5 patches, batch 256/GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg3]

One process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE); rank 0 prints ONE JSON line.
Workload (BASELINE.json configs[2]): 32x32x3 synthetic images, ConvLayer(f=5,s=2,M=512,R=10) ->
ConvLayer(f=5,s=1,M=512,R=10) -> SVGP(ConvKernel f=5,s=1,M=512,10 classes), S=10, sigma^2=5, l=5, jitter=1e-3,
non-white, trained-like variational state.  Images shard over ranks (weak scaling: 256 images per GPU).

A "step" is what the library implements of the ELBO step today -- see `config.step` in the JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

if "reference" in sys.argv:
    # CPU arm: use every host core even under torchrun (which exports OMP_NUM_THREADS=1); must happen before NumPy /
    # torch load their BLAS / OpenMP runtimes
    try:
        _n = str(len(os.sched_getaffinity(0)))
    except Exception:
        _n = str(os.cpu_count() or 1)
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = _n

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (H, W, C, [(f, s, M, R) conv layers], (f, s, M) last layer, batch per GPU, S)
    "cfg2": dict(H=28, W=28, C=1, conv=[(5, 2, 128, 10)], last=(5, 1, 128), batch=128, S=10,
                 desc="MNIST 2-layer DCGP M=128 batch 128"),
    "cfg3": dict(H=32, W=32, C=3, conv=[(5, 2, 512, 10), (5, 1, 512, 10)], last=(5, 1, 512), batch=256, S=10,
                 desc="CIFAR-10 3-layer DCGP M=512 batch 256"),
    "cfg4": dict(H=32, W=32, C=3, conv=[(5, 2, 1024, 10), (5, 1, 1024, 10)], last=(5, 1, 1024), batch=64, S=10,
                 desc="CIFAR-10 3-layer DCGP M=1024 batch 512 over 8 GPUs (64/GPU)"),
}
NUM_DATA = 50000
SIGMA2, LENGTHSCALE, JITTER = 5.0, 5.0, 1e-3


# ----------------------------------------------------------------------------------------------- synthetic model
def _np_patches(X, f, s):
    """[N,H,W,C] -> [N*P, L], (dy,dx,c) order, p = oy*OW+ox (views.py:32-54)."""
    N, H, W, C = X.shape
    oh, ow = (H - f) // s + 1, (W - f) // s + 1
    out = np.empty((N, oh, ow, f, f, C))
    for dy in range(f):
        for dx in range(f):
            out[:, :, :, dy, dx, :] = X[:, dy:dy + (oh - 1) * s + 1:s, dx:dx + (ow - 1) * s + 1:s, :]
    return out.reshape(N * oh * ow, f * f * C)


def _np_sqdist(A, B):
    return np.maximum((A * A).sum(1)[:, None] + (B * B).sum(1)[None, :] - 2.0 * A @ B.T, 0.0)


def synth_params(cfg, seed=1236, n_probe=16):
    """Seeded synthetic model in a NON-DEGENERATE trained-like state (pure NumPy, identical on every rank and in both arms).

    A handful of probe images is propagated through the stack while it is built, so that every layer's inducing patches
    come from that layer's ACTUAL input (kernels.py:147-164: the reference clusters patches of the layer's real inputs):
      * X ~ N(0,1); layer 1: variance 5, lengthscale 5 (models.py:115-116);
      * Z_l = M patches drawn from the probe input of layer l + 10 % noise;
      * deeper layers: a DS-DGP layer's samples carry the prior variance sigma^2 - k^T Kuu^-1 k (std up to sqrt(5)) whatever
        q(u) is, so with the initial lengthscale 5 and L = 250 every Kuf entry of layers >= 2 is ~5 exp(-40) = 0 (the state
        the round-1 benchmark timed).  A trained model adapts the lengthscale to its input scale; here it is set by the
        median heuristic on the layer's probe input: median(Kuf) = 0.1 sigma^2;
      * q(u) is a contraction of the prior: q_mu = Lm v, q_sqrt_r = Lm T_r with v ~ 0.7 N(0,1),
        T_r = 0.5 I + tril(N(0, 0.25/M)) -- mean and variance of every layer stay O(sigma^2).
    Each layer dict also carries `kuf_median_rel` = median(Kuf)/sigma^2 on the probe input (checked on the GPU by bench.py)."""
    rng = np.random.RandomState(seed)
    layers = []
    h, w, c = cfg["H"], cfg["W"], cfg["C"]
    F = rng.standard_normal((n_probe, h, w, c))
    specs = [(f, s, M, R, "conv") for (f, s, M, R) in cfg["conv"]] + [cfg["last"] + (10, "svgp_conv")]
    for li, (f, s, M, R, kind) in enumerate(specs):
        L = f * f * c
        oh, ow = (h - f) // s + 1, (w - f) // s + 1
        P = oh * ow
        pat = _np_patches(F, f, s)                                      # [n_probe*P, L]
        idx = rng.choice(pat.shape[0], M, replace=pat.shape[0] < M)
        Z = pat[idx] + 0.1 * pat.std() * rng.standard_normal((M, L))
        d2 = _np_sqdist(pat, Z)
        ls = LENGTHSCALE if li == 0 else float(np.sqrt(np.median(d2) / (2.0 * np.log(10.0))))
        K = SIGMA2 * np.exp(-0.5 * d2 / ls ** 2)                        # [T, M]
        Kuu = SIGMA2 * np.exp(-0.5 * _np_sqdist(Z, Z) / ls ** 2) + JITTER * np.eye(M)
        Lm = np.linalg.cholesky(Kuu)
        v = 0.7 * rng.standard_normal((M, R))
        Tr = np.tril(rng.standard_normal((R, M, M)) * (0.5 / np.sqrt(M))) + 0.5 * np.eye(M)
        lay = dict(type=kind, H=h, W=w, C=c, f=f, s=s, M=M, R=R, white=False, variance=SIGMA2, lengthscale=ls,
                   Z=Z, q_mu=Lm @ v, q_sqrt=np.matmul(Lm[None], Tr), kuf_median_rel=float(np.median(K) / SIGMA2))
        if kind == "svgp_conv":
            lay["patch_weights"] = np.ones(P)
        layers.append(lay)
        if kind == "conv":      # probe samples of this layer = the next layer's input (conditionals.py:29-65, DS/utils.py:41)
            a = np.linalg.solve(Lm, K.T)                                # [M, T]   (tiny: a dense solve is fine here)
            mean = a.T @ v                                              # alpha = Lm^-1 q_mu = v
            var = SIGMA2 - (a * a).sum(0)[:, None] + np.stack([((Tr[r].T @ a) ** 2).sum(0) for r in range(R)], 1)
            smp = mean + rng.standard_normal(mean.shape) * np.sqrt(np.maximum(var, 0.0) + JITTER)
            F = smp.reshape(n_probe, oh, ow, R)
        h, w, c = oh, ow, R
    return layers


def build_model(layers, S, device):
    import deepcgp_b200 as D
    built = []
    for lay in layers:
        kern = D.RBF(lay["f"] ** 2 * lay["C"], variance=lay["variance"], lengthscales=lay["lengthscale"])
        feat = D.PatchInducingFeatures(lay["Z"])
        if lay["type"] == "conv":
            view = D.FullView((lay["H"], lay["W"]), lay["f"], lay["C"], lay["s"])
            built.append(D.ConvLayer(kern, D.Zero(), feature=feat, view=view, white=lay["white"], gp_count=lay["R"],
                                     q_mu=lay["q_mu"], q_sqrt=lay["q_sqrt"], device=device))
        else:
            view = D.FullView((lay["H"], lay["W"], lay["C"]), lay["f"], lay["C"], lay["s"])
            built.append(D.SVGP_Layer(D.ConvKernel(kern, view, lay["patch_weights"]), lay["R"], D.Zero(lay["R"]),
                                      feature=feat, white=lay["white"], q_mu=lay["q_mu"], q_sqrt=lay["q_sqrt"],
                                      device=device))
    return D.DGP_Base(np.zeros((1, 1), np.float32), np.zeros((1, 1)), D.MultiClass(10), built, num_samples=S,
                      num_data=NUM_DATA, device=device)


def algorithmic_flops(cfg, n_img):
    """SURVEY.md 8d canonical forward flops per step (per GPU), and the share of the conditional-GEMM kernel."""
    S = cfg["S"]
    h, w, c = cfg["H"], cfg["W"], cfg["C"]
    total, cond_layers = 0.0, []
    specs = [(f, s, M, R, "conv") for (f, s, M, R) in cfg["conv"]] + [cfg["last"] + (10, "svgp_conv")]
    for i, (f, s, M, R, kind) in enumerate(specs):
        L = f * f * c
        oh, ow = (h - f) // s + 1, (w - f) // s + 1
        P = oh * ow
        n_eff = n_img if i == 0 else S * n_img
        f_m = 2.0 * M * M * L + M ** 3 / 3.0 + R * M ** 3 / 3.0 + M * M * R
        if kind == "conv":
            T = P * n_eff
            cond = T * (M * M + 2.0 * M * R + R * M * M + 2.0 * M * (R + 1))
            total += T * 2.0 * M * L + cond + f_m
        else:
            T = n_eff
            cond = T * (M * M + 2.0 * M * R + R * M * M)
            total += n_eff * (2.0 * M * L * P + 2.0 * L * P * P) + cond + f_m
        cond_layers.append(cond)
        h, w, c = oh, ow, R
    return total, cond_layers


# ----------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler(threading.Thread):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self._halt = gpu_index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        sm = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


# ----------------------------------------------------------------------------------------------- reference / CPU arm
def cpu_sample(cfg, layers, n_img, seed=99):
    """Seeded bounded sample of the workload: n_img images (float32-representable, so both arms see identical inputs),
    labels and the N(0,1) draws of every layer."""
    rng = np.random.RandomState(seed)
    S = cfg["S"]
    X = rng.standard_normal((n_img, cfg["H"] * cfg["W"] * cfg["C"])).astype(np.float32)
    Y = rng.randint(0, 10, size=(n_img, 1))
    zs = []
    h, w = cfg["H"], cfg["W"]
    for lay in layers:
        oh, ow = (h - lay["f"]) // lay["s"] + 1, (w - lay["f"]) // lay["s"] + 1
        D = oh * ow * lay["R"] if lay["type"] == "conv" else lay["R"]
        zs.append(rng.standard_normal((S, n_img, D)).astype(np.float32))
        h, w = oh, ow
    return X, Y, zs


def cpu_reference_rate(cfg, layers, n_img, seed=99, forward_only=False, keep=None):
    """The reference's CPU implementation cannot run here (TensorFlow/GPflow absent).  Timed instead,
    on a bounded sample of the same workload (`n_img` images with all S samples each, full model):
      * ELBO step  : the float64 torch-CPU restatement with autograd (oracle/dcgp_oracle_torch.py): forward + backward,
                     which is what TensorFlow executes per optimiser step (tf.gradients; Adam itself is negligible);
      * forward only: the float64 NumPy/SciPy oracle (oracle/dcgp_oracle.py, single-solve form).
    Both use all host BLAS threads.  Returns (images/s, seconds, elbo, grads or None)."""
    from oracle import dcgp_oracle as O
    S = cfg["S"]
    X, Y, zs = cpu_sample(cfg, layers, n_img, seed)
    X64, zs64 = X.astype(np.float64), [z.astype(np.float64) for z in zs]
    grads = None
    t0 = time.perf_counter()
    if forward_only:
        elbo = O.dgp_elbo(layers, X64, Y, zs64, NUM_DATA, S, JITTER, fast=True)
    else:
        from oracle import dcgp_oracle_torch as OT
        elbo, grads = OT.elbo_and_grads(layers, X64, Y, zs64, NUM_DATA, S, JITTER, keep=keep)
    dt = time.perf_counter() - t0
    return n_img / dt, dt, float(elbo), grads


def host_threads():
    """Threads the CPU arm may use: every core this process is allowed on (torchrun exports OMP_NUM_THREADS=1, which is
    overridden at the top of this file for --impl reference and here for the in-line cpu_baseline leg)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        import torch
        torch.set_num_threads(n)
    except Exception:
        pass
    return n


def workload_config(args, cfg, layers, B, n_global, scaling):
    """`config` of the JSON line -- identical for both arms (the CPU arm states its bounded sample in cpu_baseline.sample)."""
    return {"workload": args.config + ": " + cfg["desc"],
            "step": "forward ELBO only" if args.forward_only else
                    "forward ELBO + backward + gradient all-reduce + Adam (all trainables)",
            "images_per_gpu": B, "global_batch": n_global, "num_samples": cfg["S"], "scaling": scaling,
            "state": "trained-like, non-degenerate: Z from the layer's actual (probe) input, median-heuristic lengthscale "
                     "for layers >= 2, q(u) a contraction of the prior (bench.py:synth_params)",
            "variance": SIGMA2, "lengthscales": [round(l["lengthscale"], 4) for l in layers],
            "kuf_median_rel_numpy": [round(l["kuf_median_rel"], 4) for l in layers]}


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_threads()
    layers = synth_params(cfg)
    n_img = args.ref_images
    rates = []
    for i in range(args.warmup + args.steps):
        r, dt, _, _ = cpu_reference_rate(cfg, layers, n_img, seed=99 + i, forward_only=args.forward_only)
        if i >= args.warmup:
            rates.append((r, dt))
    value = float(np.mean([r for r, _ in rates]))
    ms = float(np.mean([dt for _, dt in rates])) * 1e3
    what = "forward ELBO (NumPy/SciPy float64)" if args.forward_only else "forward + backward (torch-CPU float64 autograd)"
    sample = ("each step = %d images x S=%d of the workload through the full model, %s, %d host threads; images/s = %d / step time; "
              "oracle port (TensorFlow/GPflow not installable)" % (n_img, cfg["S"], what, cores, n_img))
    B = cfg["batch"]
    line = {"impl": "reference", "metric": "ELBO-step images/sec", "value": value, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, cfg, layers, B, B * args.gpus, "weak"),
            "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------- our arm
def normwise(x, ref):
    x, ref = np.asarray(x, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return float(np.max(np.abs(x.reshape(ref.shape) - ref)) / max(float(np.max(np.abs(ref))), 1e-300))


def gpu_parity(cfg, layers, n_img, device, ref_elbo, ref_grads, ref_layers, seed=99):
    """Outside the timed region: the GPU model on the CPU arm's sample (same inputs, same draws) against the float64 oracle:
    ELBO, every layer's conditional mean / var (normwise max|d|/max|ref|) and every parameter gradient."""
    import torch
    import deepcgp_b200 as D
    S = cfg["S"]
    X, Y, zs = cpu_sample(cfg, layers, n_img, seed)
    model = build_model(layers, S, device)
    eg = D.ElboGradient(model)
    elbo, grads = eg(X, Y, zs=[torch.as_tensor(z, device=device) for z in zs])
    Fs, Fmeans, Fvars = model._fwd
    out = {"images": n_img, "elbo_rel": abs(float(elbo.item()) - ref_elbo) / abs(ref_elbo),
           "mean_rel": [normwise(m.cpu().numpy(), r[0]) for m, r in zip(Fmeans, ref_layers)],
           "var_rel": [normwise(v.cpu().numpy(), r[1]) for v, r in zip(Fvars, ref_layers)]}
    per, worst = {}, 0.0
    for i, (got, ref) in enumerate(zip(grads, ref_grads)):
        for k, v in ref.items():
            e = normwise(got[k].detach().cpu().numpy(), v)
            per["l%d.%s" % (i, k)] = e
            worst = max(worst, e)
    out["grad_rel"] = worst
    out["grad_rel_per_tensor"] = per
    out["metric"] = "max|x - ref| / max|ref| per tensor; ref = float64 oracle (torch autograd for gradients)"
    return out


def kuf_medians(model, cfg, device, n_img=4):
    """median(Kuf) / sigma^2 of every layer on its ACTUAL input (a few images propagated on the GPU): the benchmark state
    must not multiply zeros (every layer's K planes carry O(0.1 sigma^2) entries)."""
    import torch
    import deepcgp_b200 as D
    rng = np.random.RandomState(7)
    X = torch.as_tensor(rng.standard_normal((n_img, cfg["H"] * cfg["W"] * cfg["C"])).astype(np.float32), device=device)
    Fs, _, _ = model.propagate(X, S=1)
    meds, Fin = [], X
    for layer, F in zip(model.layers, Fs):
        v = layer._view
        img = Fin.reshape(n_img, int(v.input_size[0]), int(v.input_size[1]), v.feature_maps)
        mok = D.MultiOutputConvKernel(layer._base_kernel, 0, v.patch_count)
        K = mok.Kuf_images(layer.feature.Z, img, v.filter_size, v.stride)
        meds.append(float(K.median().item()) / float(layer._base_kernel.variance))
        Fin = F.reshape(n_img, -1)
    return meds


def run_ours(args, cfg):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL prints its version banner with printf when the communicator is created
        # (first collective), so file descriptor 1 points at stderr until that has happened
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            warm = torch.zeros(1, device=device)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    from deepcgp_b200 import _lib
    import deepcgp_b200 as D

    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=device)   # > 126 MB L2
    elbo_host = torch.empty(1, dtype=torch.float64).pin_memory()

    def measure(cfg_m, B, n_global, steps, warmup, e2e=True, sample_clocks=False):
        """Times `steps` ELBO steps of configuration cfg_m with B images on this rank (n_global over all ranks).
        Returns dict(ms, ms_e2e, launches, clocks, model, layers, elbo)."""
        S = cfg_m["S"]
        layers = synth_params(cfg_m)
        model = build_model(layers, S, device)
        rng = np.random.RandomState(4321 + rank)
        D_in = cfg_m["H"] * cfg_m["W"] * cfg_m["C"]
        n_batches = 4                                        # rotate distinct host batches (pinned), like a data loader
        hostX = [torch.from_numpy(rng.standard_normal((B, D_in)).astype(np.float32)).pin_memory() for _ in range(n_batches)]
        hostY = [torch.from_numpy(rng.randint(0, 10, size=(B,)).astype(np.int32)).pin_memory() for _ in range(n_batches)]
        devX = [x.to(device) for x in hostX]
        devY = [y.to(device) for y in hostY]

        # The N(0,1) draws of DS/layers.py:104 are part of the step: drawn inside the timed region, every step, by the
        # counter-based generator (indexed by the global image position, so every rank count sees the same noise).
        def draw():
            return model.draw_zs(B, n_global, rank * B)

        if args.sequential:
            eg = D.ElboGradient(model)
            opt = D.Adam(model, lr=args.lr)
            train_step = None
        else:
            train_step = D.TrainStep(model, lr=args.lr)

        def elbo_step(x, y):
            """One optimisation step: forward ELBO, backward, (all-reduce of the gradient), Adam update of every trainable.
            Default = grad.TrainStep (same arithmetic as ElboGradient + Adam.step, per-layer tails overlapped)."""
            if args.forward_only:
                return model._build_likelihood(x, y, zs=draw(), n_global=n_global)
            if args.sequential:
                e, grads = eg(x, y, zs=draw(), n_global=n_global)
                opt.step(grads)
                return e
            return train_step(x, y, zs=draw(), n_global=n_global)

        def step_resident(i):
            return elbo_step(devX[i % n_batches], devY[i % n_batches])

        def step_e2e(i):
            x = hostX[i % n_batches].to(device, non_blocking=True)
            y = hostY[i % n_batches].to(device, non_blocking=True)
            e = elbo_step(x, y)
            elbo_host.copy_(e.reshape(1), non_blocking=True)
            return e

        def timed(step_fn):
            for i in range(warmup):
                step_fn(i)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            torch.cuda.synchronize()
            for i in range(steps):
                flush.zero_()                                # L2 flush between timed iterations (not timed)
                ev[i][0].record()
                step_fn(i)
                ev[i][1].record()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            ms = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device=device)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)    # device time, max over ranks
            return float(ms.item()) / steps

        sampler = ClockSampler(local_rank) if (sample_clocks and rank == 0) else None
        if sampler:
            sampler.start()
        launches0 = _lib.lib.dcgp_launch_count()
        ms_res = timed(step_resident)
        launches = _lib.lib.dcgp_launch_count() - launches0
        ms_e2e = timed(step_e2e) if e2e else None
        clocks = sampler.stop() if sampler else None
        if train_step is not None and not args.forward_only:
            train_step.finish()
        for layer in model.layers:
            _lib.raise_if_not_pd(layer._info)
        elbo_val = float(model._elbo.item())
        if not np.isfinite(elbo_val):
            raise SystemExit("bench.py: non-finite ELBO (%r) after the timed steps" % elbo_val)
        return dict(ms=ms_res, ms_e2e=ms_e2e, launches=int(launches), clocks=clocks, model=model, layers=layers,
                    elbo=elbo_val, h2d=B * D_in * 4 + B * 4)

    S, B = cfg["S"], cfg["batch"]
    n_global = B * world
    main = measure(cfg, B, n_global, args.steps, args.warmup, e2e=True, sample_clocks=True)
    model, layers = main["model"], main["layers"]

    # every layer's K planes must carry non-negligible entries in the timed state (round-1 timed layers 2-3 on zeros)
    med0 = kuf_medians(build_model(layers, S, device), cfg, device)
    for li, m in enumerate(med0):
        if not (0.01 <= m <= 0.5):
            raise SystemExit("bench.py: degenerate synthetic state, layer %d median(Kuf)/sigma^2 = %.3g" % (li, m))

    # roofline of the dominant kernel (the tcgen05 conditional GEMM of layer 2), timed live with CUDA events
    roof = kernel_roofline(model, cfg, B, S, device, flush) if rank == 0 else None

    # strong scaling (BASELINE.json config 3/4 wording: the global minibatch is partitioned over the GPUs)
    strong = None
    if world > 1 and B % world == 0 and not args.no_extras:
        r = measure(cfg, B // world, B, args.steps, args.warmup, e2e=False)
        strong = {"global_batch": B, "images_per_gpu": B // world, "ms_per_step": r["ms"], "value": B / (r["ms"] * 1e-3),
                  "unit": "images/s", "gpu_launches": r["launches"]}
        del r

    # the other BASELINE.json configurations with a GPU leg, few steps each (parity-test cases; reported, not the headline)
    also = {}
    if not args.no_extras and args.config == "cfg3":
        for name in ("cfg2", "cfg4"):
            c = CONFIGS[name]
            r = measure(c, c["batch"], c["batch"] * world, max(5, args.steps // 2), 3, e2e=False)
            also[name] = {"workload": name + ": " + c["desc"], "images_per_gpu": c["batch"], "global_batch": c["batch"] * world,
                          "ms_per_step": r["ms"], "value": c["batch"] * world / (r["ms"] * 1e-3), "unit": "images/s",
                          "gpu_launches": r["launches"], "elbo": r["elbo"]}
            del r
            torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu, parity = None, None
    if world == 1 and not args.no_cpu_baseline:
        cores = host_threads()
        keep = []
        r, dt, ref_elbo, ref_grads = cpu_reference_rate(cfg, layers, args.ref_images, forward_only=args.forward_only, keep=keep)
        cpu = {"value": r, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": "%d images x S=%d, full model, %s in %.1f s, float64 oracle port on %d host threads (reference TF/GPflow "
                         "path not installable)" % (args.ref_images, S, "forward ELBO" if args.forward_only else
                                                    "forward + backward (torch-CPU autograd)", dt, cores)}
        if not args.forward_only:
            parity = gpu_parity(cfg, layers, args.ref_images, device, ref_elbo, ref_grads, keep)
    total_flops, _ = algorithmic_flops(cfg, B)
    if not args.forward_only:
        total_flops *= 3.0      # SURVEY 8d convention: backward = 2x forward
    config = workload_config(args, cfg, layers, B, n_global, "weak")
    config.update({"l2_flush": "256 MiB buffer written between timed steps",
                   "algorithmic_gflop_per_step_per_gpu": total_flops / 1e9, "elbo": main["elbo"],
                   "kuf_median_rel_gpu": [round(m, 4) for m in med0]})
    line = {"metric": "ELBO-step images/sec", "value": n_global / (main["ms"] * 1e-3), "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": main["ms"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": _lib.precision_string(), "data": "synthetic",
            "config": config,
            "e2e": {"value": n_global / (main["ms_e2e"] * 1e-3), "unit": "images/s", "ms_per_step": main["ms_e2e"],
                    "h2d_bytes_per_step": main["h2d"], "d2h_bytes_per_step": 8},
            "gpu_launches": main["launches"], "clocks": main["clocks"], "roofline": roof, "cpu_baseline": cpu,
            "parity": parity, "strong_scaling": strong, "also": also or None}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def ncu_summary():
    """Per-kernel numbers that only a profiler can give (DRAM traffic, tensor-pipe activity) come from the committed
    capture summary profiles/ncu_summary.json (written by tools/ncu_summarize.py from an `ncu --set full` capture of
    tools/profile_layer2.py; it names the capture CSV and the commit) -- never typed into this file."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))
    except Exception:
        return {}


def kernel_roofline(model, cfg, B, S, device, flush):
    """Dominant kernel = the tcgen05 conditional GEMM of conv layer 2 (forward): two launches of the same kernel,
    `tc_kernel<MODE_AP,128>` (a = K Lm^-T) + `tc_kernel<MODE_COND,256>` (G_r = a C_r), timed together; the two big
    backward GEMMs of the same layer (`dk_gemm_kernel`, `xf_gemm_kernel`) and the Kuf kernel are reported next to it.
    Every kernel is timed live with CUDA events recorded by the library on the launching stream around that launch
    (dcgp_set_kernel_timing), L2 flushed before every repetition, on the layer's ACTUAL input (samples of layer 1).
      achieved = ALGORITHMIC flops per launch (T*(M^2 + R*M^2 + 2MR + 2M(R+1)), triangular count, no split)
                 / launch duration;   peak = MEASURED_PEAKS.json bf16 burst (the kernel is timed alone);
      executed_* = the tensor-pipe flops the launch really issues (split products x the k-blocks not skipped as zero), as
                   counted by the library itself (dcgp_kernel_tensor_flops);
      traffic / tensor_pipe_active_pct_ncu = from profiles/ncu_summary.json (null when no capture of this build exists)."""
    import torch
    from deepcgp_b200 import _lib
    from deepcgp_b200.grad import LayerBackward
    if len(model.layers) < 3:
        return None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("bf16_tflops", 1590.0))
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    ncu = ncu_summary()
    layer = model.layers[1]
    layer._run_pending()
    n_rows = S * B
    rng = np.random.RandomState(11)
    X0 = torch.as_tensor(rng.standard_normal((B, cfg["H"] * cfg["W"] * cfg["C"])).astype(np.float32), device=device)
    Fs, _, _ = model.propagate(X0, S=S)
    X = Fs[0].reshape(n_rows, -1).contiguous()               # the layer's actual input: samples of layer 1
    gm = torch.randn((n_rows, layer.num_outputs), device=device) * 1e-3
    gv = torch.randn((n_rows, layer.num_outputs), device=device) * 1e-3
    lb = LayerBackward(layer)
    layer.prepare()
    layer._hold = True
    _lib.lib.dcgp_set_kernel_timing(1)
    for _ in range(3):
        layer._conditional(X)
        lb.t_sized(X, 1, gm, gv, True)
    torch.cuda.synchronize()
    t = {0: [], 1: [], 2: [], 3: []}
    for _ in range(10):
        flush.zero_()
        layer._conditional(X)
        t[0].append(_lib.lib.dcgp_kernel_ms(0))
        t[1].append(_lib.lib.dcgp_kernel_ms(1))
        flush.zero_()
        lb.t_sized(X, 1, gm, gv, True)
        t[2].append(_lib.lib.dcgp_kernel_ms(2))
        t[3].append(_lib.lib.dcgp_kernel_ms(3))
    exe = [float(_lib.lib.dcgp_kernel_tensor_flops(i)) for i in range(4)]
    _lib.lib.dcgp_set_kernel_timing(0)
    layer._hold = False
    ms, ms_kuf, ms_dk, ms_dq = (float(np.mean(t[i])) for i in range(4))
    M, R, P, L = layer.num_inducing, layer.gp_count, layer.patch_count, layer.patch_length
    T = P * n_rows
    alg = T * (M * M + R * M * M + 2.0 * M * R + 2.0 * M * (R + 1))
    ach = alg / (ms * 1e-3) / 1e12
    kuf_bytes = 4.0 * T * M + 4.0 * n_rows * X.shape[1]      # K planes written (hi+lo fp16) + images read
    src = "MEASURED_PEAKS.json bf16_tflops (burst)" if peaks else "fallback 1590"

    def prof(key, also=None):
        """DRAM bytes (summed over the launches a timed entry spans) and tensor-pipe activity of the dominant launch."""
        e, e2 = ncu.get(key, {}), ncu.get(also, {}) if also else {}
        traffic = e.get("dram_bytes")
        if traffic is not None and e2.get("dram_bytes") is not None:
            traffic += e2["dram_bytes"]
        return {"traffic": traffic, "tensor_pipe_active_pct_ncu": e.get("tensor_pipe_active_pct"),
                "ncu_source": e.get("source") or ncu.get("source")}      # (an entry re-captured later names its own file)

    def tensor_entry(name, key, t_ms, alg_flops, exe_flops, also=None):
        e = {"kernel": name, "bound": "tensor", "ms": t_ms, "achieved": alg_flops / (t_ms * 1e-3) / 1e12, "peak": peak,
             "unit": "TFLOP/s", "frac": alg_flops / (t_ms * 1e-3) / 1e12 / peak, "algorithmic_gflop": alg_flops / 1e9,
             "executed_tensor_gflop": exe_flops / 1e9, "executed_frac": exe_flops / (t_ms * 1e-3) / 1e12 / peak}
        e.update(prof(key, also))
        return e

    # backward GEMMs: algorithmic = the Q-form contraction 2*T*R*M^2 (dQ: its symmetric half)
    # (dK: da = sum_r s_r a SP_r, then dK = da Lm^-1 with the fused dd epilogue -- two launches timed together)
    dk = tensor_entry("dk_gemm_kernel<256,EPI_PLANES> + dk_gemm_kernel<256,EPI_DD> (da GEMM, then dK = da Lm^-1 + fused dd epilogue; "
                      "conv layer 2 backward)", "dk_gemm", ms_dk, 2.0 * T * R * M * M + 1.0 * T * M * M, exe[2], also="dk_gemm_stage2")
    dq = tensor_entry("xf_gemm_kernel<256> (dS_r = a^T diag(s_r) a, conv layer 2 backward)", "dq_gemm", ms_dq, 1.0 * T * R * M * M, exe[3])
    out = {"bound": "tensor", "kernel": "tc_kernel<MODE_AP,128> + tc_kernel<MODE_COND,256> (chained conditional GEMM, conv layer 2 forward)",
           "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "ms": ms, "algorithmic_gflop": alg / 1e9,
           "executed_tensor_gflop": exe[0] / 1e9, "executed_tflops": exe[0] / (ms * 1e-3) / 1e12,
           "executed_frac": exe[0] / (ms * 1e-3) / 1e12 / peak, "peak_source": src,
           "input": "samples of conv layer 1 (the layer's actual input in the benchmark state)"}
    out.update(prof("cond_gemm", "cond_gemm_stage1"))
    kuf = {"kernel": "kuf_tc_kernel<256,PAIR> (conv layer 2)", "bound": "hbm", "ms": ms_kuf,
           "achieved": kuf_bytes / (ms_kuf * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
           "frac": kuf_bytes / (ms_kuf * 1e-3) / 1e9 / hbm, "algorithmic_bytes": kuf_bytes,
           "executed_tensor_gflop": exe[1] / 1e9, "tensor_floor_ms": exe[1] / (peak * 1e12) * 1e3}
    kuf.update(prof("kuf"))
    out.update({"kuf": kuf, "dk_gemm": dk, "dq_gemm": dq, "cholesky": cholesky_times(device)})
    return out


def cholesky_times(device):
    """K-C is latency-bound at M <= 1024: reported as time versus M, not as a tensor-roofline fraction."""
    import torch
    from deepcgp_b200 import _lib
    out = {}
    rng = np.random.RandomState(0)
    for M in (256, 512, 1024):
        A = rng.standard_normal((M, M + 8))
        K0 = torch.as_tensor(A @ A.T + 0.5 * np.eye(M), device=device).contiguous()
        ws = torch.empty(_lib.lib.dcgp_cholesky_workspace_bytes(M), dtype=torch.uint8, device=device)
        info = torch.zeros(1, dtype=torch.int32, device=device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for i in range(6):
            K = K0.clone()
            e0.record()
            _lib.check(_lib.lib.dcgp_cholesky(_lib.ptr(K), M, _lib.ptr(ws), ws.numel(), _lib.ptr(info), _lib.stream()))
            e1.record()
            torch.cuda.synchronize()
            if i:
                ts.append(e0.elapsed_time(e1))
        out["M%d_ms" % M] = float(np.median(ts))
    out["note"] = "float64 right-looking Cholesky (potrf_f64), latency-bound: M^3/3 = 45 MFLOP at M=512"
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS))
    ap.add_argument("--ref-images", type=int, default=32, help="images in the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the strong-scaling leg and the cfg2 / cfg4 lines")
    ap.add_argument("--forward-only", action="store_true", help="time the forward ELBO alone (diagnostic)")
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--sequential", action="store_true", help="ElboGradient + Adam.step without the per-layer pipelining (diagnostic)")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
