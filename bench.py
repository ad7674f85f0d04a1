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
def synth_params(cfg, seed=1236):
    """Seeded synthetic parameters: Z = patches sampled from N(0,1) inputs of the layer's shape + N(0,0.1^2);
    trained-like q: q_mu ~ N(0,1), q_sqrt = tril(N(0,0.3^2)) + 0.5 I.  Pure numpy, identical on every rank."""
    rng = np.random.RandomState(seed)
    layers = []
    h, w, c = cfg["H"], cfg["W"], cfg["C"]
    specs = [(f, s, M, R, "conv") for (f, s, M, R) in cfg["conv"]] + [cfg["last"] + (10, "svgp_conv")]
    for (f, s, M, R, kind) in specs:
        L = f * f * c
        oh, ow = (h - f) // s + 1, (w - f) // s + 1
        nimg = max(8, 2 * M // (oh * ow) + 1)
        img = rng.standard_normal((nimg, h, w, c))
        Z = np.empty((M, L))
        for i in range(M):
            n, y, x = rng.randint(nimg), rng.randint(h - f + 1), rng.randint(w - f + 1)
            Z[i] = img[n, y:y + f, x:x + f, :].reshape(-1)
        Z += 0.1 * rng.standard_normal((M, L))
        lay = dict(type=kind, H=h, W=w, C=c, f=f, s=s, M=M, R=R, white=False, variance=SIGMA2, lengthscale=LENGTHSCALE,
                   Z=Z, q_mu=rng.standard_normal((M, R)),
                   q_sqrt=np.tril(rng.standard_normal((R, M, M)) * 0.3) + 0.5 * np.eye(M))
        if kind == "svgp_conv":
            lay["patch_weights"] = np.ones(oh * ow)
        layers.append(lay)
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
def cpu_reference_rate(cfg, layers, n_img, seed=99, forward_only=False):
    """The reference's CPU implementation cannot run here (TensorFlow/GPflow absent).  Timed instead,
    on a bounded sample of the same workload (`n_img` images with all S samples each, full 3-layer model):
      * ELBO step  : the float64 torch-CPU restatement with autograd (oracle/dcgp_oracle_torch.py): forward + backward,
                     which is what TensorFlow executes per optimiser step (tf.gradients; Adam itself is negligible);
      * forward only: the float64 NumPy/SciPy oracle (oracle/dcgp_oracle.py, single-solve form).
    Both use all host BLAS threads."""
    from oracle import dcgp_oracle as O
    rng = np.random.RandomState(seed)
    S = cfg["S"]
    X = rng.standard_normal((n_img, cfg["H"] * cfg["W"] * cfg["C"]))
    Y = rng.randint(0, 10, size=(n_img, 1))
    zs = []
    for lay in layers:
        oh, ow = O.out_image_size(lay["H"], lay["W"], lay["f"], lay["s"])
        D = oh * ow * lay["R"] if lay["type"] == "conv" else lay["R"]
        zs.append(rng.standard_normal((S, n_img, D)))
    t0 = time.perf_counter()
    if forward_only:
        elbo = O.dgp_elbo(layers, X, Y, zs, NUM_DATA, S, JITTER, fast=True)
    else:
        from oracle import dcgp_oracle_torch as OT
        elbo, _ = OT.elbo_and_grads(layers, X, Y, zs, NUM_DATA, S, JITTER)
    dt = time.perf_counter() - t0
    return n_img / dt, dt, float(elbo)


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    layers = synth_params(cfg)
    n_img = args.ref_images
    rates = []
    for i in range(args.warmup + args.steps):
        r, dt, _ = cpu_reference_rate(cfg, layers, n_img, seed=99 + i, forward_only=args.forward_only)
        if i >= args.warmup:
            rates.append((r, dt))
    value = float(np.mean([r for r, _ in rates]))
    ms = float(np.mean([dt for _, dt in rates])) * 1e3
    cores = os.cpu_count()
    what = "forward ELBO (NumPy/SciPy float64)" if args.forward_only else "forward + backward (torch-CPU float64 autograd)"
    sample = "%d images x S=%d, full 3-layer model, %s; oracle port, TF/GPflow not installable" % (n_img, cfg["S"], what)
    line = {"impl": "reference", "metric": "ELBO-step images/sec", "value": value, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.config + ": " + cfg["desc"], "step": what, "sample": sample},
            "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------- our arm
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

    S, B = cfg["S"], cfg["batch"]
    n_global = B * world
    layers = synth_params(cfg)
    model = build_model(layers, S, device)
    rng = np.random.RandomState(4321 + rank)
    D_in = cfg["H"] * cfg["W"] * cfg["C"]
    n_batches = 4                                        # rotate distinct host batches (pinned), like a data loader
    hostX = [torch.from_numpy(rng.standard_normal((B, D_in)).astype(np.float32)).pin_memory() for _ in range(n_batches)]
    hostY = [torch.from_numpy(rng.randint(0, 10, size=(B,)).astype(np.int32)).pin_memory() for _ in range(n_batches)]
    devX = [x.to(device) for x in hostX]
    devY = [y.to(device) for y in hostY]
    # The N(0,1) draws of DS/layers.py:104 are part of the step: drawn inside the timed region, every step, by the
    # counter-based generator (indexed by the global image position, so every rank count sees the same noise).
    def draw():
        return model.draw_zs(B, n_global, rank * B)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=device)   # > 126 MB L2
    elbo_host = torch.empty(1, dtype=torch.float64).pin_memory()

    import deepcgp_b200 as D
    if args.sequential:
        eg = D.ElboGradient(model)
        opt = D.Adam(model, lr=args.lr)
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

    def timed(step_fn, steps, warmup):
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

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.lib.dcgp_launch_count()
    ms_res = timed(step_resident, args.steps, args.warmup)
    launches = _lib.lib.dcgp_launch_count() - launches0
    ms_e2e = timed(step_e2e, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    if not args.forward_only and not args.sequential:
        train_step.finish()
    for layer in model.layers:
        _lib.raise_if_not_pd(layer._info)
    elbo_val = float(model._elbo.item())

    # roofline of the dominant kernel (the tcgen05 conditional GEMM of layer 2), timed live with CUDA events
    roof = None
    if rank == 0:
        roof = kernel_roofline(model, cfg, B, S, device, flush)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r, dt, _ = cpu_reference_rate(cfg, layers, args.ref_images, forward_only=args.forward_only)
        cpu = {"value": r, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
               "sample": "%d images x S=%d, full 3-layer model, %s in %.1f s, float64 oracle port (reference TF/GPflow "
                         "path not installable)" % (args.ref_images, S, "forward ELBO" if args.forward_only else
                                                    "forward + backward (torch-CPU autograd)", dt)}
    total_flops, _ = algorithmic_flops(cfg, B)
    if not args.forward_only:
        total_flops *= 3.0      # SURVEY 8d convention: backward = 2x forward
    line = {"metric": "ELBO-step images/sec", "value": n_global / (ms_res * 1e-3), "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_res, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp16x2-split tensor cores (fp32 accumulate) + f64 M-only", "data": "synthetic",
            "config": {"workload": args.config + ": " + cfg["desc"], "step": "forward ELBO only" if args.forward_only else
                       "forward ELBO + backward + gradient all-reduce + Adam (all trainables)",
                       "images_per_gpu": B, "num_samples": S, "l2_flush": "256 MiB buffer written between timed steps",
                       "algorithmic_gflop_per_step_per_gpu": total_flops / 1e9, "elbo": elbo_val},
            "e2e": {"value": n_global / (ms_e2e * 1e-3), "unit": "images/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": B * D_in * 4 + B * 4, "d2h_bytes_per_step": 8},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def kernel_roofline(model, cfg, B, S, device, flush):
    """Dominant kernel = the tcgen05 conditional GEMM of conv layer 2 (forward) -- since the chained form two launches of
    the same kernel, `tc_kernel<MODE_A,256>` (a = K Lm^-T) + `tc_kernel<MODE_COND,256>` (G_r = a C_r), timed together; the two big
    backward GEMMs of the same layer (`dk_gemm_kernel`, `xf_gemm_kernel`) and the Kuf kernel are reported next to it.
    Every kernel is timed live with CUDA events recorded by the library on the launching stream around that launch
    (dcgp_set_kernel_timing), L2 flushed before every repetition.
      achieved = ALGORITHMIC flops per launch (T*(M^2 + R*M^2 + 2MR + 2M(R+1)), triangular count, no split)
                 / launch duration;   peak = MEASURED_PEAKS.json bf16 burst (the kernel is timed alone);
      executed_* = the tensor-pipe flops the launch really issues: 3 split products x the k-blocks not skipped as zero;
      traffic = dram__bytes_read + dram__bytes_write of this launch from the committed `ncu --set full` capture
                (profiles/r1c_ncu_full_layer2.csv; dense form: profiles/r1b_ncu_full_layer2.csv)."""
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
    layer = model.layers[1]
    layer._run_pending()
    n_rows = S * B
    D_in = int(np.prod(layer.view.input_size)) * layer.view.feature_maps
    X = torch.randn((n_rows, D_in), device=device)
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
    _lib.lib.dcgp_set_kernel_timing(0)
    layer._hold = False
    ms, ms_kuf, ms_dk, ms_dq = (float(np.mean(t[i])) for i in range(4))
    M, R, P, L = layer.num_inducing, layer.gp_count, layer.patch_count, layer.patch_length
    T = P * n_rows
    alg = T * (M * M + R * M * M + 2.0 * M * R + 2.0 * M * (R + 1))
    chained = os.environ.get("DCGP_FWD_CHAINED", "1") != "0"     # two triangular stages: 3/4 of the k-blocks of each tile
    executed = 3 * 2.0 * T * ((0.75 if chained else 1.0) * (R + 1) * M * M + 256 * M)
    ach = alg / (ms * 1e-3) / 1e12
    kuf_bytes = 4.0 * T * M + 4.0 * n_rows * D_in            # K planes written (hi+lo fp16) + images read
    kuf_flops = 3 * 2.0 * T * M * 256                        # 3 split products over the padded patch length
    src = "MEASURED_PEAKS.json bf16_tflops (burst)" if peaks else "fallback 1590"

    def tensor_entry(name, t_ms, alg_flops, exe_flops, extra):
        e = {"kernel": name, "bound": "tensor", "ms": t_ms, "achieved": alg_flops / (t_ms * 1e-3) / 1e12, "peak": peak,
             "unit": "TFLOP/s", "frac": alg_flops / (t_ms * 1e-3) / 1e12 / peak, "algorithmic_gflop": alg_flops / 1e9,
             "executed_tensor_gflop": exe_flops / 1e9, "executed_frac": exe_flops / (t_ms * 1e-3) / 1e12 / peak}
        e.update(extra)
        return e

    # backward GEMMs: algorithmic = the Q-form contraction 2*T*R*M^2 (dQ: its symmetric half); executed = 3 split products
    # (dK: + the 64-deep mean tile; dQ: 6 of 8 output tiles per r)
    dk = tensor_entry("dk_gemm_kernel<256,true> (dK GEMM + fused dd epilogue, conv layer 2 backward)", ms_dk,
                      2.0 * T * R * M * M, 3 * 2.0 * T * (R * M * M + 64 * M),
                      {"traffic": 1903177696, "tensor_pipe_active_pct_ncu": 77.6,
                       "epilogue_bytes": 2.0 * 4 * T * M + 4.0 * T * M})
    dq = tensor_entry("xf_gemm_kernel<256> (dQ GEMM, in-smem column rescale, conv layer 2 backward)", ms_dq,
                      1.0 * T * R * M * M, 3 * 2.0 * T * R * M * M * 0.75,
                      {"traffic": 790520072, "tensor_pipe_active_pct_ncu": 60.5})
    return {"bound": "tensor", "kernel": "tc_kernel<MODE_A,256> + tc_kernel<MODE_COND,256> (chained conditional GEMM, conv layer 2 "
                                         "forward)" if chained else "tc_kernel<MODE_COND,256> (conditional GEMM, conv layer 2 forward)",
            "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "traffic": 1590374000 if chained else 877827921, "ms": ms, "algorithmic_gflop": alg / 1e9,
            "executed_tensor_gflop": executed / 1e9, "executed_tflops": executed / (ms * 1e-3) / 1e12,
            "executed_frac": executed / (ms * 1e-3) / 1e12 / peak,
            "peak_source": src, "tensor_pipe_active_pct_ncu": 93.2 if chained else 96.9,
            "kuf": {"kernel": "kuf_tc_kernel<256> (conv layer 2)", "bound": "hbm", "ms": ms_kuf,
                    "achieved": kuf_bytes / (ms_kuf * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                    "frac": kuf_bytes / (ms_kuf * 1e-3) / 1e9 / hbm, "algorithmic_bytes": kuf_bytes,
                    "traffic": 498662144, "executed_tensor_gflop": kuf_flops / 1e9,
                    "tensor_floor_ms": kuf_flops / (peak * 1e12) * 1e3},
            "dk_gemm": dk, "dq_gemm": dq}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS))
    ap.add_argument("--ref-images", type=int, default=32, help="images in the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
