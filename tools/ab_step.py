"""A/B timing of TrainStep settings inside ONE process (run-to-run variance between gpurun boxes is +-0.5 ms):
alternates the settings in blocks of `--block` steps, `--rounds` times, L2 flushed between steps.
    python tools/ab_step.py reserve 0 16 32 -1      (DCGP_RESERVE_SMS values; -1 = one CTA per item for the deferred GEMMs)
    python tools/ab_step.py precise 0 1             (dcgp_set_precise_stage1)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, deepcgp_b200 as D
what, vals = sys.argv[1], [int(v) for v in sys.argv[2:]]
cfg = bench.CONFIGS["cfg3"]; dev = torch.device("cuda:0")
layers = bench.synth_params(cfg); model = bench.build_model(layers, cfg["S"], dev)
B = cfg["batch"]
rng = np.random.RandomState(0)
X = torch.as_tensor(rng.standard_normal((B, 3072)).astype(np.float32), device=dev)
Y = torch.as_tensor(rng.randint(0, 10, size=(B,)).astype(np.int32), device=dev)
ts = D.TrainStep(model, lr=1e-3)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
def step():
    return ts(X, Y, zs=model.draw_zs(B, B, 0), n_global=B)
for _ in range(8):
    step()
torch.cuda.synchronize()
res = {v: [] for v in vals}
for rnd in range(4):
    for v in vals:
        if what == "reserve":
            ts.reserve_sms = v
        elif what == "prepare":          # TrainStep.early_prepare: next-step prepare queued behind the Adam slice (device-side hyp)
            ts.finish(); ts.early_prepare = bool(v)
        elif what == "precise":          # dcgp_set_precise_stage1: stage 1 of the conditional on four TMEM accumulators
            D._lib.lib.dcgp_set_precise_stage1(v)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        ev = []
        for _ in range(25):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); step(); b.record(); ev.append((a, b))
        torch.cuda.synchronize()
        res[v].append(float(np.mean([a.elapsed_time(b) for a, b in ev])))
for v in vals:
    print(what, v, "ms/step per round:", ["%.3f" % x for x in res[v]], "mean %.3f" % np.mean(res[v]))
