"""GPU dev tool: Kineto trace of a few pipelined ELBO steps -> gpurun_out/trace_<tag>.json.gz + a per-stream summary."""
import gzip, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, deepcgp_b200 as D
from torch.profiler import profile, ProfilerActivity
tag = sys.argv[1] if len(sys.argv) > 1 else "step"
seq = len(sys.argv) > 2 and sys.argv[2] == "seq"
cfg = bench.CONFIGS[os.environ.get("DCGP_TRACE_CFG", "cfg3")]; dev = torch.device("cuda:0")
layers = bench.synth_params(cfg); model = bench.build_model(layers, cfg["S"], dev)
S, B = cfg["S"], cfg["batch"]
X = torch.randn((B, 3072), device=dev); Y = torch.randint(0, 10, (B,), device=dev, dtype=torch.int32)
zs = [torch.randn((S, B, l.num_outputs), device=dev) for l in model.layers]
if seq:
    eg, opt = D.ElboGradient(model), D.Adam(model, lr=1e-3)
    def step():
        e, g = eg(X, Y, zs=zs); opt.step(g)
else:
    ts = D.TrainStep(model, lr=1e-3)
    def step():
        ts(X, Y, zs=zs)
for _ in range(4):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)
path = "gpurun_out/trace_%s.json" % tag
prof.export_chrome_trace(path)
ev = json.load(open(path))["traceEvents"]
k = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
slim = [{"name": e["name"][:80], "ts": e["ts"], "dur": e["dur"], "stream": e["args"].get("stream"), "cat": e["cat"]} for e in k]
json.dump(slim, gzip.open("gpurun_out/trace_%s_kernels.json.gz" % tag, "wt"))
os.remove(path)
print("kernels", len(slim))
