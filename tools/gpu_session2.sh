#!/bin/bash
tag=${1:-s}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/${tag}_pytest.log
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $?"; tail -c 300 gpurun_out/${tag}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("ms/step %.3f (%.0f img/s) e2e %.3f" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"]), "launches/step", d["gpu_launches"] / d["steps"])
print("cond %.3f ms frac %.3f exe %.3f | kuf %.3f frac %.3f | dk %.3f frac %.3f exe %.3f | dq %.3f frac %.3f exe %.3f" % (
    r["ms"], r["frac"], r["executed_frac"], r["kuf"]["ms"], r["kuf"]["frac"], r["dk_gemm"]["ms"], r["dk_gemm"]["frac"], r["dk_gemm"]["executed_frac"],
    r["dq_gemm"]["ms"], r["dq_gemm"]["frac"], r["dq_gemm"]["executed_frac"]))
print("chol", r["cholesky"])
p = d["parity"]; print("parity elbo %.1e mean %s var %s grad %.1e" % (p["elbo_rel"], ["%.1e" % x for x in p["mean_rel"]], ["%.1e" % x for x in p["var_rel"]], p["grad_rel"]))
print("clocks", d["clocks"], "cpu", d["cpu_baseline"]["value"], "also", {k: round(v["ms_per_step"], 3) for k, v in (d["also"] or {}).items()})
PY
timeout 300 python tools/trace_step.py ${tag} > gpurun_out/${tag}_trace.log 2>&1; tail -2 gpurun_out/${tag}_trace.log
python tools/analyze_trace.py gpurun_out/trace_${tag}_kernels.json.gz 2>&1 | tail -60
