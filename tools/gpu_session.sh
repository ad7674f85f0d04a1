#!/bin/bash
# One gpurun call: GPU test-suite, then the bench with the three split-product settings.  Usage: tools/gpu_session.sh <tag>
tag=${1:-s}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rP > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
for pr in 333 311 113 111; do
  c=${pr:0:1}; k=${pr:1:1}; q=${pr:2:1}
  DCGP_PROD_COND=$c DCGP_PROD_DK=$k DCGP_PROD_DQ=$q timeout 600 python bench.py --steps 10 --warmup 3 --no-extras \
      > gpurun_out/${tag}_bench_${pr}.json 2> gpurun_out/${tag}_bench_${pr}.err
  echo "bench $pr exit $?"; tail -c 600 gpurun_out/${tag}_bench_${pr}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_${pr}.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("$pr", "ms/step %.3f e2e %.3f" % (d["ms_per_step"], d["e2e"]["ms_per_step"]), "cond %.3f ms frac %.3f exe %.3f | kuf %.3f | dk %.3f exe %.3f | dq %.3f exe %.3f" % (
        r["ms"], r["frac"], r["executed_frac"], r["kuf"]["ms"], r["dk_gemm"]["ms"], r["dk_gemm"]["executed_frac"], r["dq_gemm"]["ms"], r["dq_gemm"]["executed_frac"]))
    print("parity", json.dumps(d["parity"]))
    print("clocks", d["clocks"], "launches", d["gpu_launches"], "kuf medians", d["config"]["kuf_median_rel_gpu"])
except Exception as e:
    print("$pr parse failed", e)
PY
done
