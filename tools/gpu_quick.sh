#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_grad.py tests/test_gpu_backward_pieces.py tests/test_gpu_bench_size.py -q > gpurun_out/q3_pytest.log 2>&1
tail -3 gpurun_out/q3_pytest.log; grep -E "^E  " gpurun_out/q3_pytest.log | cut -c1-250 | head -12
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/q3_bench.json 2> gpurun_out/q3_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/q3_bench.json").read().strip().splitlines()[-1]); r = d["roofline"]
print("ms/step %.3f e2e %.3f | cond %.3f kuf %.3f dk %.3f dq %.3f (exe %.3f)" % (d["ms_per_step"], d["e2e"]["ms_per_step"], r["ms"], r["kuf"]["ms"], r["dk_gemm"]["ms"], r["dq_gemm"]["ms"], r["dq_gemm"]["executed_frac"]))
PY
