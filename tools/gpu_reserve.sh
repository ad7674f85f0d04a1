#!/bin/bash
mkdir -p gpurun_out
for k in 0 8 16 32; do
  DCGP_RESERVE_SMS=$k timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/rs_$k.json 2> gpurun_out/rs_$k.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/rs_$k.json").read().strip().splitlines()[-1])
print("reserve $k: ms/step %.3f e2e %.3f" % (d["ms_per_step"], d["e2e"]["ms_per_step"]), d["clocks"])
PY
done
