#!/bin/bash
# multi-GPU session: tools/gpu_mgpu.sh <tag> <ngpus>
tag=${1:-mg}; n=${2:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded.py > gpurun_out/${tag}_sharded.log 2>&1
echo "check_sharded exit $?"; grep -E "rank|Error|error" gpurun_out/${tag}_sharded.log | head -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/${tag}_bench${n}.json 2> gpurun_out/${tag}_bench${n}.err
echo "bench exit $?"; tail -c 400 gpurun_out/${tag}_bench${n}.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench${n}.json").read().strip().splitlines()[-1])
print("weak: %.3f ms/step, %.0f img/s | strong:" % (d["ms_per_step"], d["value"]), d.get("strong_scaling"), "| also:", {k: (round(v["ms_per_step"], 3), round(v["value"])) for k, v in (d.get("also") or {}).items()})
PY
