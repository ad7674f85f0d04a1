#!/bin/bash
# launch list of the bench command under ncu (per-launch durations; cold-cache, serialised): tools/gpu_launchlist.sh <tag>
tag=${1:-ll}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/${tag}_bench_under_ncu.log 2>&1
echo "ncu exit $?"
python tools/summarize_launches.py gpurun_out/${tag}_launches.csv 3 | head -60
