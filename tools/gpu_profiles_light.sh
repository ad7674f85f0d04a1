#!/bin/bash
# launch list + step timeline only (no ncu --set full): tools/gpu_profiles_light.sh <tag>
tag=${1:-r2}
bash tools/gpu_launchlist.sh ${tag} | tail -45 > gpurun_out/${tag}_launch_summary.txt
timeout 300 python tools/trace_step.py ${tag} > gpurun_out/${tag}_trace.log 2>&1
python tools/analyze_trace.py gpurun_out/trace_${tag}_kernels.json.gz 150 > gpurun_out/${tag}_step_timeline.txt 2>&1
head -4 gpurun_out/${tag}_step_timeline.txt
