#!/bin/bash
# Evidence for profiles/: ncu --set full capture of conv layer 2, launch list of the bench command, step timeline.  tools/gpu_profiles.sh <tag>
tag=${1:-r2}
bash tools/gpu_ncu_full.sh ${tag}
bash tools/gpu_launchlist.sh ${tag} | tail -45 > gpurun_out/${tag}_launch_summary.txt
timeout 300 python tools/trace_step.py ${tag} > gpurun_out/${tag}_trace.log 2>&1
python tools/analyze_trace.py gpurun_out/trace_${tag}_kernels.json.gz 150 > gpurun_out/${tag}_step_timeline.txt 2>&1
ls -la gpurun_out | head -30; du -sh gpurun_out
