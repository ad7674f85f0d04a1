#!/bin/bash
# ncu --set full of the dS kernel only (conv layer 2 of cfg3): tools/gpu_ncu_xf.sh <tag>
tag=${1:-xf}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:"xf_gemm_kernel" -s 1 -c 1 -o gpurun_out/${tag}_xf -f python tools/profile_layer2.py 2 > gpurun_out/${tag}_ncu_xf.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/${tag}_xf.ncu-rep --page raw --csv 2>/dev/null > gpurun_out/${tag}_xf_raw.csv
python - <<PY
import csv
rows = list(csv.reader(open("gpurun_out/${tag}_xf_raw.csv", errors="replace")))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
head, body = rows[hi], rows[hi + 2:]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "sm__inst_issued.avg.per_cycle_active", "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in body:
    print({w: r[head.index(w)] for w in want if w in head})
PY
rm -f gpurun_out/${tag}_xf.ncu-rep
