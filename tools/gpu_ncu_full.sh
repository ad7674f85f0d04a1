#!/bin/bash
# ncu --set full capture of conv layer 2's kernels (one iteration after one warm-up iteration): tools/gpu_ncu_full.sh <tag>
tag=${1:-r2}
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on \
    -k regex:"^tc_kernel|kuf_tc_kernel|dk_gemm_kernel|xf_gemm_kernel" -c 48 -o gpurun_out/${tag}_full -f \
    python tools/profile_layer2.py 2 > gpurun_out/${tag}_ncu_full.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/${tag}_ncu_full.log; ls -la gpurun_out/${tag}_full.ncu-rep
ncu -i gpurun_out/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_full_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}_full.ncu-rep --page details --csv 2>/dev/null | grep -E "dk_gemm_kernel<256, 1>|xf_gemm_kernel<256>|tc_kernel<0, 256>" | grep -E "Stall|Bank|Shared|L2|DRAM Throughput|Issue" | head -60 > gpurun_out/${tag}_details_excerpt.csv
python tools/ncu_summarize.py gpurun_out/${tag}_full_raw.csv gpurun_out/${tag}_ncu_full_layer2.csv "ncu --set full -k regex:^tc_kernel|kuf_tc_kernel|dk_gemm_kernel|xf_gemm_kernel -c 48 python tools/profile_layer2.py 2" | head -60
rm -f gpurun_out/${tag}_full.ncu-rep gpurun_out/${tag}_full_raw.csv     # (gpurun copies back at most 64 MiB)
