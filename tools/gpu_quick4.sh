#!/bin/bash
mkdir -p gpurun_out
python tools/ab_step.py reserve 0 -1 2>&1 | tail -3
timeout 300 python tools/trace_step.py q4 > gpurun_out/q4_trace.log 2>&1; tail -1 gpurun_out/q4_trace.log
python tools/analyze_trace.py gpurun_out/trace_q4_kernels.json.gz 300 2>&1 | head -40
