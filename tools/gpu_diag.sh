#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_backward_pieces.py -q -rP > gpurun_out/diag_pieces.log 2>&1
grep -E "normwise|passed|failed|Error|assert" gpurun_out/diag_pieces.log | head -40
