#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bench_size.py tests/test_gpu_backward_pieces.py -q -rP > gpurun_out/quick_pytest.log 2>&1
grep -E "products \(|gradient normwise|passed|failed" gpurun_out/quick_pytest.log | cut -c1-900
grep -E "^E  " gpurun_out/quick_pytest.log | cut -c1-300 | head
