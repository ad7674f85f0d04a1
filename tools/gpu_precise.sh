#!/bin/bash
# precise stage 1 (four TMEM accumulators): accuracy on the ill-conditioned layer, cost at M = 512 / 1024, cfg4 gradients
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/diag_accum.py time 2>&1 | tail -12
for mode in 0 1; do
  echo "== cfg4 gradient test, DCGP_PRECISE_STAGE1=$mode"
  DCGP_PRECISE_STAGE1=$mode timeout 900 python -m pytest tests/test_gpu_bench_size.py -q -s -k "forward_elbo_gradients and cfg4" 2>&1 | grep -E "gradient normwise|passed|failed|Error" | cut -c1-1200
done
timeout 900 python -m pytest tests/test_gpu_parity.py -q 2>&1 | tail -3
