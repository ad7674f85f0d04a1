#!/bin/bash
# Same-box A/B of builds of libdcgp.so: deepcgp_b200/lib/libdcgp.so ("new") against every deepcgp_b200/lib_ab/libdcgp_<name>.so.
# (the box's copy of the tree is scratch: swapping the file in place is safe there)   usage: gpu_ab_lib.sh [bench|kuf] [rounds]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cp deepcgp_b200/lib/libdcgp.so /tmp/new.so
one() {
  if [ "$2" = "kuf" ]; then echo "== $1"; python tools/diag_kuf.py 2>&1 | tail -4; return; fi
  python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('%-8s ms/step %.3f e2e %.3f | cond %.3f kuf %.3f dk %.3f dq %.3f | %s MHz' % ('$1', d['ms_per_step'], d['e2e']['ms_per_step'], r['ms'], r['kuf']['ms'], r['dk_gemm']['ms'], r['dq_gemm']['ms'], d['clocks']['sm_mhz']))"
}
for rnd in $(seq 1 ${2:-2}); do
  cp /tmp/new.so deepcgp_b200/lib/libdcgp.so; one new $1
  for f in deepcgp_b200/lib_ab/libdcgp_*.so; do
    n=$(basename $f .so); n=${n#libdcgp_}
    cp $f deepcgp_b200/lib/libdcgp.so; one $n $1
  done
done
cp /tmp/new.so deepcgp_b200/lib/libdcgp.so
