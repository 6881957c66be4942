#!/bin/bash
# parameter sweep of the persistent traversal (GPU box): library variants x refill / wait thresholds
run() {  # label, env...
  label=$1; shift
  env "$@" python bench.py --spp 128 --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$label', round(d['value'],1), 'Mrays/s  extend_ms', round(d['roofline']['kernel_ms'],1), 'step_ms', round(d['ms_per_step'],1), 'share', round(d['roofline']['kernel_share_of_step'],3))"
}
for v in "" _mb5 _mb4 _mb8; do
  run "lib=$v refill=8 wait=8" CRAY_B200_LIB=$PWD/craytracer_b200/libcray_b200$v.so
done
for cfg in "4 8" "16 8" "8 4" "8 16" "8 33" "12 12"; do
  set -- $cfg
  run "lib=default refill=$1 wait=$2" CRAY_REFILL_LANES=$1 CRAY_WAIT_LANES=$2
done
