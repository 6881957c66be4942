#!/bin/bash
# sweep of library variants (make VARIANT=...) and traversal thresholds on the GPU box
run() {  # label, env...
  label=$1; shift
  env "$@" python bench.py --spp 128 --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$label', round(d['value'],1), 'Mrays/s  extend_ms', round(d['roofline']['kernel_ms'],1), 'step_ms', round(d['ms_per_step'],1), 'share', round(d['roofline']['kernel_share_of_step'],3))"
}
for v in "" $VARIANTS; do
  run "lib=$v" CRAY_B200_LIB=$PWD/craytracer_b200/libcray_b200$v.so
done
