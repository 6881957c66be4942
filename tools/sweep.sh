#!/bin/bash
# sweep of library variants (make VARIANT=...) on the GPU box:  VARIANTS="_a _b" SPP=256 bash tools/sweep.sh
run() {  # label, env...
  label=$1; shift
  env "$@" python bench.py --spp ${SPP:-256} --steps 2 --warmup 2 --no-cpu-baseline --no-f32-leg 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
st=d['stage_ms_per_step']
print('$label', round(d['value'],1), 'Mrays/s  step_ms', round(d['ms_per_step'],1), ' '.join(k+'='+str(round(v,1)) for k,v in st.items()))"
}
for v in "" $VARIANTS; do
  run "lib=$v" CRAY_B200_LIB=$PWD/craytracer_b200/libcray_b200$v.so
done
