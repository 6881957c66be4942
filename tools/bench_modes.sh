#!/bin/bash
# one short bench line per traversal mode on the GPU box:  MODES="fast f32" SPP=256 bash tools/bench_modes.sh
for m in ${MODES:-fast f32}; do
  python bench.py --spp ${SPP:-256} --steps 2 --warmup 2 --no-cpu-baseline --no-f32-leg --mode $m 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
st=d['stage_ms_per_step']
print('mode=$m', round(d['value'],1), 'Mrays/s  step_ms', round(d['ms_per_step'],1), ' '.join(k+'='+str(round(v,1)) for k,v in st.items()))"
done
