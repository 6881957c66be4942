#!/bin/bash
# refill / wait thresholds of the persistent traversal (GPU box)
for cfg in ${CONFIGS:-"8 8" "4 8" "12 8" "16 8" "8 4" "8 16" "8 33" "6 12"}; do
  set -- $cfg
  CRAY_REFILL_LANES=$1 CRAY_WAIT_LANES=$2 python bench.py --spp ${SPP:-256} --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); st=d['stage_ms_per_step']
print('refill=$1 wait=$2', round(d['value'],1), 'Mrays/s', ' '.join(k+'='+str(round(v,1)) for k,v in st.items() if k in ('extend','shadow')))"
done
