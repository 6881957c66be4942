#!/bin/bash
# pool-size sweep on the GPU box
for p in ${POOLS:-21 22 23 24}; do
CRAY_POOL_LOG2=$p python bench.py --spp ${SPP:-256} --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); st=d['stage_ms_per_step']
print('pool=2^$p', round(d['value'],1), 'Mrays/s  step_ms', round(d['ms_per_step'],1), ' '.join(k+'='+str(round(v,1)) for k,v in st.items()))"
done
