#!/bin/bash
# A/B of environment switches on the GPU box: ENVS="A=1 B=2|C=3" bash tools/sweep_env.sh  ('|' separates runs; the first run is plain)
IFS='|' read -ra RUNS <<< "|${ENVS}"
for envs in "${RUNS[@]}"; do
  env $envs python bench.py --spp ${SPP:-256} --steps 2 --warmup 2 --no-cpu-baseline --no-f32-leg 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); st=d['stage_ms_per_step']
print('[$envs]', round(d['value'],1), 'Mrays/s', ' '.join(k+'='+str(round(v,1)) for k,v in st.items() if k in ('extend','shade','shadow','generate')))"
done
