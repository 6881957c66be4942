#!/bin/bash
# parameter sweep of the persistent traversal (GPU box)
for cfg in "8 10" "4 10" "16 10" "8 4" "8 16" "8 24" "12 16" "16 20" "1 1"; do
  set -- $cfg
  CRAY_REFILL_LANES=$1 CRAY_PRIM_LANES=$2 python bench.py --spp 128 --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('refill=$1 prim=$2', round(d['value'],1), 'Mrays/s extend_ms', round(d['roofline']['kernel_ms'],1), 'step_ms', round(d['ms_per_step'],1))"
done
