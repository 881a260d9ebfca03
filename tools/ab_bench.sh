#!/bin/bash
# Same-box A/B of one environment switch on the 8-view bench step (run through gpurun):
#   tools/ab_bench.sh MA_GEMM_STREAMK 0 1 [extra bench.py args]
# alternates the two settings twice (A B A B) and prints value / ms / e2e / GEMM and attention ms per step for each run.
VAR=$1; A=$2; B=$3; shift 3
for v in $A $B $A $B; do
  env $VAR=$v python bench.py --no-eager --no-cpu --steps 10 --warmup 3 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
f=d['roofline']['families_ms']
print('$VAR=$v', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'gemm', f.get('gemm'), 'attn', f.get('attention'), 'gemm_frac', round(d['roofline']['frac'],3), 'step_frac', round(d['roofline']['step_frac'],3))"
done
