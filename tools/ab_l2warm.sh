#!/bin/bash
# A/B of the L2 warm-up of the next block's C (xl_set_option l2_prefetch_mb); run under gpurun
for mb in 0 16 32 48 64 96 160; do
  echo "== l2_prefetch_mb=$mb"
  python bench.py --steps 200 --warmup 10 --no-cpu-baseline --profile-steps 5 --opt l2_prefetch_mb=$mb "$@" | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']
print(round(d['value']), round(d['e2e']['value']), round(d['ms_per_step']*1e3,1), 'us/step; state avg', round(r['avg_launch_us'],1), 'us share', round(r['share_of_step'],3))"
done
