#!/bin/bash
# L2 warm-up: size x eviction policy on the headline workload; run under gpurun
run() { timeout 200 python bench.py --no-cpu-baseline --profile-steps 0 --steps 200 --warmup 10 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print(round(d['value']), round(d['e2e']['value']), round(d['ms_per_step']*1e3,1), 'us/step')"; }
for pol in 0 1; do for mb in 32 48 64 80 96 112; do echo "== policy=$pol l2_prefetch_mb=$mb"; run --opt l2_prefetch_mb=$mb --opt l2_prefetch_policy=$pol; done; done
echo "== default"; run
