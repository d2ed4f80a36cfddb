#!/bin/bash
# finer sweep of l2_prefetch_mb on the headline workload + the two big configs; run under gpurun
run() { timeout 200 python bench.py --no-cpu-baseline --profile-steps 0 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print(round(d['value']), round(d['e2e']['value']), round(d['ms_per_step']*1e3,1), 'us/step')"; }
for mb in 0 24 32 40 48 56 64 72 0; do echo "== 48M x 64 l2_prefetch_mb=$mb"; run --steps 200 --warmup 10 --opt l2_prefetch_mb=$mb; done
for mb in 0 48 96 0; do echo "== 206M x 128 l2_prefetch_mb=$mb"; run --model 206M --envs 128 --domains mixed --steps 30 --warmup 3 --opt l2_prefetch_mb=$mb; done
for mb in 0 48 0; do echo "== 110M x 256 discrete l2_prefetch_mb=$mb"; run --model 110M --envs 256 --discrete --steps 30 --warmup 3 --opt l2_prefetch_mb=$mb; done
for mb in 0 48; do echo "== 16M x 1 l2_prefetch_mb=$mb"; run --model 16M --envs 1 --domains dmcontrol --steps 300 --warmup 10 --opt l2_prefetch_mb=$mb; done
