#!/bin/bash
# one env, 6-kernel path: marginal in-graph cost of each kernel class (debug_skip: results garbage, timing valid)
for m in 16M 48M; do for sk in 0 1 2 4 8 16 32 63; do echo "== $m x 1 env debug_skip=$sk"; timeout 100 python bench.py --model $m --envs 1 --domains dmcontrol --steps 300 --warmup 10 --no-cpu-baseline --profile-steps 0 --opt smallm=0 --opt debug_skip=$sk 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(round(d['ms_per_step']*1e3,1), 'us/step')"; done; done
