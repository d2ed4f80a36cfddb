#!/bin/bash
# one env: state-stream row split (CTA count of the stream kernel vs partial sums in finalize); run under gpurun
for m in 16M 48M; do for rs in 0 1 2 4 8 16; do echo "== $m x 1 env state_rows_split=$rs"; timeout 100 python bench.py --model $m --envs 1 --domains dmcontrol --steps 300 --warmup 10 --no-cpu-baseline --profile-steps 0 --opt state_rows_split=$rs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(round(d['value']), round(d['ms_per_step']*1e3,1), 'us/step')"; done; done
