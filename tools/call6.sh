#!/bin/bash
# round-2 late: persistent conv kernel, pipelined prep2, SFU SiLU default -- parity subset, prefill A/B, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "prefill" > gpurun_out/r02e_tests.log 2>&1; tail -4 gpurun_out/r02e_tests.log
for o in "prefill_conv_persist=0" "prefill_conv_persist=1" "prefill_conv_persist=1 --opt prefill_conv_run=32" "prefill_conv_persist=1 --opt prefill_conv_run=8"; do
  echo "== $o"
  timeout 200 python tools/bench_prefill.py --model 206M --envs 1 --tokens 50000 --rollout 20 --check 0 --reps 3 --opt $o 2>&1 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['options'], round(d['prefill_ms'],2), round(d['prefill_tokens_per_s']))"
done > gpurun_out/r02e_ab_prefill.log 2>&1
cat gpurun_out/r02e_ab_prefill.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 200 --csv --log-file gpurun_out/r02e_prefill_launches_206M_B1.csv python tools/bench_prefill.py --model 206M --envs 1 --tokens 50000 --rollout 5 --check 0 --reps 1 > gpurun_out/ncu_pf.log 2>&1
python tools/agg_launches.py gpurun_out/r02e_prefill_launches_206M_B1.csv > gpurun_out/r02e_prefill_launches_206M_B1_summary.txt; head -8 gpurun_out/r02e_prefill_launches_206M_B1_summary.txt
