#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "scan_epilogue or prefill_cells" > gpurun_out/r02l_tests.log 2>&1; tail -3 gpurun_out/r02l_tests.log
for o in "prefill_scan_split=2" "prefill_scan_split=4"; do
  for m in "206M --envs 1" "206M --envs 8" "48M --envs 1"; do
  timeout 200 python tools/bench_prefill.py --model $m --tokens 50000 --rollout 20 --check 0 --reps 3 --opt $o 2>&1 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['model'], d['envs'], d['options'], round(d['prefill_ms'],2), round(d['prefill_tokens_per_s']))"
  done
done 2>&1 | tee gpurun_out/r02l_ab_scan_split.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 120 --csv --log-file gpurun_out/r02l_launches.csv python tools/bench_prefill.py --model 206M --envs 1 --tokens 50000 --rollout 5 --check 0 --reps 1 --opt prefill_scan_split=4 > gpurun_out/ncu_pf.log 2>&1
python tools/agg_launches.py gpurun_out/r02l_launches.csv | head -6
