#!/bin/bash
# round-2 late: side-stream overlap in the tcgen05 prefill cell -- parity subset, prefill A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "prefill" > gpurun_out/r02f_tests.log 2>&1; tail -4 gpurun_out/r02f_tests.log
for o in "prefill_tc_overlap=0" "prefill_tc_overlap=1"; do
  for m in "206M --envs 1" "110M --envs 1" "48M --envs 1" "206M --envs 8"; do
  echo "== $o $m"
  timeout 200 python tools/bench_prefill.py --model $m --tokens 50000 --rollout 20 --check 0 --reps 3 --opt $o 2>&1 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['model'], d['envs'], d['options'], round(d['prefill_ms'],2), round(d['prefill_tokens_per_s']))"
  done
done > gpurun_out/r02f_ab_prefill_overlap.log 2>&1
cat gpurun_out/r02f_ab_prefill_overlap.log
