#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "prefill or rollout or persist" > gpurun_out/r02n_tests.log 2>&1; tail -3 gpurun_out/r02n_tests.log
for m in "206M --envs 1" "206M --envs 8" "48M --envs 64 --tokens 300" "110M --envs 3"; do
  timeout 300 python tools/bench_prefill.py --model $m --tokens 50000 --rollout 20 --check 16 --reps 3 2>&1 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['model'], d['envs'], d['context_tokens'], round(d['prefill_ms'],2), round(d['prefill_tokens_per_s']), d.get('check_max_rel_C_diff_vs_stepping'))"
done 2>&1 | tee gpurun_out/r02n_prefill.log
