#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "prefill or real_batch or golden" > gpurun_out/r02i_tests.log 2>&1; tail -3 gpurun_out/r02i_tests.log
for m in "206M --envs 1" "48M --envs 1"; do
timeout 200 python tools/bench_prefill.py --model $m --tokens 50000 --rollout 20 --check 0 --reps 3 2>&1 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['model'], d['envs'], round(d['prefill_ms'],2), round(d['prefill_tokens_per_s']))"
done
AB_STEPS=200 timeout 300 python tools/ab_options.py 48M:64 "" ""
