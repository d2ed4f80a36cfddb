timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
run() { timeout 300 python tools/bench_prefill.py --model 206M --envs 1 --rollout 10 --check 0 --reps 2 "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['options'], round(d['prefill_ms'],1), 'ms', round(d['prefill_tokens_per_s']))"; }
run
run --opt prefill_rows=2048
python bench.py --steps 300 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('headline', d['value'], d['ms_per_step'], d['e2e']['value'], d.get('whole_step'))"
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 300 --csv --log-file gpurun_out/r02_prefill_tc_launches_206M_B1.csv python tools/bench_prefill.py --model 206M --envs 1 --tokens 6048 --check 0 --rollout 2 --reps 1 --opt prefill_rows=2048 > /dev/null 2>&1; python tools/agg_launches.py gpurun_out/r02_prefill_tc_launches_206M_B1.csv > gpurun_out/r02_prefill_tc_launches_206M_B1_summary.txt; head -8 gpurun_out/r02_prefill_tc_launches_206M_B1_summary.txt
