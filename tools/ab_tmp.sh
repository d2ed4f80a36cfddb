timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
run() { timeout 300 python tools/bench_prefill.py --model ${MODEL:-206M} --envs ${ENVS:-1} --rollout 10 --check "${CHK:-0}" --reps 2 "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['model'], d['envs'], d['options'], round(d['prefill_ms'],1), 'ms', round(d['prefill_tokens_per_s']), d.get('check_max_rel_C_diff_vs_stepping'))"; }
CHK=64 run
ENVS=8 run
MODEL=110M run
MODEL=48M run
bash tools/sanitize.sh 2sm 2>&1 | tail -12
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 200 --csv --log-file gpurun_out/r02_prefill_tc_launches_206M_B1.csv python tools/bench_prefill.py --model 206M --envs 1 --tokens 49152 --check 0 --rollout 2 --reps 1 > /dev/null 2>&1; python tools/agg_launches.py gpurun_out/r02_prefill_tc_launches_206M_B1.csv > gpurun_out/r02_prefill_tc_launches_206M_B1_summary.txt; head -8 gpurun_out/r02_prefill_tc_launches_206M_B1_summary.txt
