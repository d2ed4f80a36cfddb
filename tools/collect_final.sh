#!/bin/bash
# final data collection of round 2 (late): bench lines of the named configs, prefill lines, launch lists, ncu full
mkdir -p gpurun_out
python bench.py --model 206M --envs 128 --domains mixed --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r02h_bench_206M_B128.json 2>/dev/null
python bench.py --model 110M --envs 256 --discrete --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r02h_bench_110M_B256_discrete.json 2>/dev/null
python bench.py --model 16M --envs 1 --domains dmcontrol --steps 300 --warmup 10 --no-cpu-baseline > gpurun_out/r02h_bench_16M_B1.json 2>/dev/null
for f in 206M_B128 110M_B256_discrete 16M_B1; do python -c "
import json,sys; d=json.loads(open('gpurun_out/r02h_bench_$f.json').read().strip().splitlines()[-1]); r=d['roofline']; w=d.get('whole_step') or {}
print('$f', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],4), 'state', round(r['avg_launch_us'],1), 'us', round(r['frac'],3), 'whole', round(w.get('frac',0),3))"; done
for m in "206M --envs 1" "206M --envs 8" "110M --envs 1" "48M --envs 1"; do
  n=$(echo $m | sed 's/ --envs /_B/')
  timeout 300 python tools/bench_prefill.py --model $m --tokens 50000 --rollout 200 --check 64 --reps 3 --out gpurun_out/r02h_prefill_$n.json 2>&1 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['model'], d['envs'], round(d['prefill_ms'],2), round(d['prefill_tokens_per_s']), d.get('check_max_rel_C_diff_vs_stepping'), d['rollout']['after_prefill']['p50_ms'], d['rollout']['cold']['p50_ms'])"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 200 --csv --log-file gpurun_out/r02h_prefill_launches_206M_B1.csv python tools/bench_prefill.py --model 206M --envs 1 --tokens 50000 --rollout 5 --check 0 --reps 1 > gpurun_out/ncu_pf.log 2>&1
python tools/agg_launches.py gpurun_out/r02h_prefill_launches_206M_B1.csv > gpurun_out/r02h_prefill_launches_206M_B1_summary.txt; head -20 gpurun_out/r02h_prefill_launches_206M_B1_summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 300 --csv --log-file gpurun_out/r02h_launches_48M_B64.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --profile-steps 0 > gpurun_out/ncu_step.log 2>&1
python tools/agg_launches.py gpurun_out/r02h_launches_48M_B64.csv > gpurun_out/r02h_launches_48M_B64_summary.txt; cat gpurun_out/r02h_launches_48M_B64_summary.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv_qkv_gates_seq2_kernel|prep2_kernel" -s 20 -c 2 -o gpurun_out/r02h_prefill_elementwise python tools/bench_prefill.py --model 206M --envs 1 --tokens 50000 --rollout 5 --check 0 --reps 1 > gpurun_out/ncu_pf2.log 2>&1
ncu -i gpurun_out/r02h_prefill_elementwise.ncu-rep --page raw --csv > gpurun_out/r02h_prefill_elementwise_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv_qkv_gates_pk_kernel" -s 30 -c 1 -o gpurun_out/r02h_conv_pk python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --profile-steps 0 > gpurun_out/ncu_step2.log 2>&1
ncu -i gpurun_out/r02h_conv_pk.ncu-rep --page raw --csv > gpurun_out/r02h_conv_pk_raw.csv 2>/dev/null
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02h_bench_reference_arm_48M_B64.json 2>/dev/null; tail -c 600 gpurun_out/r02h_bench_reference_arm_48M_B64.json
