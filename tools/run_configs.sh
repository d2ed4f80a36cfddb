# bench lines of the non-headline BASELINE.json configs (single GPU shards)
python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_48M_B64.json
python bench.py --model 206M --envs 128 --domains mixed --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_206M_B128.json
python bench.py --model 110M --envs 256 --discrete --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_110M_B256_discrete.json
python bench.py --model 16M --envs 1 --domains dmcontrol --steps 300 --warmup 10 --no-cpu-baseline > gpurun_out/bench_16M_B1.json
for f in 48M_B64 206M_B128 110M_B256_discrete 16M_B1; do python -c "
import json,sys; d=json.load(open('gpurun_out/bench_$f.json')); r=d['roofline']
print('$f', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],3), 'state', round(r['avg_launch_us'],1), 'us', round(r['achieved']), 'GB/s', round(r['frac'],3), 'share', round(r['share_of_step'],3))"; done
