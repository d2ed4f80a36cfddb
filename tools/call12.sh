#!/bin/bash
# N = 2 sanity of the final build: driver-style torchrun bench (48M x 64 per GPU) + bit-exactness check vs one GPU and the oracle
mkdir -p gpurun_out
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$1" --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "${@:2}"; }
tr 2 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/r02h_bench_2gpu_48M_B64.json 2> gpurun_out/r02h_bench_2gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/r02h_bench_2gpu_48M_B64.json').read().strip().splitlines()[-1]); print('N=2', round(d['value']), round(d['e2e']['value']), d['ms_per_step'], d['scaling'], d['clocks'])"
tr 2 tools/multi_gpu_check.py --model 110M --envs 256 --discrete --steps 4 2>&1 | grep -E "multi-gpu check" | tee gpurun_out/r02h_mgpu_bit_exact.log
tr 2 tools/multi_gpu_check.py --model 48M --envs 128 --steps 6 2>&1 | grep -E "multi-gpu check" | tee -a gpurun_out/r02h_mgpu_bit_exact.log
tr 2 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -c 300
