#!/bin/bash
# The named multi-GPU configs of BASELINE.json on ONE 8 x B200 box (gpurun --gpus 8 -- bash tools/run_multi_gpu_configs.sh):
#   configs[2]  206M x 128 mixed-domain envs per GPU, weak scaling, N = 1, 4, 8 (1024 envs at N = 8)
#   configs[4]  110M, 256 Atari-style envs in total, discrete head, strong scaling 256/128/64/32 per GPU, N = 1, 2, 4, 8,
#               + bit-exactness of the sharded tokens vs one GPU and vs the fp32 oracle (tools/multi_gpu_check.py)
#   configs[1]  48M x 64 per GPU, weak, N = 8 (the driver's own scaling run covers N = 1, 2, 4, 8 of this one)
# One JSON line per run -> gpurun_out/r02_mgpu_<tag>.json
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$1" --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "${@:2}"; }
run() { # tag, N, bench args...
  local tag=$1 n=$2; shift 2
  if [ "$n" = 1 ]; then python bench.py --gpus 1 --no-cpu-baseline "$@" > gpurun_out/r02_mgpu_${tag}.json 2> gpurun_out/r02_mgpu_${tag}.err
  else tr "$n" bench.py --gpus "$n" --no-cpu-baseline "$@" > gpurun_out/r02_mgpu_${tag}.json 2> gpurun_out/r02_mgpu_${tag}.err; fi
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r02_mgpu_{tag}.json").read().strip().splitlines()[-1])
    print(tag, "n_gpus", d["n_gpus"], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1),
          "scaling", d["scaling"], "envs/gpu", d["config"]["envs_per_gpu"], "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as ex:
    print(tag, "FAILED", ex)
PY
}
for n in 8 4 2 1; do run 110M_strong256_N$n $n --model 110M --envs 256 --scaling strong --discrete --domains mixed --steps 100 --warmup 5; done
for n in 8 4 1; do run 206M_weak128_N$n $n --model 206M --envs 128 --domains mixed --steps 40 --warmup 5; done
run 48M_weak64_N8 8 --steps 200 --warmup 10
for n in 8 4; do tr $n tools/multi_gpu_check.py --model 110M --envs 256 --discrete --steps 4 2>&1 | grep -E "multi-gpu check" | tee -a gpurun_out/r02_mgpu_bit_exact.log; done
tr 8 tools/multi_gpu_check.py --model 206M --envs 256 --steps 3 2>&1 | grep -E "multi-gpu check" | tee -a gpurun_out/r02_mgpu_bit_exact.log
tr 8 tools/multi_gpu_check.py --model 48M --envs 512 --steps 6 --every 4 2>&1 | grep -E "multi-gpu check" | tee -a gpurun_out/r02_mgpu_bit_exact.log
