B="python bench.py --steps 100 --warmup 5 --no-cpu-baseline --profile-steps 0 --opt gemm_up_bn=64 --opt gemm_up_splits=1 --opt gemm_down_bn=64 --opt gemm_down_splits=4"
for o in 0 1 2 4 8 16 32 55 63 47 31; do
  echo "== skip $o"; $B --opt debug_skip=$o | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step']*1000,1))"
done
