set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
B="python bench.py --steps 100 --warmup 5 --no-cpu-baseline"
for o in "" "--opt gemm_splitk=1" "--opt gemm_up_splits=1" "--opt gemm_down_splits=1" "--opt gemm_down_bn=128 --opt gemm_down_splits=8" "--opt gemm_up_bn=128 --opt gemm_up_splits=2" "--opt gemm_down_bn=32 --opt gemm_down_splits=3" "--opt gemm_up_bn=64 --opt gemm_up_splits=1 --opt gemm_down_bn=64 --opt gemm_down_splits=4"; do
  echo "== $o"; $B $o | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],4), round(d['roofline']['avg_launch_us'],1), round(d['roofline']['share_of_step'],3))"
done
