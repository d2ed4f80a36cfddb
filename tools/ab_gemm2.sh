B="python bench.py --steps 100 --warmup 5 --no-cpu-baseline --profile-steps 0"
run() { echo "== up $1/$2 down $3/$4"; $B --opt gemm_up_bn=$1 --opt gemm_up_splits=$2 --opt gemm_down_bn=$3 --opt gemm_down_splits=$4 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step']*1000,1))"; }
for d in "64 2" "64 3" "64 4" "64 6" "64 8" "32 2" "32 4" "128 4" "128 6"; do run 64 1 $d; done
for u in "64 2" "64 3" "128 1" "128 2" "32 1" "32 2"; do run $u 64 4; done
