#!/bin/bash
# compute-sanitizer over the hot path (SURVEY.md §5 "race detection"): memcheck, racecheck, synccheck and initcheck on
# tools/sanitize_step.py. Run on the GPU box:   gpurun --timeout 1500 -- bash tools/sanitize.sh
# Logs land in gpurun_out/sanitize_<tool>.log; copy the summaries to profiles/.
# racecheck tracks shared-memory hazards only and is ~50x slower: it gets the cases whose kernels hand data between
# warps through shared memory + mbarriers (state stream ring, tcgen05 Linear pipeline, prefill cell).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CS=${CS:-/usr/local/cuda/bin/compute-sanitizer}
run() { # tool, per-case timeout, cases...
  local tool=$1 tmo=$2; shift 2
  local log=gpurun_out/sanitize_${tool}.log
  : > "$log"
  for c in "$@"; do
    echo "=== $tool :: $c" >> "$log"
    timeout "$tmo" "$CS" --tool "$tool" --error-exitcode 9 --print-limit 20 \
      python tools/sanitize_step.py "$c" >> "$log" 2>&1
    echo "=== exit $? ($c)" >> "$log"
  done
  echo "--- $tool"; grep -E "=== exit|ERROR SUMMARY|RACECHECK SUMMARY|sanitize_step" "$log"
}
R2="state_fuse1 state_fuse2 up_fuse gemm_bm64 gemm_cluster token_ring prefill_tc small_fuse gemm_2sm"
if [ "${1:-all}" = "2sm" ]; then
  run memcheck 600 gemm_2sm
  run synccheck 600 gemm_2sm
  run racecheck 900 gemm_2sm
  exit 0
fi
if [ "${1:-all}" = "r2b" ]; then  # the prefill cell on tcgen05 and the row-split cluster finalize (late round 2)
  run memcheck 900 prefill_tc small_fuse
  run synccheck 900 prefill_tc small_fuse
  run racecheck 1200 small_fuse
  # racecheck of the whole prefill_tc case ends without a report ("process didn't terminate successfully", no hazard
  # printed, also with the pre-existing kernels only); the kernels that hand data through shared memory are therefore
  # instrumented one at a time (--kernel-regex): the case runs to its oracle check each time.
  log=gpurun_out/sanitize_racecheck.log
  for k in bgemm update_scan prep_kernel pmat_kernel gate_pre gate_scan_seq conv_qkv_gates_seq finalize_seq; do
    echo "=== racecheck :: prefill_tc, kernels matching $k" >> "$log"
    timeout 900 "$CS" --tool racecheck --error-exitcode 9 --print-limit 20 --kernel-regex kns=$k \
      python tools/sanitize_step.py prefill_tc >> "$log" 2>&1
    echo "=== exit $? (prefill_tc / $k)" >> "$log"
  done
  grep -E "=== exit|RACECHECK SUMMARY|sanitize_step" "$log"
  exit 0
fi
if [ "${1:-all}" = "r2c" ]; then  # packed-fp32 conv kernels (step + prefill), single-read prep, side-stream overlap
  run memcheck 900 conv_pk conv_pk_toy prefill_tc prefill_tc_ragged
  run synccheck 900 conv_pk prefill_tc prefill_tc_ragged
  run racecheck 900 conv_pk
  log=gpurun_out/sanitize_racecheck.log
  for k in conv_qkv_gates_seq2 prep2_kernel; do
    echo "=== racecheck :: prefill_tc_ragged, kernels matching $k" >> "$log"
    timeout 900 "$CS" --tool racecheck --error-exitcode 9 --print-limit 20 --kernel-regex kns=$k \
      python tools/sanitize_step.py prefill_tc_ragged >> "$log" 2>&1
    echo "=== exit $? (prefill_tc_ragged / $k)" >> "$log"
  done
  grep -E "=== exit|RACECHECK SUMMARY|sanitize_step" "$log"
  exit 0
fi
if [ "${1:-all}" = "r2" ]; then   # only the kernels / options added in round 2
  run memcheck 600 $R2
  run synccheck 600 $R2
  run racecheck 900 state_fuse1 state_fuse2 up_fuse gemm_cluster
  exit 0
fi
run memcheck 600 fused_eager fused_graph per_token one_env discrete real_16M slstm prefill $R2
run synccheck 600 fused_eager per_token one_env real_16M slstm prefill $R2
run initcheck 600 fused_eager one_env real_16M prefill
run racecheck 900 fused_eager one_env real_16M prefill state_fuse1 state_fuse2 up_fuse gemm_cluster
