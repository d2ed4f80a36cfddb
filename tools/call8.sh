#!/bin/bash
# round-2 late: packed-fp32 step-path conv kernel (conv_impl=2): parity, A/B on the named configs, sanitizer pass r2c
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "conv_impl or side_stream" > gpurun_out/r02g_tests.log 2>&1; tail -3 gpurun_out/r02g_tests.log
(AB_STEPS=200 timeout 300 python tools/ab_options.py 48M:64 "conv_impl=0" "conv_impl=2" "conv_impl=0" "conv_impl=2"
 AB_STEPS=50 timeout 300 python tools/ab_options.py 206M:128 "conv_impl=0" "conv_impl=2"
 AB_STEPS=200 timeout 300 python tools/ab_options.py 16M:1 "conv_impl=0" "conv_impl=2") > gpurun_out/r02g_ab_conv_impl.log 2>&1
cat gpurun_out/r02g_ab_conv_impl.log
bash tools/sanitize.sh r2c > gpurun_out/r02c_sanitize_summary.txt 2>&1
cp gpurun_out/sanitize_memcheck.log gpurun_out/r02c_sanitize_memcheck.log; cp gpurun_out/sanitize_synccheck.log gpurun_out/r02c_sanitize_synccheck.log; cp gpurun_out/sanitize_racecheck.log gpurun_out/r02c_sanitize_racecheck.log
cat gpurun_out/r02c_sanitize_summary.txt
