#!/bin/bash
# full GPU suite + default bench line + step A/B after the finalize / LayerNorm latency changes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02h_gputests.log 2>&1; tail -4 gpurun_out/r02h_gputests.log
(AB_STEPS=200 timeout 300 python tools/ab_options.py 48M:64 "conv_impl=0" "conv_impl=2" "conv_impl=2"
 AB_STEPS=50 timeout 300 python tools/ab_options.py 206M:128 "conv_impl=2"
 AB_STEPS=50 timeout 300 python tools/ab_options.py 110M:256 "conv_impl=2"
 AB_STEPS=200 timeout 300 python tools/ab_options.py 16M:1 "conv_impl=2") > gpurun_out/r02h_ab_step.log 2>&1
cat gpurun_out/r02h_ab_step.log
timeout 600 python bench.py > gpurun_out/r02h_bench_default_48M_B64.json 2> gpurun_out/r02h_bench.err; tail -c 1500 gpurun_out/r02h_bench_default_48M_B64.json
