#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_smallm.py tests/test_gpu_parity.py -x -q -m gpu -k "small or one_env or b1 or predict or golden or real_batch or rollout" > gpurun_out/r02j_tests.log 2>&1; tail -3 gpurun_out/r02j_tests.log
(AB_STEPS=300 timeout 200 python tools/ab_options.py 16M:1 "" ""
 AB_STEPS=300 timeout 200 python tools/ab_options.py 48M:1 ""
 AB_STEPS=200 timeout 200 python tools/ab_options.py 206M:1 ""
 AB_STEPS=300 timeout 200 python tools/ab_options.py 16M:4 ""
 AB_STEPS=200 timeout 200 python tools/ab_options.py 48M:64 "") 2>&1 | tee gpurun_out/r02j_ab_one_env.log
