#!/bin/bash
mkdir -p gpurun_out
(AB_STEPS=200 timeout 900 python tools/ab_options.py 48M:64 "" "gemm_up_bn=64,gemm_up_splits=2" "gemm_up_bn=128,gemm_up_splits=2" "gemm_up_bn=128,gemm_up_splits=3" "gemm_up_bn=128,gemm_up_splits=1" "gemm_up_bn=256,gemm_up_splits=4" "gemm_down_splits=4" "gemm_down_splits=8" "gemm_down_bn=128,gemm_down_splits=6" "gemm_down_bn=32,gemm_down_splits=3" "" ) 2>&1 | tee gpurun_out/r02m_ab_gemm_plan.log
