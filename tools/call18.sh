#!/bin/bash
mkdir -p gpurun_out
(AB_STEPS=40 timeout 400 python tools/ab_options.py 206M:128 "state_fuse=0" "state_fuse=2" "state_fuse=0" "state_fuse=2"
 AB_STEPS=40 timeout 400 python tools/ab_options.py 110M:256 "state_fuse=0" "state_fuse=2"
 AB_STEPS=100 timeout 400 python tools/ab_options.py 48M:256 "state_fuse=0" "state_fuse=2") 2>&1 | tee gpurun_out/r02o_ab_state_fuse.log
