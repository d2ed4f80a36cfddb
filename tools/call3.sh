#!/bin/bash
# refresh of the non-headline bench lines + launch list of the current default step; run under gpurun
bash tools/run_configs.sh
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_48M_B64_v5.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-steps 0 > gpurun_out/ncu_v5.log 2>&1
python tools/agg_launches.py gpurun_out/launches_48M_B64_v5.csv > gpurun_out/launches_48M_B64_v5_summary.txt; head -12 gpurun_out/launches_48M_B64_v5_summary.txt
