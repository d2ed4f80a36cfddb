#!/bin/bash
# final verification of the round: smoke(), full GPU suite, default bench line
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02k_gputests.log 2>&1; tail -3 gpurun_out/r02k_gputests.log
timeout 600 python bench.py > gpurun_out/r02k_bench_default_48M_B64.json 2> gpurun_out/r02k_bench.err; python -c "
import json; d=json.loads(open('gpurun_out/r02k_bench_default_48M_B64.json').read().strip().splitlines()[-1]); print(round(d['value']), round(d['e2e']['value']), d['ms_per_step'], d['roofline']['frac'], d['whole_step']['frac'], d['context_prefill']['tokens_per_s'], d['clocks'])"
