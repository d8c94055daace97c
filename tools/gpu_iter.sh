#!/bin/bash
# quick iteration: correctness of both tiers + timing breakdown + bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/t_all.log
tail -6 gpurun_out/t_all.log
PLAIN_FIRST=1 timeout 300 python tools/prof_render.py standard bdpt 16 > gpurun_out/wf_timing.log 2>&1; cat gpurun_out/wf_timing.log
PLAIN_FIRST=1 timeout 300 python tools/prof_render.py standard pt_full 64 2>&1 | tail -3 | cut -c1-250
PLAIN_FIRST=1 timeout 300 python tools/prof_render.py standard pt_shipped 64 2>&1 | tail -3 | cut -c1-250
if [ -n "$BENCH" ]; then timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err; fi
