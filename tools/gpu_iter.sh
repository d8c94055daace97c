#!/bin/bash
# quick iteration: correctness of both tiers + timing breakdown (+ optional ncu: NCU=1)
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_exact.py tests/test_gpu_shading.py tests/test_gpu_render.py -q -x --no-header -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/t_all.log
tail -6 gpurun_out/t_all.log
PLAIN_FIRST=1 timeout 300 python tools/prof_render.py standard bdpt 16 > gpurun_out/wf_timing.log 2>&1; cat gpurun_out/wf_timing.log
PLAIN_FIRST=1 timeout 300 python tools/prof_render.py standard pt_full 64 2>&1 | tail -3 | cut -c1-250
PLAIN_FIRST=1 timeout 300 python tools/prof_render.py standard pt_shipped 64 2>&1 | tail -3 | cut -c1-250
if [ -n "$NCU" ]; then
timeout 300 python tools/prof_render.py standard bdpt 4 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_shade|k_extend|k_mis|k_connect|k_shadow_q' -s 60 -c 5 -o gpurun_out/prof_wf python tools/prof_render.py standard bdpt 4 > gpurun_out/ncu2.log 2>&1
fi
