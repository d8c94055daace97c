#!/bin/bash
# fresh ncu capture of the BDPT wavefront kernels (4 spp render), after a plain run of the same command
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt
PLAIN_FIRST=1 timeout 300 python tools/prof_render.py standard bdpt 16 > gpurun_out/wf_timing.log 2>&1; cat gpurun_out/wf_timing.log
timeout 300 python tools/prof_render.py standard bdpt 4 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_shade|k_extend|k_mis|k_connect|k_shadow_q|k_expand' -s 60 -c 6 -o gpurun_out/prof_wf2 python tools/prof_render.py standard bdpt 4 > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
