#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_render.py -q --no-header -p no:cacheprovider > gpurun_out/t_render2.log 2>&1; echo "exit $?" >> gpurun_out/t_render2.log
tail -25 gpurun_out/t_render2.log
timeout 300 python tools/prof_render.py standard bdpt 16 > gpurun_out/wf_timing.log 2>&1; cat gpurun_out/wf_timing.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench exit $?"; cat gpurun_out/bench1.json; tail -5 gpurun_out/bench1.err
timeout 300 python tools/prof_render.py standard bdpt 2 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python tools/prof_render.py standard bdpt 2 > gpurun_out/ncu1.log 2>&1
timeout 300 python tools/prof_render.py standard bdpt 2 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_shade|k_extend|k_mis|k_connect|k_shadow_q' -s 40 -c 10 -o gpurun_out/prof_wf python tools/prof_render.py standard bdpt 2 > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out
