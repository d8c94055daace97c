#!/bin/bash
# the artefacts profiles/ holds for a round: bench line, launch list of the bench command, per-launch counters of one
# step (instructions, fp32 flops, DRAM and L2 bytes: tools/ncu_step_counters.py + tools/ncu_traffic.py read the CSV)
mkdir -p gpurun_out
R=${ROUND:-r01f}
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${R}_bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/${R}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>> gpurun_out/bench.err; tail -c 700 gpurun_out/${R}_bench_reference.json
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 125 -c 125 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1   # the second frame of the command (a frame is 123 + 2 launches)
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum
timeout 300 python tools/prof_render.py standard bdpt 16 > gpurun_out/plain2.log 2>&1 && \
ncu --metrics $M --clock-control none -k regex:'k_' -c 1300 --csv --log-file gpurun_out/${R}_step_counters.csv python tools/prof_render.py standard bdpt 16 > gpurun_out/ncu2.log 2>&1
tail -n 2 gpurun_out/ncu1.log; tail -n 2 gpurun_out/ncu2.log
