#!/bin/bash
# run-time A/B of the BDPT launch schedule (no rebuild): ENVS="A=1;B=2 C=3" + step counters of the build
T=${TAG:-iter}
mkdir -p gpurun_out
exec > >(tee gpurun_out/${T}_iter.log) 2>&1
run() { for cfg in "standard bdpt 16" "bunny bdpt 8"; do env $1 PLAIN_FIRST=1 timeout 300 python tools/prof_render.py $cfg 2>&1 | grep "no per-kernel" | cut -c1-200; done; }
echo "== default"; run "X=1"
IFS=';' read -ra V <<< "$ENVS"
for v in "${V[@]}"; do [ -z "$v" ] && continue; echo "== $v"; run "$v"; done
if [ -n "$COUNTERS" ]; then
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum
timeout 300 python tools/prof_render.py standard bdpt 16 > gpurun_out/plain2.log 2>&1 && \
ncu --metrics $M --clock-control none -k regex:'k_' -c 1300 --csv --log-file gpurun_out/${T}_step_counters.csv python tools/prof_render.py standard bdpt 16 > gpurun_out/ncu2.log 2>&1
tail -n 2 gpurun_out/ncu2.log
fi
