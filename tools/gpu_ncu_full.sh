#!/bin/bash
# `ncu --set full` (+ L2 / issue counters, source) of every kernel of a mid-frame round of the tree as built: BDPT (Cornell),
# PathTrace (Cornell), PathTrace (bunny: budget + long kernels) — the capture part of tools/gpu_final.sh on its own.
#   gpurun --timeout 900 -- 'ROUND=r05h bash tools/gpu_ncu_full.sh'
R=${ROUND:-r05h}
mkdir -p gpurun_out
EXTRA=lts__t_bytes.sum,lts__t_sectors_srcunit_tex_lookup_hit.sum,lts__t_sectors_srcunit_tex_lookup_miss.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum,sm__inst_issued.sum,sm__inst_issued.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__cycles_active.avg
cap() {   # name scene mode spp kernels skip count
  timeout 300 python tools/prof_render.py $2 $3 $4 > gpurun_out/plain_$1.log 2>&1 && \
  timeout 900 ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:"$5" -s $6 -c $7 -o gpurun_out/${R}_$1 -f python tools/prof_render.py $2 $3 $4 > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log | cut -c1-200
}
cap bdpt standard bdpt 8 'k_path|k_expand|k_connect|k_shadow_q|k_mis' 25 5
cap pt standard pt_full 4 'k_pt_shade|k_pt_extend|k_pt_shadow' 30 6
cap bunny_pt bunny pt_full 2 'k_pt_' 30 10
ls -la gpurun_out/${R}_*
