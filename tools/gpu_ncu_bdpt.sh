#!/bin/bash
# `ncu --set full` (+ L2 / issue counters, source) of one mid-frame round of the BDPT pipeline: k_path, k_expand, k_connect,
# k_shadow_q, k_mis.   R=r02k bash tools/gpu_ncu_bdpt.sh
R=${R:-r02k}
mkdir -p gpurun_out
EXTRA=lts__t_bytes.sum,lts__t_sectors_srcunit_tex_lookup_hit.sum,lts__t_sectors_srcunit_tex_lookup_miss.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum,sm__inst_issued.sum,sm__inst_issued.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__cycles_active.avg
timeout 300 python tools/prof_render.py standard bdpt 8 > gpurun_out/plain_bdpt.log 2>&1 && \
timeout 900 ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:'k_path|k_expand|k_connect|k_shadow_q|k_mis' -s 25 -c 5 -o gpurun_out/${R}_bdpt -f python tools/prof_render.py standard bdpt 8 > gpurun_out/ncu_bdpt.log 2>&1
tail -2 gpurun_out/ncu_bdpt.log | cut -c1-200
