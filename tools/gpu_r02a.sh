#!/bin/bash
# Round 2, first GPU call: (1) the round's profile artefacts of the tree as it is, (2) `ncu --set full` of EVERY kernel of
# the shipped build — BDPT (Cornell), PathTrace (Cornell), PathTrace + BDPT traversal on the bunny — with the L2 / issue
# counters the verdict asked for, (3) A/B of the experiments built in round 1, GPU tests on each.
#   gpurun --timeout 1500 -- 'bash tools/gpu_r02a.sh'
R=${ROUND:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${R}_gpu.txt
ROUND=$R bash tools/gpu_profiles.sh
EXTRA=lts__t_bytes.sum,lts__t_sectors_srcunit_tex_lookup_hit.sum,lts__t_sectors_srcunit_tex_lookup_miss.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum,sm__inst_issued.sum,sm__inst_issued.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__cycles_active.avg
cap() {   # name scene mode spp kernels skip count
  timeout 300 python tools/prof_render.py $2 $3 $4 > gpurun_out/plain_$1.log 2>&1 && \
  timeout 900 ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:"$5" -s $6 -c $7 -o gpurun_out/${R}_$1 -f python tools/prof_render.py $2 $3 $4 > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log | cut -c1-200
}
cap bdpt standard bdpt 4 'k_shade|k_extend|k_expand|k_connect|k_shadow_q|k_mis' 120 6
cap pt standard pt_full 4 'k_pt_shade|k_pt_extend|k_pt_shadow' 30 6
cap bunny_pt bunny pt_full 2 'k_pt_shade|k_pt_extend|k_pt_shadow' 24 6
cap bunny_bdpt bunny bdpt 2 'k_shade|k_extend|k_expand|k_connect|k_shadow_q|k_mis' 60 6
VARIANTS="-DTPT_WIDE_TRIS;-DWF_BIN_ACTIVE;-DTPT_WIDE_TRIS -DWF_BIN_ACTIVE" TESTV=1 bash tools/gpu_ab.sh > gpurun_out/${R}_ab_bdpt.log 2>&1
cut -c1-250 gpurun_out/${R}_ab_bdpt.log | grep -v "^$" | tail -60
touch toypathtracer-games101-assignment7_b200/csrc/*.cu; make -C toypathtracer-games101-assignment7_b200 -j8 libtpt.so 2>&1 | grep -E "error"
VARIANTS="-DTPT_BUDGET_WALK=16;-DTPT_BUDGET_WALK=32;-DTPT_BUDGET_WALK=48" bash tools/gpu_ab_bunny.sh > gpurun_out/${R}_ab_budget_walk.log 2>&1
timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider >> gpurun_out/${R}_ab_budget_walk.log 2>&1
cut -c1-250 gpurun_out/${R}_ab_budget_walk.log | tail -40
touch toypathtracer-games101-assignment7_b200/csrc/*.cu; make -C toypathtracer-games101-assignment7_b200 -j8 libtpt.so 2>&1 | grep -E "error"
ls -la gpurun_out | tail -30
