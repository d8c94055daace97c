#!/bin/bash
# First GPU call of a new round, in one trip (about 8 minutes of box time):
#   1. the round's profile artefacts of the tree as it is (bench line, reference arm, ncu launch list, DRAM traffic)
#   2. GPU tests + BDPT timing of the shipped build, then of -DTPT_WIDE_TRIS, -DWF_BIN_ACTIVE and both (the frames are
#      already known to be unchanged: host mirror / block emulator, tests/test_host_mirror.py) with the GPU tests on top
#   3. Cornell + bunny timing of the shipped build against -DTPT_BUDGET_WALK (16 / 32 / 48 node visits per turn), then
#      the GPU tests on the last of them (bunny renders, BASELINE config 4)
#   gpurun --timeout 1200 -- 'ROUND=r02a bash tools/gpu_next_round.sh'
R=${ROUND:-r02a}
mkdir -p gpurun_out
ROUND=$R bash tools/gpu_profiles.sh
VARIANTS="-DTPT_WIDE_TRIS;-DWF_BIN_ACTIVE;-DTPT_WIDE_TRIS -DWF_BIN_ACTIVE" TESTV=1 bash tools/gpu_ab.sh > gpurun_out/${R}_ab_bdpt.log 2>&1
tail -50 gpurun_out/${R}_ab_bdpt.log | cut -c1-250
touch toypathtracer-games101-assignment7_b200/csrc/*.cu; make -C toypathtracer-games101-assignment7_b200 -j8 libtpt.so 2>&1 | grep -E "error"
VARIANTS="-DTPT_BUDGET_WALK=16;-DTPT_BUDGET_WALK=32;-DTPT_BUDGET_WALK=48" bash tools/gpu_ab_bunny.sh > gpurun_out/${R}_ab_budget_walk.log 2>&1
timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider >> gpurun_out/${R}_ab_budget_walk.log 2>&1
tail -40 gpurun_out/${R}_ab_budget_walk.log | cut -c1-250
# leave the tree's library as shipped
touch toypathtracer-games101-assignment7_b200/csrc/*.cu; make -C toypathtracer-games101-assignment7_b200 -j8 libtpt.so 2>&1 | grep -E "error"
