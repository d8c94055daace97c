#!/bin/bash
# First GPU call of a new round, in one trip (about 4 minutes of box time):
#   1. the round's profile artefacts of the tree as it is (bench line, reference arm, ncu launch list, DRAM traffic)
#   2. GPU tests + timing of the shipped build, then of every compile-time experiment that the host mirror has already
#      shown to be exact (tests/test_host_mirror.py EXPERIMENTS), each with the GPU tests on top
#   gpurun --timeout 900 -- 'ROUND=r02a bash tools/gpu_next_round.sh'
R=${ROUND:-r02a}
mkdir -p gpurun_out
ROUND=$R bash tools/gpu_profiles.sh
VARIANTS=${VARIANTS:--DTPT_WIDE_TRIS} TESTV=1 bash tools/gpu_ab.sh > gpurun_out/${R}_ab.log 2>&1
tail -40 gpurun_out/${R}_ab.log | cut -c1-250
