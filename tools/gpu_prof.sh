#!/bin/bash
# ncu capture of selected kernels of a BDPT render: KERNELS='k_shade|k_extend|k_strategies' SKIP=60 COUNT=6 SPP=4
mkdir -p gpurun_out
K=${KERNELS:-k_shade|k_extend|k_strategies}; S=${SKIP:-60}; C=${COUNT:-6}; P=${SPP:-4}; OUT=${OUT:-prof_wf}
timeout 300 python tools/prof_render.py standard bdpt $P > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"$K" -s $S -c $C -o gpurun_out/$OUT -f python tools/prof_render.py standard bdpt $P > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
