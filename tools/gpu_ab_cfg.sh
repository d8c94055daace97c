#!/bin/bash
# One configuration, as built and with compile-time variants:
#   CFG="standard bdpt 16" VARIANTS="-DX=1;-DY=2" TAG=r02w bash tools/gpu_ab_cfg.sh
T=${TAG:-ab}
CFG=${CFG:-standard bdpt 16}
mkdir -p gpurun_out
exec > >(tee gpurun_out/${T}_iter.log) 2>&1
run() { PLAIN_FIRST=1 timeout 200 python tools/prof_render.py $CFG 2>&1 | tail -4 | head -3 | cut -c1-220; }
echo "== as built ($CFG)"; run
IFS=';' read -ra V <<< "$VARIANTS"
for v in "${V[@]}"; do
  [ -z "$v" ] && continue
  echo "== variant $v"
  touch toypathtracer-games101-assignment7_b200/csrc/*.cu
  make -C toypathtracer-games101-assignment7_b200 -j8 NVEXTRA="$v" libtpt.so 2>&1 | grep -E "error"
  run
done
