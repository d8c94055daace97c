#!/bin/bash
# A/B compile-time variants on the bunny scene (hierarchy walk): VARIANTS="-DX=1;-DX=2" bash tools/gpu_ab_bunny.sh
mkdir -p gpurun_out
run() { PLAIN_FIRST=1 timeout 200 python tools/prof_render.py bunny pt_full 32 2>&1 | tail -4 | head -3 | cut -c1-200; PLAIN_FIRST=1 timeout 200 python tools/prof_render.py bunny bdpt 8 2>&1 | tail -4 | head -3 | cut -c1-200; }
echo "== build as shipped"; run
IFS=';' read -ra V <<< "$VARIANTS"
for v in "${V[@]}"; do
  echo "== variant $v"
  touch toypathtracer-games101-assignment7_b200/csrc/*.cu
  make -C toypathtracer-games101-assignment7_b200 -j8 NVEXTRA="$v" libtpt.so 2>&1 | grep -E "error"
  run
done
