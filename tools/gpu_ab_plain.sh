#!/bin/bash
# A/B compile-time variants, un-instrumented frame time (two plain renders, the second reported) + the GPU tests per variant:
#   VARIANTS="-DX=1;-DX=2" bash tools/gpu_ab_plain.sh
mkdir -p gpurun_out
run() { PLAIN_FIRST=1 timeout 200 python tools/prof_render.py standard bdpt 16 2>&1 | head -1; }
echo "== build as shipped"; run
IFS=';' read -ra V <<< "$VARIANTS"
for v in "${V[@]}"; do
  echo "== variant $v"
  touch toypathtracer-games101-assignment7_b200/csrc/*.cu
  make -C toypathtracer-games101-assignment7_b200 -j8 NVEXTRA="$v" libtpt.so 2>&1 | grep -E "error"
  run
  if [ -n "$TESTV" ]; then timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -2 | cut -c1-200; fi
done
