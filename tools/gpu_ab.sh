#!/bin/bash
# A/B compile-time variants on the GPU box: VARIANTS="-DX=1;-DX=2" bash tools/gpu_ab.sh   (first: the build as shipped)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/t_all.log
tail -12 gpurun_out/t_all.log | cut -c1-300
echo "== baseline build"; PLAIN_FIRST=1 timeout 200 python tools/prof_render.py standard bdpt 16 2>&1 | tail -3
IFS=';' read -ra V <<< "$VARIANTS"
for v in "${V[@]}"; do
  echo "== variant $v"
  touch toypathtracer-games101-assignment7_b200/csrc/*.cu
  make -C toypathtracer-games101-assignment7_b200 -j8 NVEXTRA="$v" libtpt.so 2>&1 | grep -E "error" 
  PLAIN_FIRST=1 timeout 200 python tools/prof_render.py standard bdpt 16 2>&1 | tail -3
  if [ -n "$TESTV" ]; then timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -8 | cut -c1-300; fi
done
