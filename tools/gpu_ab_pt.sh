#!/bin/bash
# A/B compile-time variants on the PathTrace pipeline: VARIANTS="-DX=1;-DX=2" bash tools/gpu_ab_pt.sh   (first: as shipped + GPU tests)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/t_all.log
tail -4 gpurun_out/t_all.log | cut -c1-300
run() { for m in pt_full pt_shipped; do PLAIN_FIRST=1 timeout 200 python tools/prof_render.py standard $m 64 2>&1 | tail -4 | head -3 | cut -c1-200; done; }
echo "== build as shipped"; run
IFS=';' read -ra V <<< "$VARIANTS"
for v in "${V[@]}"; do
  echo "== variant $v"
  touch toypathtracer-games101-assignment7_b200/csrc/*.cu
  make -C toypathtracer-games101-assignment7_b200 -j8 NVEXTRA="$v" libtpt.so 2>&1 | grep -E "error"
  run
done
