#!/bin/bash
# PathTrace A/B: GPU tests of the tree as built, then Cornell / bunny PathTrace timings per compile-time variant
T=${TAG:-iter}
mkdir -p gpurun_out
exec > >(tee gpurun_out/${T}_iter.log) 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -6 | cut -c1-300
run() {
  for cfg in "standard pt_full 64" "standard pt_shipped 64" "bunny pt_full 32" "bunny pt_shipped 32"; do
    PLAIN_FIRST=1 timeout 300 python tools/prof_render.py $cfg 2>&1 | tail -4 | head -2 | cut -c1-200
  done
}
echo "== as built"; run
IFS=';' read -ra V <<< "$VARIANTS"
for v in "${V[@]}"; do
  [ -z "$v" ] && continue
  echo "== variant $v"
  touch toypathtracer-games101-assignment7_b200/csrc/pt_wavefront.cu
  make -C toypathtracer-games101-assignment7_b200 -j8 NVEXTRA="$v" libtpt.so 2>&1 | grep -E "error"
  run
done
