#!/bin/bash
# The bunny PathTrace configurations (hierarchy walk) as built and with compile-time variants, after the GPU tests of
# the exact and render tiers that touch them: VARIANTS="-DX=1;-DX=2" TAG=r02s bash tools/gpu_ab_bunny_pt.sh
T=${TAG:-bunny}
mkdir -p gpurun_out
exec > >(tee gpurun_out/${T}_iter.log) 2>&1
[ -z "$NOTEST" ] && timeout 600 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider -k "bunny or pathtrace or deterministic or small_renders" 2>&1 | tail -4 | cut -c1-300
run() {
  for cfg in "bunny pt_full 32" "bunny pt_shipped 32" "standard pt_full 64"; do
    [ -n "$ONLY_FULL" ] && [ "$cfg" != "bunny pt_full 32" ] && continue
    PLAIN_FIRST=1 timeout 200 python tools/prof_render.py $cfg 2>&1 | tail -4 | head -3 | cut -c1-220
  done
}
echo "== as built"; run
IFS=';' read -ra V <<< "$VARIANTS"
for v in "${V[@]}"; do
  [ -z "$v" ] && continue
  echo "== variant $v"
  touch toypathtracer-games101-assignment7_b200/csrc/*.cu
  make -C toypathtracer-games101-assignment7_b200 -j8 NVEXTRA="$v" libtpt.so 2>&1 | grep -E "error"
  run
done
