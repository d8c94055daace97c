#!/bin/bash
# Generic A/B on the B200: GPU tests selected by TESTS, then the configurations in CFGS ("scene mode spp;...") as built and
# for every compile-time variant in VARIANTS ("-DX=1;-DX=2").
#   CFGS="bunny pt_full 32;bunny bdpt 8" VARIANTS="-DTPT_PEND_WAIT=0;-DTPT_PEND_WAIT=6" TAG=r04d bash tools/gpu_ab.sh
T=${TAG:-ab}
mkdir -p gpurun_out
exec > >(tee gpurun_out/${T}_iter.log) 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
[ -z "$NOTEST" ] && timeout 900 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider -k "${TESTS:-bunny or pathtrace or deterministic or small_renders}" 2>&1 | tail -4 | cut -c1-300
run() {
  IFS=';' read -ra C <<< "${CFGS:-bunny pt_full 32}"
  for cfg in "${C[@]}"; do
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
