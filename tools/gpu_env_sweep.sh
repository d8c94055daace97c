#!/bin/bash
# One environment variable swept over values on one configuration: VAR=TPT_LONG_WAIT VALUES="0 4 8 16" CFG="bunny pt_full 32" TAG=r04i bash tools/gpu_env_sweep.sh
T=${TAG:-envsweep}
mkdir -p gpurun_out
exec > >(tee gpurun_out/${T}_iter.log) 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
[ -z "$NOTEST" ] && timeout 900 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider -k "${TESTS:-bunny or pathtrace or deterministic}" 2>&1 | tail -4 | cut -c1-300
for v in $VALUES; do
  echo "== $VAR=$v"
  env $VAR=$v PLAIN_FIRST=1 timeout 200 python tools/prof_render.py ${CFG:-bunny pt_full 32} 2>&1 | tail -4 | head -3 | cut -c1-220
done
