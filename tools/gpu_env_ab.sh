#!/bin/bash
# A/B run-time variants (environment switches of wavefront.cu) on the GPU box:
#   ENVS="TPT_WF_CGRID=8;TPT_WF_CGRID=2" bash tools/gpu_env_ab.sh      (first: defaults, with the GPU tests)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/t_all.log
tail -12 gpurun_out/t_all.log | cut -c1-300
echo "== defaults"; PLAIN_FIRST=1 timeout 200 python tools/prof_render.py standard bdpt 16 2>&1 | tail -4 | cut -c1-250
IFS=';' read -ra V <<< "$ENVS"
for v in "${V[@]}"; do
  echo "== $v"
  env $v PLAIN_FIRST=1 timeout 200 python tools/prof_render.py standard bdpt 16 2>&1 | tail -4 | head -3 | cut -c1-250
done
