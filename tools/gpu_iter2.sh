#!/bin/bash
# Round 2 iteration call: GPU tests, then the timing set (Cornell BDPT / PathTrace, glass BDPT, bunny PathTrace / BDPT)
# of the tree as built, then of each compile-time variant in VARIANTS="-DX=1;-DY" (rebuilt on the box; TESTV=1 runs the
# GPU tests on each variant too).  TAG names the log: gpurun_out/${TAG}_iter.log
T=${TAG:-iter}
mkdir -p gpurun_out
exec > >(tee gpurun_out/${T}_iter.log) 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -8 | cut -c1-300
run() {
  for cfg in "standard bdpt 16" "standard pt_full 64" "standard pt_shipped 64" "refractive bdpt 16" "bunny pt_full 32" "bunny bdpt 8"; do
    PLAIN_FIRST=1 timeout 300 python tools/prof_render.py $cfg 2>&1 | tail -4 | head -3 | cut -c1-260
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
  if [ -n "$TESTV" ]; then timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -8 | cut -c1-300; fi
done
if [ -n "$BENCH" ]; then
  touch toypathtracer-games101-assignment7_b200/csrc/*.cu; make -C toypathtracer-games101-assignment7_b200 -j8 libtpt.so 2>&1 | grep -E "error"
  timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/${T}_bench.json; tail -5 gpurun_out/bench.err
fi
