#!/bin/bash
# first GPU trip: parity tests of what exists + timing of the validation renderer
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
timeout 900 python -m pytest tests/test_gpu_exact.py tests/test_gpu_shading.py -q -x --no-header -p no:cacheprovider > gpurun_out/t_exact.log 2>&1; echo "exit $?" >> gpurun_out/t_exact.log
timeout 900 python -m pytest tests/test_gpu_render.py -q --no-header -p no:cacheprovider -k "megakernel" > gpurun_out/t_render.log 2>&1; echo "exit $?" >> gpurun_out/t_render.log
timeout 600 python - > gpurun_out/mega_timing.log 2>&1 <<'PY'
import sys, time, numpy as np
sys.path.insert(0, '.')
import tpt_b200 as T
s = T.Scene('standard', 784, 784)
for mode, spp in (('bdpt', 16), ('pt_full', 64), ('pt_shipped', 64)):
    for rep in range(2):
        t0 = time.time(); img, st = s.render(mode, spp, pipeline=T.PIPE_MEGAKERNEL); dt = time.time() - t0
        print(mode, spp, 'device_ms %.2f wall %.3f Msamples/s %.2f ref_rays %d traced %d mean %s' % (st['device_ms'], dt, st['samples'] / st['device_ms'] / 1e3, st['ref_rays'], st['traced_rays'], img.mean((0, 1))))
PY
tail -5 gpurun_out/t_exact.log gpurun_out/t_render.log gpurun_out/mega_timing.log
