#!/bin/bash
# 2-GPU call: every GPU test (incl. the multi-GPU C ABI and the drop-in binary), bench line at N = 1 and N = 2
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02d.log) 2>&1
nvidia-smi --query-gpu=index,name --format=csv
timeout 1500 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider --durations=8 2>&1 | tail -25 | cut -c1-300
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02d_bench.json 2> gpurun_out/bench.err; tail -c 5000 gpurun_out/r02d_bench.json; tail -5 gpurun_out/bench.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02d_bench_n2.json 2> gpurun_out/bench2.err; tail -c 5000 gpurun_out/r02d_bench_n2.json; tail -5 gpurun_out/bench2.err
