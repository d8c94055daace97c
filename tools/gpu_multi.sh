#!/bin/bash
# Two (or N) GPUs of one box: the multi-GPU tests of the C ABI (tests/test_multi_gpu_abi.py, the unchanged main.cpp with
# TPT_GPUS), then the bench under torchrun as the driver launches it.
#   gpurun --gpus 2 --timeout 900 -- 'N=2 ROUND=r04z bash tools/gpu_multi.sh'
N=${N:-2}; R=${ROUND:-r04z}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -k "multi or two_gpus or gpus" > gpurun_out/${R}_gpu_tests_n${N}.log 2>&1; tail -3 gpurun_out/${R}_gpu_tests_n${N}.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${R}_bench_n${N}.json 2> gpurun_out/bench_n${N}.err; tail -c 1500 gpurun_out/${R}_bench_n${N}.json; tail -3 gpurun_out/bench_n${N}.err
