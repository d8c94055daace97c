#!/bin/bash
# The short closing check of a tree whose kernels were measured before (tools/gpu_final.sh): GPU tests, smoke(), the
# driver's bench command and the reference arm.   gpurun --timeout 900 -- 'ROUND=r04y bash tools/gpu_close.sh'
R=${ROUND:-close}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/${R}_gpu_tests.log 2>&1; tail -3 gpurun_out/${R}_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; tail -4 gpurun_out/${R}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench.json 2> gpurun_out/bench.err; head -c 700 gpurun_out/${R}_bench.json; echo; tail -2 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>> gpurun_out/bench.err; head -c 400 gpurun_out/${R}_bench_reference.json; echo
