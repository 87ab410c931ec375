#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" | tee -a gpurun_out/summary.txt; tail -n 14 gpurun_out/$name.log | tee -a gpurun_out/summary.txt; }
run conv 300 python -m pytest tests/test_gpu_kernels.py -q -k "conv_engine"
run net 600 python -m pytest tests/test_gpu_net.py -q -s
run timing 300 python tools/time_forward.py 8 64
SUO_PROFILE_DUMP=gpurun_out/per_op.csv run bench 900 python bench.py --steps 10 --warmup 3 --frames-per-step 8 --no-cpu-baseline
run bench32 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline
ls -la gpurun_out >> gpurun_out/summary.txt
