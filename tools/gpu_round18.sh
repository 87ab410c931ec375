#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-8} gpurun_out/$name.log | cut -c1-600 | tee -a gpurun_out/summary.txt; }
TAILN=12 run conv 900 python -m pytest tests/test_gpu_kernels.py -q -k "conv_engine"
run net 600 python -m pytest tests/test_gpu_net.py -q -x
run timeline 300 python tools/conv_timeline.py gpurun_out/timeline.csv 64
SUO_PROFILE_DUMP=gpurun_out/per_op.csv run bench 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline
