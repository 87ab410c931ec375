#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" | tee -a gpurun_out/summary.txt; tail -n 12 gpurun_out/$name.log | tee -a gpurun_out/summary.txt; }
run conv 300 python -m pytest tests/test_gpu_kernels.py -q -k "conv_engine and (1-3- or 2-3-)"
run g2o 300 python -m pytest tests/test_gpu_geom.py -q -k "g2o"
run timeline 300 python tools/conv_timeline.py gpurun_out/timeline.csv
run timing 300 python tools/time_forward.py 8 64
ls -la gpurun_out >> gpurun_out/summary.txt
