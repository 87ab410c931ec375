#!/bin/bash
# First GPU trip: run every parity suite under its own timeout so a hung kernel cannot eat the box.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" | tee -a gpurun_out/summary.txt; tail -n 6 gpurun_out/$name.log | tee -a gpurun_out/summary.txt; }
run geom 400 python -m pytest tests/test_gpu_geom.py -x -q
run kernels_noconv 300 python -m pytest tests/test_gpu_kernels.py -q -k "not conv_engine"
run conv_simt 300 python -m pytest tests/test_gpu_kernels.py -q -k "conv_engine and 0-3-"
run conv_tc3 300 python -m pytest tests/test_gpu_kernels.py -q -k "conv_engine and 1-3-"
run conv_tc1 300 python -m pytest tests/test_gpu_kernels.py -q -k "conv_engine and 1-1-"
run net 600 python -m pytest tests/test_gpu_net.py -q
run timing 600 python tools/time_forward.py 8 64
