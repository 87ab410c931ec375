#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-8} gpurun_out/$name.log | cut -c1-600 | tee -a gpurun_out/summary.txt; }
TAILN=40 run conv 900 python -m pytest tests/test_gpu_kernels.py -q -k "conv_engine" 
SUO_RAW_TMA=0 run net_noraw 600 python -m pytest tests/test_gpu_net.py -q -x -k "golden and net_small-2"
ls -la gpurun_out >> gpurun_out/summary.txt
