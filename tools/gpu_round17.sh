#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-8} gpurun_out/$name.log | cut -c1-300 | tee -a gpurun_out/summary.txt; }
TAILN=12 run conv 900 python -m pytest tests/test_gpu_kernels.py -q -k "conv_engine and (case15 or case17 or case18)" 
TAILN=12 SUO_RAW_TMA=0 run conv_noraw 900 python -m pytest tests/test_gpu_kernels.py -q -k "conv_engine and (case15 or case17 or case18)" 
TAILN=40 run bisect 600 python tools/bisect_raw.py
