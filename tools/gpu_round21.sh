#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-400 | tee -a gpurun_out/summary.txt; }
SUO_TRACE=gpurun_out/trace SUO_GRID_CAP=74 run s2_74 300 python tools/dual_stream.py 2 32
