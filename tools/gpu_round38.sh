#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-8} gpurun_out/$name.log | cut -c1-3000 | tee -a gpurun_out/summary.txt; }
rm -f gpurun_out/summary.txt
TAILN=40 run raggedtests 300 python -m pytest tests/test_gpu_net.py -q -x -k "ragged"
