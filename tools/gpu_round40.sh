#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-8} gpurun_out/$name.log | cut -c1-700 | tee -a gpurun_out/summary.txt; }
rm -f gpurun_out/summary.txt gpurun_out/trace.*.csv
TAILN=25 run nettests 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_net.py -q -x
run bench_pdl 600 python bench.py --no-cpu-baseline
SUO_PDL=0 run bench_nopdl 600 python bench.py --no-cpu-baseline
run bench_pdl2 600 python bench.py --no-cpu-baseline
