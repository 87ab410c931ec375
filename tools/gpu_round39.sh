#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-8} gpurun_out/$name.log | cut -c1-600 | tee -a gpurun_out/summary.txt; }
rm -f gpurun_out/summary.txt gpurun_out/trace.*.csv
SUO_TRACE=gpurun_out/trace run bench_trace 600 python bench.py --no-cpu-baseline --steps 20
for f in gpurun_out/trace.*.csv; do python tools/trace_gaps.py $f | tee -a gpurun_out/summary.txt; done
