#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-8} gpurun_out/$name.log | cut -c1-3000 | tee -a gpurun_out/summary.txt; }
rm -f gpurun_out/summary.txt
TAILN=30 run pairtest 240 python -m pytest tests/test_gpu_kernels.py -q -x -k "pair"
if grep -q "passed" gpurun_out/pairtest.log && ! grep -q "failed" gpurun_out/pairtest.log; then
TAILN=15 run pairnet 300 python -m pytest tests/test_gpu_net.py -q -x -k "cta_pair"
run bench 600 python bench.py --no-cpu-baseline
SUO_PAIR=0 run bench_nopair 600 python bench.py --no-cpu-baseline
fi
