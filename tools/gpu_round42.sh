#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-500 | tee -a gpurun_out/summary.txt; }
rm -f gpurun_out/summary.txt
run alltests 1200 python -m pytest tests -q -m gpu -x
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
run bench 600 python bench.py --no-cpu-baseline
