#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-8} gpurun_out/$name.log | cut -c1-2500 | tee -a gpurun_out/summary.txt; }
rm -f gpurun_out/summary.txt
run alltests 1500 python -m pytest tests -q -m gpu
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
