#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-8} gpurun_out/$name.log | cut -c1-3000 | tee -a gpurun_out/summary.txt; }
rm -f gpurun_out/summary.txt
run probe 400 ncu --set full --clock-control none --import-source on -k regex:conv -f -o gpurun_out/pair_probe python tools/pair_probe.py
run bench64 600 python bench.py --no-cpu-baseline --frames-per-step 64
TAILN=5 run alltests 900 python -m pytest tests -m gpu -q -x
