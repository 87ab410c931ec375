#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-8} gpurun_out/$name.log | cut -c1-2500 | tee -a gpurun_out/summary.txt; }
rm -f gpurun_out/summary.txt
TAILN=40 run newtests 900 python -m pytest tests -q -m gpu -k "prior or global or large or malformed or g2o or curr_only or joint"
