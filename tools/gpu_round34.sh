#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-8} gpurun_out/$name.log | cut -c1-3000 | tee -a gpurun_out/summary.txt; }
rm -f gpurun_out/summary.txt
SUO_FUSE=2 TAILN=70 run timeline 300 python tools/fused_timeline.py gpurun_out/fused2_timeline.csv
