#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" | tee -a gpurun_out/summary.txt; tail -n 8 gpurun_out/$name.log | tee -a gpurun_out/summary.txt; }
run timeline 300 python tools/conv_timeline.py gpurun_out/timeline.csv 64
# DRAM traffic of every launch of one timed step (cheap metrics: one replay pass)
run ncu_traffic 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 636 -c 215 --csv --log-file gpurun_out/traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline
ls -la gpurun_out >> gpurun_out/summary.txt
