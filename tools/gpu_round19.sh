#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-8} gpurun_out/$name.log | cut -c1-400 | tee -a gpurun_out/summary.txt; }
for f in 1 2 4; do
SUO_PROFILE_DUMP=gpurun_out/per_op_f$f.csv run bench_f$f 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --frames-per-step $f
done
run ncu_list 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 636 -c 215 --csv --log-file gpurun_out/traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline
