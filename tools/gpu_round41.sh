#!/bin/bash
# final round-1 evidence run: tests, smoke, bench (both arms), ncu launch list + one full capture of the dominant kernels
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-700 | tee -a gpurun_out/summary.txt; }
rm -f gpurun_out/summary.txt gpurun_out/trace.*.csv
run alltests 1200 python -m pytest tests -q -m gpu -x
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
SUO_BENCH_PER_OP=gpurun_out/per_op.csv run bench 900 python bench.py

run bench_ref 600 python bench.py --impl reference --steps 3 --warmup 1
run ncu_list 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 636 -c 215 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline



ls -la gpurun_out >> gpurun_out/summary.txt
