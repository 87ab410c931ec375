#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" | tee -a gpurun_out/summary.txt; tail -n 8 gpurun_out/$name.log | tee -a gpurun_out/summary.txt; }
run alltests 1200 python -m pytest tests -q -m gpu -x
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
SUO_PROFILE_DUMP=gpurun_out/per_op.csv run bench 900 python bench.py
run bench_ref 600 python bench.py --impl reference --steps 5 --warmup 2
ls -la gpurun_out >> gpurun_out/summary.txt
