#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" | tee -a gpurun_out/summary.txt; tail -n 8 gpurun_out/$name.log | tee -a gpurun_out/summary.txt; }
run geom 400 python -m pytest tests/test_gpu_geom.py -q -s
run net512 400 python -m pytest tests/test_gpu_net.py -q -s -k "tless or 256"
SUO_PROFILE_DUMP=gpurun_out/per_op.csv run bench 900 python bench.py --steps 10 --warmup 3
run bench_ref 600 python bench.py --impl reference --steps 3 --warmup 1
run ncu_list 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 450 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline
run ncu_full 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 320 -c 8 -o gpurun_out/prof_conv python bench.py --steps 1 --warmup 3 --no-cpu-baseline
ls -la gpurun_out >> gpurun_out/summary.txt
