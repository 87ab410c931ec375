#!/bin/bash
# GPU trip 2: parity suites for both tcgen05 kernel variants, smoke, bench, ncu launch list + full capture.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" | tee -a gpurun_out/summary.txt; tail -n 8 gpurun_out/$name.log | tee -a gpurun_out/summary.txt; }
run geom 400 python -m pytest tests/test_gpu_geom.py -q
run conv_persistent 300 python -m pytest tests/test_gpu_kernels.py -q -k "conv_engine"
SUO_CONV_PERSISTENT=0 run conv_onetile 300 python -m pytest tests/test_gpu_kernels.py -q -k "conv_engine and 1-3-"
run net 600 python -m pytest tests/test_gpu_net.py -q -s
run timing_persistent 300 python tools/time_forward.py 8 64
SUO_CONV_PERSISTENT=0 run timing_onetile 300 python tools/time_forward.py 64
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
run bench 900 python bench.py --steps 10 --warmup 3
run bench_ref 600 python bench.py --impl reference --steps 3 --warmup 1
run ncu_list 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 450 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline
run ncu_full 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 320 -c 8 -o gpurun_out/prof_conv python bench.py --steps 1 --warmup 3 --no-cpu-baseline
ls -la gpurun_out >> gpurun_out/summary.txt
