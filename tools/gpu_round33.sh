#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-8} gpurun_out/$name.log | cut -c1-3000 | tee -a gpurun_out/summary.txt; }
rm -f gpurun_out/summary.txt
TAILN=30 run fusetest 200 python -m pytest tests/test_gpu_net.py -q -x -k "fused_bottleneck"
if grep -q "passed" gpurun_out/fusetest.log && ! grep -q "failed" gpurun_out/fusetest.log; then
SUO_FUSE=2 run bench_fuse2 600 python bench.py --no-cpu-baseline
SUO_FUSE=2 TAILN=40 run timeline 300 python tools/fused_timeline.py gpurun_out/fused2_timeline.csv
fi
