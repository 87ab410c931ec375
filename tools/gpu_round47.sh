#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-400 | tee -a gpurun_out/summary.txt; }
rm -f gpurun_out/summary.txt
TAILN=30 run halotest 100 python -m pytest tests/test_gpu_kernels.py -q -x -k "halo"
if grep -q "passed" gpurun_out/halotest.log && ! grep -q "failed" gpurun_out/halotest.log; then
TAILN=30 run halonet 120 python -m pytest tests/test_gpu_net.py -q -x -k "halo"
SUO_HALO=1 run bench_halo 200 python bench.py --no-cpu-baseline
run bench_nohalo 200 python bench.py --no-cpu-baseline
python - <<'PY'
import json
for f in ("bench_halo", "bench_nohalo"):
    for l in open(f"gpurun_out/{f}.log"):
        if l.startswith("{"):
            d = json.loads(l); print(f, d["value"], {k: v["ms"] for k, v in d["conv_engine"]["classes"].items() if "3x3" in k})
PY
fi
