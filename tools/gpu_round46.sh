#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-300 | tee -a gpurun_out/summary.txt; }
rm -f gpurun_out/summary.txt
TAILN=25 run nettests 240 python -m pytest tests/test_gpu_net.py -q -x
run bench_stem 300 python bench.py --no-cpu-baseline
SUO_STEM_TMA=0 run bench_nostem 300 python bench.py --no-cpu-baseline
python - <<'PY'
import json
for f in ("bench_stem", "bench_nostem"):
    for l in open(f"gpurun_out/{f}.log"):
        if l.startswith("{"):
            d = json.loads(l); print(f, d["value"], d["conv_engine"]["classes"].get("7x7/2 stem"))
PY
