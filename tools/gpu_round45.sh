#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-300 | tee -a gpurun_out/summary.txt; }
rm -f gpurun_out/summary.txt
run alltests 400 python -m pytest tests -q -m gpu -x
run smoke 200 python -c "import __graft_entry__ as g; g.smoke()"
SUO_BENCH_PER_OP=gpurun_out/per_op.csv run bench 600 python bench.py
python - <<'PY'
import json
for l in open("gpurun_out/bench.log"):
    if l.startswith("{"):
        d = json.loads(l); print("value", d["value"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"]); print("pose_err", d.get("pose_err"))
PY
