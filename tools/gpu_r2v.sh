#!/bin/bash
TAG=${1:-r2v}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 900 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("$O/bench_$TAG.json"))
print("value %.1f samples/s  %.3f ms/step  e2e %.1f launches %d"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["gpu_launches"]))
print(d["roofline"]["kernel"], d["roofline"]["frac"], d["cpu_baseline"])
PY
timeout 300 python tools/timeline.py --out $O/timeline_$TAG.json > $O/timeline_$TAG.log 2>&1; tail -3 $O/timeline_$TAG.log
