#!/bin/bash
TAG=${1:-r2o}
O=gpurun_out
mkdir -p $O
export BDN_PROJ_BWD_CAP8=296
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
B="python bench.py --steps 50 --warmup 5 --top 40 --no-cpu-baseline"
timeout 600 $B > $O/bench_$TAG.json 2> $O/bench_$TAG.err; tail -c 300 $O/bench_$TAG.err
python - <<PY
import json
d=json.load(open("$O/bench_$TAG.json"))
print("$TAG value %.1f samples/s  %.3f ms/step  e2e %.1f launches %d"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["gpu_launches"]))
PY
timeout 300 python tools/timeline.py --out $O/timeline_$TAG.json > $O/timeline_$TAG.log 2>&1; tail -3 $O/timeline_$TAG.log
