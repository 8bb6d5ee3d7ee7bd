#!/bin/bash
TAG=${1:-r2l}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
B="python bench.py --steps 40 --warmup 5 --top 40 --no-cpu-baseline"
BDN_PROJ_BWD_CAP8=296 timeout 600 $B > $O/bench_$TAG.json 2> $O/bench_$TAG.err; tail -c 400 $O/bench_$TAG.err
python - <<PY
import json
for t in ("$TAG",):
    d=json.load(open("$O/bench_%s.json"%t))
    print(t,"value %.1f samples/s  %.3f ms/step  e2e %.1f  kernel-sum %.0f us/step launches %d"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["kernel_time_us_per_step"],d["gpu_launches"]))
PY
BDN_PROJ_BWD_CAP8=296 timeout 300 python tools/timeline.py --out $O/timeline_$TAG.json > $O/timeline_$TAG.log 2>&1; tail -3 $O/timeline_$TAG.log
bash tools/gpu_heads_src.sh $TAG
