#!/bin/bash
# parity tests, bench (fp32), the same with the few-pixel projection-backward grid capped at 2 / 1 blocks per SM, timeline
TAG=${1:-r2k}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
B="python bench.py --steps 40 --warmup 5 --top 40 --no-cpu-baseline"
timeout 600 $B > $O/bench_$TAG.json 2> $O/bench_$TAG.err; tail -c 400 $O/bench_$TAG.err
BDN_PROJ_BWD_CAP8=296 timeout 600 $B > $O/bench_${TAG}_cap296.json 2> $O/err.log
BDN_PROJ_BWD_CAP8=148 timeout 600 $B > $O/bench_${TAG}_cap148.json 2> $O/err.log
python - <<PY
import json
for t in ("$TAG","${TAG}_cap296","${TAG}_cap148"):
    d=json.load(open("$O/bench_%s.json"%t))
    print(t,"value %.1f samples/s  %.3f ms/step  e2e %.1f  kernel-sum %.0f us/step launches %d"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["kernel_time_us_per_step"],d["gpu_launches"]))
    for k in d["top_kernels"]:
        if t=="$TAG" or "project" in k["kernel"]: print("   %-22s %5.1f x %8.1f us/step  %5.1f%%"%(k["kernel"],k["launches_per_step"],k["us_per_step"],100*k["share"]))
PY
timeout 300 python tools/timeline.py --out $O/timeline_$TAG.json > $O/timeline_$TAG.log 2>&1; tail -3 $O/timeline_$TAG.log
