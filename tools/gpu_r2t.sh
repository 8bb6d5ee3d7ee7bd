#!/bin/bash
TAG=${1:-r2t}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
B="python bench.py --steps 50 --warmup 5 --top 40 --no-cpu-baseline"
run() { n=$1; shift; env "$@" timeout 600 $B > $O/bench_${TAG}_$n.json 2> $O/err.log
  python - <<PY
import json
d=json.load(open("$O/bench_${TAG}_$n.json"))
print("%-14s value %.1f samples/s  %.3f ms/step  e2e %.1f"%("$n",d["value"],d["ms_per_step"],d["e2e"]["value"]))
for k in d["top_kernels"]:
    if any(k["kernel"].startswith(t) for t in ("gw_reduce","mse","lift/","winv")): print("     %-20s %5.1f x %7.1f us/step"%(k["kernel"],k["launches_per_step"],k["us_per_step"]))
PY
}
run new BDN_GW_TILED=1
run oldgw BDN_GW_TILED=0
for g in 1 0; do BDN_GW_TILED=$g timeout 600 python bench.py --steps 12 --warmup 4 --no-cpu-baseline --batch-per-gpu 32 > $O/bench_${TAG}_b32_gw$g.json 2>$O/err.log; python -c "import json;d=json.load(open('$O/bench_${TAG}_b32_gw$g.json'));print('b32 gw_tiled=$g',d['value'],d['ms_per_step'])"; done
timeout 300 python tools/timeline.py --out $O/timeline_$TAG.json > $O/timeline_$TAG.log 2>&1; tail -3 $O/timeline_$TAG.log
