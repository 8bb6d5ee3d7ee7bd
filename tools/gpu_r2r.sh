#!/bin/bash
TAG=${1:-r2r}
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
    if k["kernel"].startswith("wfwd") or k["kernel"].startswith("mse"): print("     %-20s %5.1f x %7.1f us/step"%(k["kernel"],k["launches_per_step"],k["us_per_step"]))
PY
}
run fold BDN_WFWD_FOLD=1
run nofold BDN_WFWD_FOLD=0
BDN_WFWD_FOLD=1 timeout 600 python bench.py --steps 12 --warmup 4 --no-cpu-baseline --batch-per-gpu 32 > $O/bench_${TAG}_b32_fold.json 2>$O/err.log; python -c "import json;d=json.load(open('$O/bench_${TAG}_b32_fold.json'));print('b32 fold',d['value'],d['ms_per_step'])"
BDN_WFWD_FOLD=0 timeout 600 python bench.py --steps 12 --warmup 4 --no-cpu-baseline --batch-per-gpu 32 > $O/bench_${TAG}_b32_nofold.json 2>$O/err.log; python -c "import json;d=json.load(open('$O/bench_${TAG}_b32_nofold.json'));print('b32 nofold',d['value'],d['ms_per_step'])"
timeout 300 python tools/timeline.py --out $O/timeline_$TAG.json > $O/timeline_$TAG.log 2>&1; tail -3 $O/timeline_$TAG.log
