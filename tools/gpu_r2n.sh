#!/bin/bash
# knob matrix on the default bench: weights staged / single table in the heads' core2d, PDL modes
TAG=${1:-r2n}
O=gpurun_out
mkdir -p $O
export BDN_PROJ_BWD_CAP8=296
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
B="python bench.py --steps 40 --warmup 5 --top 40 --no-cpu-baseline"
run() { # name env...
  n=$1; shift
  env "$@" timeout 600 $B > $O/bench_${TAG}_$n.json 2> $O/err.log
  python - <<PY
import json
d=json.load(open("$O/bench_${TAG}_$n.json"))
print("%-22s value %.1f samples/s  %.3f ms/step  e2e %.1f"%("$n",d["value"],d["ms_per_step"],d["e2e"]["value"]))
PY
}
run base BDN_PDL=0
run nowstage BDN_CORE_WSTAGE=0
run wstage_twotab BDN_CORE_ONETAB=0
run pdl1 BDN_PDL=1
run pdl2 BDN_PDL=2
run pdl1_nowstage BDN_PDL=1 BDN_CORE_WSTAGE=0
for v in "BDN_PDL=0" "BDN_PDL=0 BDN_CORE_WSTAGE=0" "BDN_PDL=1"; do env $v HEADS_TAG="$v" timeout 200 python tools/heads_bench.py 300 2>&1 | tail -1; done | tee $O/heads_bench_$TAG.jsonl
timeout 300 python tools/timeline.py --out $O/timeline_$TAG.json > $O/timeline_$TAG.log 2>&1; tail -3 $O/timeline_$TAG.log
