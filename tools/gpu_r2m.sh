#!/bin/bash
TAG=${1:-r2m}
O=gpurun_out
mkdir -p $O
export BDN_PROJ_BWD_CAP8=296
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for pdl in 0 1; do BDN_PDL=$pdl timeout 200 python tools/heads_bench.py 300 2>&1 | tail -1; done | tee $O/heads_bench_$TAG.jsonl
B="python bench.py --steps 40 --warmup 5 --top 40 --no-cpu-baseline"
timeout 600 $B > $O/bench_$TAG.json 2> $O/bench_$TAG.err; tail -c 400 $O/bench_$TAG.err
BDN_PDL=1 timeout 600 $B > $O/bench_${TAG}_pdl.json 2> $O/err.log
python - <<PY
import json
for t in ("$TAG","${TAG}_pdl"):
    d=json.load(open("$O/bench_%s.json"%t))
    print(t,"value %.1f samples/s  %.3f ms/step  e2e %.1f  kernel-sum %.0f us/step launches %d"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["kernel_time_us_per_step"],d["gpu_launches"]))
PY
timeout 300 python tools/timeline.py --out $O/timeline_$TAG.json > $O/timeline_$TAG.log 2>&1; tail -3 $O/timeline_$TAG.log
bash tools/gpu_heads_src.sh $TAG > $O/heads_src_$TAG.log 2>&1; head -32 $O/heads_${TAG}_summary.txt
rm -f $O/heads_$TAG.ncu-rep
