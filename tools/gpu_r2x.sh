#!/bin/bash
TAG=${1:-r2x}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
B="python bench.py --steps 50 --warmup 5 --top 40 --no-cpu-baseline"
run() { n=$1; shift; env "$@" timeout 600 $B > $O/bench_${TAG}_$n.json 2> $O/err.log
  python - <<PY
import json
d=json.load(open("$O/bench_${TAG}_$n.json"))
w=" ".join("%s=%.1f"%(k["kernel"],k["us_per_step"]/k["launches_per_step"]) for k in d["top_kernels"] if k["kernel"] in ("lift/4","project_bwd/12","mse_heads"))
print("%-10s value %.1f samples/s  %.3f ms/step | %s"%("$n",d["value"],d["ms_per_step"],w))
PY
}
run base A=1
run oldlift BDN_LIFT_BAGS4=0
run pp2 BDN_PROJ_BWD_PP8=2
run pp2_cap592 BDN_PROJ_BWD_PP8=2 BDN_PROJ_BWD_CAP8=592
timeout 300 python tools/timeline.py --out $O/timeline_$TAG.json > $O/timeline_$TAG.log 2>&1; tail -3 $O/timeline_$TAG.log
BDN_PROJ_BWD_PP8=2 timeout 300 python tools/timeline.py --out $O/timeline_${TAG}_pp2.json > $O/timeline_$TAG.log 2>&1
python - <<PY
import json
for t in ("$TAG","${TAG}_pp2"):
    d=json.load(open("$O/timeline_%s.json"%t)); print(t, round(d["span_us_per_step"],1), {k.split("bdn::")[-1][:30]:round(v[1]/v[0],1) for k,v in d["by_kernel_us_per_step"].items() if "lift" in k or "project_bwd_kernel<12" in k or "mse" in k})
PY
