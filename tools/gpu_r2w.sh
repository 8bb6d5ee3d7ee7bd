#!/bin/bash
# winv tiling knob sweep for the per-snapshot net (lines x channels-per-item), default bench
TAG=${1:-r2w}
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 50 --warmup 5 --top 40 --no-cpu-baseline"
run() { n=$1; shift; env "$@" timeout 600 $B > $O/bench_${TAG}_$n.json 2> $O/err.log
  python - <<PY
import json
d=json.load(open("$O/bench_${TAG}_$n.json"))
w=" ".join("%s=%.1f"%(k["kernel"],k["us_per_step"]/k["launches_per_step"]) for k in d["top_kernels"] if k["kernel"] in ("winv_layer_bwd/4","winv_layer_fwd/4","mse_heads"))
print("%-10s value %.1f samples/s  %.3f ms/step | %s"%("$n",d["value"],d["ms_per_step"],w))
PY
}
run base A=1
run t6_2 BDN_WINV_TILE=6,2
run t7_2 BDN_WINV_TILE=7,2
run t3_4 BDN_WINV_TILE=3,4
run t6_4 BDN_WINV_TILE=6,4
run t13_2 BDN_WINV_TILE=13,2
run t4_1 BDN_WINV_TILE=4,1
