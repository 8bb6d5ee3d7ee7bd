#!/bin/bash
# A/B of library tuning knobs in ONE gpurun call (same box, same clocks): every argument is an environment setting
# ("BDN_PDL=1", "BDN_CORE_WSTAGE=0 BDN_CORE_ONETAB=0", "A=1" for the default) the default bench is run under.
#   gpurun -- 'bash tools/gpu_ab.sh "A=1" "BDN_WFWD_FOLD=0" "BDN_PROJ_BWD_PP8=1"'
# Knobs (csrc/*.cu, read once per process): BDN_PDL (0/1/2), BDN_CORE_WSTAGE, BDN_CORE_ONETAB, BDN_WFWD_FOLD (0/1/2),
# BDN_WFWD_TC_AUTO (0/1/2), BDN_GW_TILED, BDN_LIFT_BAGS4, BDN_PROJ_BWD_CAP8, BDN_PROJ_BWD_PP8 (1/2/4),
# BDN_WINV_TILE ("lines,channels"), BDN_MSE_PIX_PER_BLOCK, BDN_PROJ_FWD_PP (2/4/8), BDN_PROJ_FWD_CAP (blocks); build-time: BDN_NVCC_EXTRA="-DBDN_PDL_LATE=1".
O=gpurun_out
mkdir -p $O
EXTRA=${BENCH_ARGS:-"--steps 50 --warmup 5 --top 40 --no-cpu-baseline"}
i=0
for v in "$@"; do
  i=$((i+1))
  env $v timeout 600 python bench.py $EXTRA > $O/bench_ab_$i.json 2> $O/err_ab_$i.log || tail -3 $O/err_ab_$i.log
  python - "$v" $O/bench_ab_$i.json <<'PY'
import json, sys
d = json.load(open(sys.argv[2]))
top = " ".join("%s=%.1f" % (k["kernel"], k["us_per_step"] / k["launches_per_step"]) for k in d["top_kernels"][:12])
print("%-40s %.1f samples/s  %.4f ms/step  e2e %.1f | %s" % (sys.argv[1], d["value"], d["ms_per_step"], d["e2e"]["value"], top))
PY
done
