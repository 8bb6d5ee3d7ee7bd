#!/bin/bash
# programmatic dependent launch with the trigger late in every kernel (library rebuilt on the box with -DBDN_PDL_LATE=1)
TAG=${1:-r2y}
O=gpurun_out
mkdir -p $O
B0="python bench.py --steps 50 --warmup 5 --top 40 --no-cpu-baseline"
for v in "A=1" "BDN_PROJ_BWD_PP8=4" "BDN_MSE_PIX_PER_BLOCK=100000" "BDN_MSE_PIX_PER_BLOCK=4096"; do env $v timeout 600 $B0 > $O/bench_${TAG}_k.json 2>$O/err.log; python -c "import json;d=json.load(open('$O/bench_${TAG}_k.json'));print('$v',round(d['value'],1),round(d['ms_per_step'],4),[(k['kernel'],round(k['us_per_step']/k['launches_per_step'],1)) for k in d['top_kernels'] if k['kernel'] in ('mse_heads','project_bwd/12')])"; done
timeout 300 python tools/timeline.py --out $O/timeline_${TAG}.json > $O/timeline_$TAG.log 2>&1; python -c "import json;d=json.load(open('$O/timeline_${TAG}.json'));print('timeline',round(d['span_us_per_step'],1),{k.split('bdn::')[-1][:28]:round(v[1]/v[0],1) for k,v in d['by_kernel_us_per_step'].items() if 'mse' in k or 'project_bwd_kernel<12' in k})"
export BDN_NVCC_EXTRA="-DBDN_PDL_LATE=1"
python -c "from blindno_b200 import build; print(build.build())" 2>&1 | tail -1
B="python bench.py --steps 50 --warmup 5 --top 40 --no-cpu-baseline"
run() { n=$1; shift; env "$@" timeout 600 $B > $O/bench_${TAG}_$n.json 2> $O/err_$n.log || tail -3 $O/err_$n.log
  python - <<PY
import json
d=json.load(open("$O/bench_${TAG}_$n.json"))
print("%-10s value %.1f samples/s  %.3f ms/step  e2e %.1f  loss %s"%("$n",d["value"],d["ms_per_step"],d["e2e"]["value"],d.get("final_loss")))
PY
}
run pdl0 BDN_PDL=0
run pdl1 BDN_PDL=1
run pdl2 BDN_PDL=2
for v in 0 1; do BDN_PDL=$v HEADS_TAG="late trigger, BDN_PDL=$v" timeout 200 python tools/heads_bench.py 300 2>&1 | tail -1; done | tee $O/heads_bench_$TAG.jsonl
BDN_PDL=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "default_shape or graph_replay or golden" 2>&1 | tail -3
BDN_PDL=1 timeout 300 python tools/timeline.py --out $O/timeline_${TAG}_pdl1.json > $O/timeline_$TAG.log 2>&1; tail -3 $O/timeline_$TAG.log
