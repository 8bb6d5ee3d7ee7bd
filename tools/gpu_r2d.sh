#!/bin/bash
TAG=${1:-r2d}
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 50 --warmup 5 --top 40 --no-cpu-baseline"
for v in 0 1 0 1; do BDN_WFWD_TC_AUTO=$v timeout 600 $B > $O/bench_${TAG}_tc$v.json 2> $O/err.log; python -c "
import json;d=json.load(open('$O/bench_${TAG}_tc$v.json'));print('tc_auto=$v',round(d['value'],1),round(d['ms_per_step'],4),[(k['kernel'],round(k['us_per_step']/k['launches_per_step'],1)) for k in d['top_kernels'] if k['kernel'].startswith('wfwd')])"; done
BDN_WFWD_TC_AUTO=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "default_shape or golden or graph_replay or end_metric" 2>&1 | tail -3
BDN_WFWD_TC_AUTO=1 timeout 600 python bench.py --steps 12 --warmup 4 --no-cpu-baseline --batch-per-gpu 32 > $O/bench_${TAG}_b32_tc1.json 2>$O/err.log; python -c "import json;d=json.load(open('$O/bench_${TAG}_b32_tc1.json'));print('b32 tc_auto=1',round(d['value'],1),round(d['ms_per_step'],3))"
