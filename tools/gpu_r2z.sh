#!/bin/bash
TAG=${1:-r2z}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
B="python bench.py --steps 50 --warmup 5 --top 40 --no-cpu-baseline"
timeout 600 $B > $O/bench_$TAG.json 2> $O/err.log; python -c "import json;d=json.load(open('$O/bench_$TAG.json'));print('default',round(d['value'],1),round(d['ms_per_step'],4),round(d['e2e']['value'],1),[(k['kernel'],round(k['us_per_step']/k['launches_per_step'],1)) for k in d['top_kernels'] if k['kernel'] in ('project/4','project/12','wfwd_gelu/12','winv_layer_fwd/4')])"
timeout 600 python bench.py --steps 12 --warmup 4 --top 40 --no-cpu-baseline --batch-per-gpu 32 > $O/bench_b32_$TAG.json 2> $O/err.log; python -c "import json;d=json.load(open('$O/bench_b32_$TAG.json'));print('b32',round(d['value'],1),round(d['ms_per_step'],4),[(k['kernel'],round(k['us_per_step']/k['launches_per_step'],1)) for k in d['top_kernels'] if k['kernel'].startswith('core2d')])"
timeout 200 python tools/heads_bench.py 300 2>&1 | tail -1
timeout 300 python tools/timeline.py --out $O/timeline_$TAG.json > $O/timeline_$TAG.log 2>&1; python -c "import json;d=json.load(open('$O/timeline_$TAG.json'));print('timeline',{k:round(v,1) for k,v in d.items() if k!='by_kernel_us_per_step'})"
