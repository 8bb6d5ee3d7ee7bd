#!/bin/bash
TAG=${1:-r2c}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --workload blindno_2d --steps 30 --warmup 5 --top 60 --no-cpu-baseline > $O/bench_blindno_$TAG.json 2> $O/err.log; python -c "
import json;d=json.load(open('$O/bench_blindno_$TAG.json'));print('blindno',round(d['value'],1),round(d['ms_per_step'],3));print([(k['kernel'],round(k['us_per_step'],1)) for k in d['top_kernels'] if k['kernel'].startswith('bagattn')])"
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_$TAG.json 2> $O/err.log; python -c "import json;d=json.load(open('$O/bench_$TAG.json'));print('default',round(d['value'],1),round(d['ms_per_step'],4),round(d['e2e']['value'],1))"
