#!/bin/bash
TAG=${1:-r2ba}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_custom_ops.py -x -q -m gpu -k "bag_attention" 2>&1 | tail -15
timeout 600 python -m pytest tests -x -q -m gpu -k "blindno" 2>&1 | tail -8
timeout 600 python bench.py --workload blindno_2d --steps 30 --warmup 5 --top 60 --no-cpu-baseline > $O/bench_blindno_$TAG.json 2> $O/err.log; python -c "
import json;d=json.load(open('$O/bench_blindno_$TAG.json'));print('blindno',round(d['value'],1),round(d['ms_per_step'],3));print([(k['kernel'],round(k['us_per_step'],1)) for k in d['top_kernels'] if k['kernel'].startswith('bagattn')])"
