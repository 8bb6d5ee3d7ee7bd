#!/bin/bash
# N-GPU bench, default arm (short) -- validates the multi-rank path end to end
TAG=${1:-r2u}
N=${2:-2}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_${N}gpu_$TAG.json 2> $O/bench_${N}gpu_$TAG.err; echo "bench$N exit $?"; cut -c1-300 $O/bench_${N}gpu_$TAG.json; grep -v "OMP_NUM\|\*\*\*\*\|^$" $O/bench_${N}gpu_$TAG.err | tail -5
python - <<PY
import json
try:
    d=json.load(open("$O/bench_${N}gpu_$TAG.json")); print("value", d["value"], "ms", d["ms_per_step"]); print("scaling_breakdown", json.dumps(d.get("scaling_breakdown"))[:900])
except Exception as e: print("parse failed", e)
PY
