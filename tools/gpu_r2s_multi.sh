#!/bin/bash
# N-GPU pass: bench (both arms) under torchrun, with and without the NCCL-registered gradient buffer
TAG=${1:-r2s}
N=${2:-2}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29521 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_${N}gpu_$TAG.json 2> $O/bench_${N}gpu_$TAG.err; echo "bench$N exit $?"; cut -c1-300 $O/bench_${N}gpu_$TAG.json; tail -3 $O/bench_${N}gpu_$TAG.err
python - <<PY
import json
try:
    d=json.load(open("$O/bench_${N}gpu_$TAG.json")); print("scaling_breakdown", json.dumps(d.get("scaling_breakdown"))[:900])
except Exception as e: print("parse failed", e)
PY
BDN_NCCL_REGISTER=1 NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,REG timeout 900 $TR --master-port 29524 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_${N}gpu_reg_$TAG.json 2> $O/bench_${N}gpu_reg_$TAG.err; echo "bench$N registered exit $?"; cut -c1-200 $O/bench_${N}gpu_reg_$TAG.json
grep -i -m5 "nvls\|register" $O/bench_${N}gpu_reg_$TAG.err | cut -c1-200
grep -v "NCCL INFO" $O/bench_${N}gpu_reg_$TAG.err | tail -3
grep -i "NVLS\|Registered\|regist" $O/bench_${N}gpu_reg_$TAG.err | head -40 > $O/nccl_reg_${N}gpu_$TAG.txt
timeout 600 $TR --master-port 29522 bench.py --gpus $N --steps 5 --warmup 2 --impl reference > $O/bench_${N}gpu_ref_$TAG.json 2> $O/err2.log; echo "ref$N exit $?"; cut -c1-200 $O/bench_${N}gpu_ref_$TAG.json
