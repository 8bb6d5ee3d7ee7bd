#!/bin/bash
# 8-GPU pass: bench (both arms) under torchrun, a reduced configs[4] sweep on 8 GPUs
TAG=${1:-x}
N=${2:-8}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus $N --steps 50 --warmup 5 > $O/bench_${N}gpu_$TAG.json 2> $O/bench_${N}gpu_$TAG.err; echo "bench$N exit $?"; cut -c1-330 $O/bench_${N}gpu_$TAG.json
timeout 600 $TR --master-port 29522 bench.py --gpus $N --steps 3 --warmup 1 --impl reference > $O/bench_${N}gpu_ref_$TAG.json 2> $O/err2.log; echo "ref$N exit $?"; cut -c1-200 $O/bench_${N}gpu_ref_$TAG.json
timeout 900 $TR --master-port 29523 tools/sweep.py --graphs --steps 4 --warmup 2 --bags 100,400 --grids 61,128,256 --modes 12,32,64 > $O/sweep_${N}gpu_$TAG.jsonl 2> $O/sweep_${N}gpu_$TAG.err; echo "sweep$N exit $?"; cut -c1-250 $O/sweep_${N}gpu_$TAG.jsonl
