#!/bin/bash
# 2-GPU pass: parity tests on GPU 0, then the bench under torchrun (split backward + overlapped all-reduce)
TAG=${1:-x}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -5 $O/pytest_$TAG.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > $O/bench_2gpu_$TAG.json 2> $O/bench_2gpu_$TAG.err; echo "bench2 exit $?"
tail -c 400 $O/bench_2gpu_$TAG.err; cut -c1-700 $O/bench_2gpu_$TAG.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 1 --impl reference > $O/bench_2gpu_ref_$TAG.json 2> $O/bench_2gpu_ref_$TAG.err; echo "ref2 exit $?"; cut -c1-300 $O/bench_2gpu_ref_$TAG.json
