#!/bin/bash
# GPU status-quo comparator (the reference's algorithm on stock PyTorch CUDA kernels) next to this repo's arm, per workload
TAG=${1:-x}
O=gpurun_out
mkdir -p $O
for wl in 2d_FPE 2d_NC 1d_FPE; do
  timeout 600 python bench.py --impl reference --ref-device cuda --workload $wl --steps 20 --warmup 3 > $O/bench_refcuda_${wl}_$TAG.json 2> $O/err.log; echo "refcuda $wl exit $?"; cut -c1-200 $O/bench_refcuda_${wl}_$TAG.json
  timeout 600 python bench.py --workload $wl --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_${wl}_$TAG.json 2> $O/err.log; echo "ours $wl exit $?"; cut -c1-200 $O/bench_${wl}_$TAG.json
done
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-graphs > $O/bench_2d_FPE_eager_$TAG.json 2> $O/err.log; echo "eager exit $?"; cut -c1-200 $O/bench_2d_FPE_eager_$TAG.json
python -c "import __graft_entry__ as g; g.smoke()"
