#!/bin/bash
# GPU pass 3: FFMA/FFMA2/MUFU issue-rate microbenchmark, ncu --set full over one eager step
TAG=${1:-x}
O=gpurun_out
mkdir -p $O
./tools/ubench_ffma2.bin | tee $O/ubench_ffma2_$TAG.json
B="python bench.py --steps 1 --warmup 1 --no-graphs --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --csv --page raw --log-file $O/full_$TAG.csv -k regex:_kernel -s 90 -c 90 $B > $O/ncu_full.log 2>&1; echo "ncu full exit $?"
python tools/ncucsv.py $O/full_$TAG.csv --json $O/full_$TAG.json > $O/full_${TAG}_summary.txt 2>&1; head -60 $O/full_${TAG}_summary.txt
