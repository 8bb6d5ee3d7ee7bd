#!/bin/bash
# ncu --set full with sources over the kernels of one output head (eager, third iteration): per-kernel table + SASS-level
# stall samples of the latency-bound few-image kernels
TAG=${1:-x}
O=gpurun_out
mkdir -p $O
timeout 300 python tools/heads_prof.py 3 > $O/heads_plain_$TAG.log 2>&1; echo "plain exit $?"; tail -2 $O/heads_plain_$TAG.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"core2d_kernel|winv_kernel|wfwd_pipe_kernel|project_bwd_kernel|project_kernel|gw_reduce|lift" -s 50 -c 25 -o $O/heads_$TAG python tools/heads_prof.py 3 > $O/heads_ncu_$TAG.log 2>&1; echo "ncu exit $?"
python tools/ncuraw.py $O/heads_$TAG.ncu-rep | cut -c1-210 | tee $O/heads_${TAG}_summary.txt
ncu -i $O/heads_$TAG.ncu-rep --page source --csv > $O/heads_src_$TAG.csv 2>/dev/null
for pat in "core2d_kernel<(bool)0" "winv_kernel<(int)1" "winv_kernel<(int)2" "wfwd_pipe_kernel" "project_bwd_kernel"; do python tools/srcpage.py $O/heads_src_$TAG.csv "$pat" 0; done | tee -a $O/heads_${TAG}_summary.txt
ls -la $O/heads_$TAG.ncu-rep
