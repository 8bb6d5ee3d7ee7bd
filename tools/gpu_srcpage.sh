#!/bin/bash
# ncu --set full with sources for the dominant kernel pair (project_bwd, project of the per-snapshot net): SASS-level
# instruction mix and stall samples of their inner loops (tools/srcpage.py).
TAG=${1:-x}
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 1 --warmup 1 --no-graphs --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"project_bwd_kernel|project_kernel" -s 6 -c 6 -o $O/proj_$TAG $B > $O/ncu_src.log 2>&1; echo "ncu exit $?"
ncu -i $O/proj_$TAG.ncu-rep --page source --csv > $O/proj_src_$TAG.csv 2>/dev/null
for pat in "project_bwd_kernel<(int)4" "project_kernel<(int)4"; do python tools/srcpage.py $O/proj_src_$TAG.csv "$pat" 0; done | tee $O/proj_src_${TAG}_summary.txt
ls -la $O/proj_$TAG.ncu-rep
