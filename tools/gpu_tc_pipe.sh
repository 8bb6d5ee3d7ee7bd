#!/bin/bash
# ncu --set full of the fused tcgen05 layer kernels (3xTF32) at three shapes: the per-snapshot net at the default bag
# (300 images x 4 channels), at batch 32 (2400 x 4) and the heads (4 and 32 images x 12 channels, 32 modes);
# per-kernel duration, tensor-pipe %, DRAM bytes, issue %.  Each program is run plain first.
TAG=${1:-x}
O=gpurun_out
mkdir -p $O
: > $O/tc_pipe_${TAG}_summary.txt
for shape in snap snap32 heads heads32; do
  timeout 120 python tools/tc_layer_prof.py $shape 2 > $O/tc_plain_$shape.log 2>&1 || { echo "plain $shape failed"; tail -3 $O/tc_plain_$shape.log; continue; }
  timeout 600 ncu --set full --clock-control none -k regex:"p_kernel|q_kernel" -s 4 -c 4 --csv --page raw --log-file $O/tc_pipe_${TAG}_$shape.csv python tools/tc_layer_prof.py $shape 2 > $O/tc_ncu_$shape.log 2>&1; echo "ncu $shape exit $?"
  echo "== $shape" >> $O/tc_pipe_${TAG}_summary.txt
  python tools/ncucsv.py $O/tc_pipe_${TAG}_$shape.csv --json $O/tc_pipe_${TAG}_$shape.json >> $O/tc_pipe_${TAG}_summary.txt 2>&1
done
cat $O/tc_pipe_${TAG}_summary.txt | cut -c1-230
