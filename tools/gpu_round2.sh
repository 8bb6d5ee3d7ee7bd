#!/bin/bash
# GPU pass 2: all parity tests, the sweep points that failed, one ncu --set full pass over an eager step.
TAG=${1:-x}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -25 $O/pytest_$TAG.log
timeout 300 python tools/sweep.py --grids 128 --modes 64 --graphs >> $O/sweep_$TAG.jsonl 2>> $O/sweep_$TAG.err; echo "sweep 128 exit $?"
timeout 400 python tools/sweep.py --grids 256 --modes 32,64 --graphs >> $O/sweep_$TAG.jsonl 2>> $O/sweep_$TAG.err; echo "sweep 256 exit $?"
cut -c1-330 $O/sweep_$TAG.jsonl
B="python bench.py --steps 1 --warmup 1 --no-graphs --no-cpu-baseline"
timeout 300 $B > $O/plain_$TAG.log 2>&1; echo "plain exit $?"
timeout 900 ncu --set full --clock-control none --csv --page raw --log-file $O/full_$TAG.csv -k regex:bdn -s 80 -c 80 $B > $O/ncu_full.log 2>&1; echo "ncu full exit $?"
python tools/ncucsv.py $O/full_$TAG.csv --json $O/full_$TAG.json > $O/full_${TAG}_summary.txt 2>&1; head -50 $O/full_${TAG}_summary.txt
timeout 600 ncu --set full --clock-control none --csv --page raw --log-file $O/full_tc_$TAG.csv -k regex:tc_ -s 8 -c 8 $B --prec tf32 > $O/ncu_full_tc.log 2>&1; echo "ncu tc exit $?"
python tools/ncucsv.py $O/full_tc_$TAG.csv > $O/full_tc_${TAG}_summary.txt 2>&1; cat $O/full_tc_${TAG}_summary.txt
