#!/bin/bash
# round-2 measurement pass on one GPU: both bench arms, the 3xTF32 mode, batch 32, the other workloads, the ncu launch list
TAG=${1:-r2q}
O=gpurun_out
mkdir -p $O
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1].split("/")[-1], "value %.1f %s  %.3f ms/step  e2e %s  launches %s"%(d["value"],d["unit"],d.get("ms_per_step",0),d.get("e2e",{}).get("value"),d.get("gpu_launches")))
except Exception as e: print(sys.argv[1], "FAILED", e)
PY
}
timeout 900 python bench.py --steps 50 --warmup 5 --top 40 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; show $O/bench_$TAG.json
timeout 900 python bench.py --impl reference --steps 10 --warmup 2 > $O/bench_ref_$TAG.json 2> $O/err.log; show $O/bench_ref_$TAG.json
timeout 600 python bench.py --steps 50 --warmup 5 --top 40 --no-cpu-baseline --prec tf32x3 > $O/bench_tf32x3_$TAG.json 2> $O/err.log; show $O/bench_tf32x3_$TAG.json
timeout 600 python bench.py --steps 12 --warmup 4 --top 40 --no-cpu-baseline --batch-per-gpu 32 > $O/bench_b32_$TAG.json 2> $O/err.log; show $O/bench_b32_$TAG.json
timeout 600 python bench.py --steps 12 --warmup 4 --top 40 --no-cpu-baseline --batch-per-gpu 32 --prec tf32x3 > $O/bench_b32_tf32x3_$TAG.json 2> $O/err.log; show $O/bench_b32_tf32x3_$TAG.json
for wl in 1d_FPE 1d_GPE 2d_NC blindno_2d; do
  timeout 600 python bench.py --workload $wl --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_${wl}_$TAG.json 2> $O/err_$wl.log; show $O/bench_${wl}_$TAG.json
done
timeout 900 python tools/sweep.py --graphs --steps 4 --warmup 2 --bags 100,400 --grids 61,256 --modes 12,64 > $O/sweep_$TAG.jsonl 2> $O/sweep_$TAG.err; cut -c1-260 $O/sweep_$TAG.jsonl
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $B > $O/plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 400 --csv --log-file $O/launches_$TAG.csv $B > $O/ncu1.log 2>&1
python tools/launchlist.py $O/launches_$TAG.csv > $O/launches_${TAG}_summary.txt 2>&1; head -40 $O/launches_${TAG}_summary.txt
