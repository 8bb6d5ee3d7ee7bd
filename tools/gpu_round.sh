#!/bin/bash
# One GPU pass for profiles/: parity tests, both bench arms, the C5 sweep (1 GPU), an ncu launch list.
# usage: tools/gpu_round.sh TAG
TAG=${1:-x}
O=gpurun_out
mkdir -p $O
python -c "import torch; print(torch.cuda.get_device_name(0))"
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -12 $O/pytest_$TAG.log
timeout 600 python bench.py --steps 50 --warmup 5 --top 40 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench exit $?"
tail -c 600 $O/bench_$TAG.err
python - <<PY
import json
try:
    d=json.load(open("$O/bench_$TAG.json"))
    print("value %.1f samples/s  %.3f ms/step  e2e %.1f  launches %d cpu %s"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["gpu_launches"],d.get("cpu_baseline")))
    print("roofline", {k:v for k,v in d["roofline"].items() if k!="note"})
    for k in d["top_kernels"][:28]: print("%-22s %5.1f x %8.1f us/step  %5.1f%%"%(k["kernel"],k["launches_per_step"],k["us_per_step"],100*k["share"]))
except Exception as e: print("bench parse failed", e)
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_$TAG.json 2> $O/bench_ref_$TAG.err; echo "ref exit $?"; cat $O/bench_ref_$TAG.json | cut -c1-400
for n in 61 80 128 256; do
  timeout 400 python tools/sweep.py --grids $n --graphs --steps 5 --warmup 2 >> $O/sweep_$TAG.jsonl 2>> $O/sweep_$TAG.err; echo "sweep $n exit $?"
done
cut -c1-260 $O/sweep_$TAG.jsonl
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 400 --csv --log-file $O/launches_$TAG.csv $B > $O/ncu1.log 2>&1; echo "ncu exit $?"
python tools/launchlist.py $O/launches_$TAG.csv > $O/launches_${TAG}_summary.txt 2>&1; head -40 $O/launches_${TAG}_summary.txt
