#!/bin/bash
# One GPU pass: parity tests, a bench line with the full kernel table, and (optionally) an ncu launch list.
# usage: tools/gpu_check.sh TAG [ncu]
TAG=${1:-x}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 30 --warmup 5 --top 40 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 800 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("value %.1f samples/s  %.3f ms/step  e2e %.1f  kernel-sum %.0f us/step launches %d"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["kernel_time_us_per_step"],d["gpu_launches"]))
for k in d["top_kernels"]: print("%-22s %5.1f x %8.1f us/step  %5.1f%%"%(k["kernel"],k["launches_per_step"],k["us_per_step"],100*k["share"]))
PY
if [ "$2" == "ncu" ]; then
  B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
  timeout 300 $B > gpurun_out/plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/ncu1.log 2>&1
  python tools/launchlist.py gpurun_out/launches_$TAG.csv | head -60
fi
