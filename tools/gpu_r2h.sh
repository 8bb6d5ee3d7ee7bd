#!/bin/bash
# final default-arm lines of round 2: bench (both arms), launch list, timeline, ncu --set full of the tcgen05 W-forward on the default path
TAG=${1:-r2h}
O=gpurun_out
mkdir -p $O
timeout 900 python bench.py --steps 50 --warmup 5 --top 40 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; python -c "import json;d=json.load(open('$O/bench_$TAG.json'));print('default',round(d['value'],1),round(d['ms_per_step'],4),round(d['e2e']['value'],1),d['gpu_launches'],d['cpu_baseline']['value'])"
timeout 900 python bench.py --impl reference --steps 10 --warmup 2 > $O/bench_ref_$TAG.json 2> $O/err.log; cut -c1-200 $O/bench_ref_$TAG.json
timeout 600 python bench.py --steps 12 --warmup 4 --top 40 --no-cpu-baseline --batch-per-gpu 32 > $O/bench_b32_$TAG.json 2> $O/err.log; python -c "import json;d=json.load(open('$O/bench_b32_$TAG.json'));print('b32',round(d['value'],1),round(d['ms_per_step'],3))"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $B > $O/plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file $O/launches_$TAG.csv $B > $O/ncu1.log 2>&1
python tools/launchlist.py $O/launches_$TAG.csv > $O/launches_${TAG}_summary.txt 2>&1; head -8 $O/launches_${TAG}_summary.txt | cut -c1-160
B2="python bench.py --steps 1 --warmup 1 --no-graphs --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none -k regex:tc_wfwd -c 3 --csv --page raw --log-file $O/tc_wfwd_$TAG.csv $B2 > $O/ncu2.log 2>&1; python tools/ncucsv.py $O/tc_wfwd_$TAG.csv --json $O/tc_wfwd_$TAG.json 2>&1 | cut -c1-230
timeout 300 python tools/timeline.py --out $O/timeline_$TAG.json > $O/timeline_$TAG.log 2>&1; python -c "import json;d=json.load(open('$O/timeline_$TAG.json'));print('timeline',{k:round(v,1) for k,v in d.items() if k!='by_kernel_us_per_step'})"
