#!/bin/bash
# round-2 final measurement pass on one GPU: both bench arms, the 3xTF32 mode, batch 32, the other workloads, sweep corners,
# the unchanged reference scripts through the harness, the ncu launch list, the warm timeline
TAG=${1:-r2f}
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
for wl in 1d_FPE 1d_GPE 2d_NC blindno_2d; do
  timeout 600 python bench.py --workload $wl --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_${wl}_$TAG.json 2> $O/err_$wl.log; show $O/bench_${wl}_$TAG.json
done
timeout 900 python tools/sweep.py --graphs --steps 4 --warmup 2 --bags 100,400 --grids 61,256 --modes 12,64 > $O/sweep_$TAG.jsonl 2> $O/sweep_$TAG.err; cut -c1-200 $O/sweep_$TAG.jsonl
: > $O/scripts_$TAG.jsonl
R=oracle/_ref
for s in 2d_FPE/train_fno.py 2d_Non_conservative_FPE/train_fno.py 2d_FPE/train_nio.py; do
  timeout 600 python tools/run_reference_script.py $R/$s --steps 20 --warmup 3 --bag 100 --workdir $O/script_run 2>$O/err_script.log | grep '^{' | tail -1 >> $O/scripts_$TAG.jsonl
done
timeout 600 python tools/run_reference_script.py $R/1d_FPE/train_fno.py --steps 20 --warmup 3 --bag 100 --samples 160 --workdir $O/script_run 2>>$O/err_script.log | grep '^{' | tail -1 >> $O/scripts_$TAG.jsonl
timeout 600 python tools/run_reference_script.py $R/1d_GPE/train_nio_GPE.py --steps 20 --warmup 3 --bag 100 --samples 80 --workdir $O/script_run 2>>$O/err_script.log | grep '^{' | tail -1 >> $O/scripts_$TAG.jsonl
timeout 600 python tools/run_reference_script.py $R/2d_FPE/train_fno.py --steps 6 --warmup 1 --samples 20 --workdir $O/script_run/train --save-ckpt $O/script_run/ckpt.pt --ckpt-prefix module. --seed 5 2>>$O/err_script.log | grep '^{' | tail -1 >> $O/scripts_$TAG.jsonl
timeout 600 python tools/run_reference_script.py $R/2d_FPE/eval_fno.py --samples 8 --workdir $O/script_run/eval --script-args "--ckpt $PWD/$O/script_run/ckpt.pt --outdir out --start 0 --end 5" 2>>$O/err_script.log | grep '^{' | tail -1 >> $O/scripts_$TAG.jsonl
cut -c1-260 $O/scripts_$TAG.jsonl; tail -3 $O/err_script.log | cut -c1-300
rm -rf $O/script_run
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $B > $O/plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file $O/launches_$TAG.csv $B > $O/ncu1.log 2>&1
python tools/launchlist.py $O/launches_$TAG.csv > $O/launches_${TAG}_summary.txt 2>&1; head -12 $O/launches_${TAG}_summary.txt | cut -c1-160
timeout 300 python tools/timeline.py --out $O/timeline_$TAG.json > $O/timeline_$TAG.log 2>&1; tail -2 $O/timeline_$TAG.log
