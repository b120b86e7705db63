#!/bin/bash
# Developer tool: one GPU-box round = parity tests + sweep timing + (optional) bench / ncu.  usage: gpu_round.sh [tag] [steps...]
TAG=${1:-run}; shift
mkdir -p gpurun_out
for step in "$@"; do
  case $step in
    test) timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log;;
    prof) timeout 300 python tools/prof_aux.py --batch 4096 --reps 3 2>&1 | tail -1 | cut -c1-400;;
    bench) timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_$TAG.json')); print(d['value'], d['ms_per_step'], d['roofline']['phase_ms'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline'])";;
    benchrk) timeout 600 python bench.py --steps 5 --warmup 3 --mode rk45 --no-cpu-baseline > gpurun_out/bench_rk45_$TAG.json 2> gpurun_out/bench_rk45_$TAG.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_rk45_$TAG.json')); print(d['value'], d['ms_per_step'], d['roofline']['phase_ms'])";;
    ncu) timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_riccati_bdf -c 1 -f -o gpurun_out/${TAG}_bdf python tools/prof_aux.py --batch 4096 --reps 1 > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu rc=$?";;
    launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/launches_$TAG.log 2>&1; echo "launches rc=$?";;
    ref) timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref_$TAG.json;;
  esac
done
