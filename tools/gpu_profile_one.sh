#!/bin/bash
# Developer tool: one `ncu --set full` capture of the named kernels (gpurun box), condensed into gpurun_out/.
# usage: bash tools/gpu_profile_one.sh k_stage_hessian k_newton_step
mkdir -p gpurun_out
for k in "$@"; do
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:$k -c 1 -f -o gpurun_out/r02_$k \
      python bench.py --steps 1 --warmup 0 --chunks 1 --no-cpu-baseline > gpurun_out/r02_ncu_$k.log 2>&1
  echo "$k rc=$?"
  python tools/ncu_summary.py gpurun_out/r02_$k.ncu-rep $k 4096 50 gpurun_out/r02_ncu_$k.json > /dev/null
  python tools/ncu_hotspots.py gpurun_out/r02_$k.ncu-rep $k 40 > gpurun_out/r02_ncu_${k}_hotspots.txt 2>&1
  rm -f gpurun_out/r02_$k.ncu-rep
done
