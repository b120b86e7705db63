#!/bin/bash
# Developer tool: the round-2 profile set for profiles/ (run on the GPU box through gpurun).
#   launch list of one bench step (gpu__time_duration, cold + serialised: shares, not absolutes)
#   one `ncu --set full` capture per hot kernel, condensed by tools/ncu_summary.py / ncu_hotspots.py
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bdf.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_launches.log 2>&1
echo "launch list rc=$?"
for k in k_riccati_bdf k_aux_forward k_stage_hessian k_stage_adjoint k_newton_step; do
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:$k -c 1 -f -o gpurun_out/r02_$k \
      python bench.py --steps 1 --warmup 0 --chunks 1 --no-cpu-baseline > gpurun_out/r02_ncu_$k.log 2>&1
  echo "$k rc=$?"
  python tools/ncu_summary.py gpurun_out/r02_$k.ncu-rep $k 4096 50 gpurun_out/r02_ncu_$k.json > /dev/null
  python tools/ncu_hotspots.py gpurun_out/r02_$k.ncu-rep $k 30 > gpurun_out/r02_ncu_${k}_hotspots.txt 2>&1
  rm -f gpurun_out/r02_$k.ncu-rep
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_ncu_k_*.json")):
    d = json.load(open(f))
    print(d["kernel"], "ms %.2f" % d["duration_ms_under_ncu"], "regs %d" % d["registers_per_thread"], "ipc %.2f" % d["ipc_per_sm"],
          "issue %.1f%%" % d["issue_active_pct"], "fp64 %.1f%%" % d["fp64_pipe_active_pct"], "warps %.1f%%" % d["warps_active_pct"],
          "dram MB %.0f" % ((d["dram_bytes_read"] + d["dram_bytes_write"]) / 1e6), {k: round(v, 2) for k, v in d["stall_per_issue"].items() if v and v > 0.3})
PY
