#!/usr/bin/env python3
"""Developer tool: shared-memory wavefronts per source line of one kernel in an .ncu-rep (--set full, --import-source on):
actual vs ideal wavefronts, i.e. where the bank conflicts are.  usage: ncu_smem_lines.py report.ncu-rep kernel-substring [top] [--global]   (--global: rank the lines by global-memory sectors)"""
import collections
import os
import sys

sys.path.insert(0, "/opt/nvidia/nsight-compute/2025.2.1/extras/python")
import ncu_report  # noqa: E402

rep = ncu_report.load_report(sys.argv[1])
pat = sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 25
rng = rep.range_by_idx(0)
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for ai in range(rng.num_actions()):
    act = rng.action_by_idx(ai)
    if pat not in act.name():
        continue
    by = collections.defaultdict(lambda: collections.defaultdict(float))
    for nm in ("memory_l1_wavefronts_shared", "memory_l1_wavefronts_shared_ideal", "memory_l2_theoretical_sectors_global",
               "memory_l2_theoretical_sectors_global_ideal", "inst_executed"):
        m = act.metric_by_name(nm)
        if m is None:
            print("metric missing:", nm)
            continue
        ids = m.correlation_ids()
        for i in range(ids.num_instances()):
            si = act.source_info(ids.as_uint64(i))
            key = (os.path.basename(si.file_name()), si.line()) if si else ("?", 0)
            by[key][nm] += m.as_uint64(i)
    tot = sum(v["memory_l1_wavefronts_shared"] for v in by.values())
    toti = sum(v["memory_l1_wavefronts_shared_ideal"] for v in by.values())
    print("kernel %s: shared wavefronts %.3e (ideal %.3e)" % (act.name(), tot, toti))
    srcs = {}
    order = "memory_l2_theoretical_sectors_global" if "--global" in sys.argv else "memory_l1_wavefronts_shared"
    for (fn, ln), v in sorted(by.items(), key=lambda kv: -kv[1][order])[:top]:
        if fn not in srcs:
            cand = [os.path.join(dp, fn) for dp, _, fs in os.walk(root) if fn in fs]
            srcs[fn] = open(cand[0]).read().split("\n") if cand else []
        text = srcs[fn][ln - 1].strip()[:80] if 0 < ln <= len(srcs[fn]) else ""
        print("%-20s %4d  smem wavefronts %5.1f%%  x%.2f of ideal | global sectors %.2e (ideal %.2e) | %s" % (
            fn, ln, 100 * v["memory_l1_wavefronts_shared"] / max(tot, 1),
            v["memory_l1_wavefronts_shared"] / max(v["memory_l1_wavefronts_shared_ideal"], 1),
            v["memory_l2_theoretical_sectors_global"], v["memory_l2_theoretical_sectors_global_ideal"], text))
    break
