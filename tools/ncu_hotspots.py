#!/usr/bin/env python3
"""Developer tool: per-source-line hot spots of one kernel in an .ncu-rep (needs -lineinfo and --import-source on).
usage: ncu_hotspots.py report.ncu-rep [kernel-substring] [top]"""
import collections
import os
import re
import sys

sys.path.insert(0, "/opt/nvidia/nsight-compute/2025.2.1/extras/python")
import ncu_report  # noqa: E402


def main():
    rep = ncu_report.load_report(sys.argv[1])
    pat = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    rng = rep.range_by_idx(0)
    for ai in range(rng.num_actions()):
        act = rng.action_by_idx(ai)
        if pat not in act.name():
            continue
        inst = act.metric_by_name("inst_executed")
        smp = act.metric_by_name("smsp__pcsamp_sample_count")
        names = ["barrier", "short_scoreboard", "long_scoreboard", "wait", "no_instructions", "math_pipe_throttle", "mio_throttle", "branch_resolving", "selected"]
        stall = {n: act.metric_by_name("smsp__pcsamp_warps_issue_stalled_" + n) for n in names}
        pcs = inst.correlation_ids()
        by = collections.defaultdict(lambda: collections.defaultdict(float))
        n = pcs.num_instances()
        for i in range(n):
            pc = pcs.as_uint64(i)
            si = act.source_info(pc)
            key = (os.path.basename(si.file_name()), si.line()) if si else ("?", 0)
            by[key]["inst"] += inst.as_uint64(i)
        for nm, m in list(stall.items()) + [("samples", smp)]:
            if m is None:
                continue
            ids = m.correlation_ids()
            for i in range(ids.num_instances()):
                si = act.source_info(ids.as_uint64(i))
                key = (os.path.basename(si.file_name()), si.line()) if si else ("?", 0)
                by[key][nm] += m.as_uint64(i)
        tot_i = sum(v["inst"] for v in by.values())
        tot_s = sum(sum(v[nm] for nm in names) for v in by.values())
        print("kernel %s: warp instructions %.3e, stall samples %d" % (act.name(), tot_i, tot_s))
        srcs = {}
        rows = sorted(by.items(), key=lambda kv: -sum(kv[1][nm] for nm in names))
        for (fn, ln), v in rows[:top]:
            s = sum(v[nm] for nm in names)
            if fn not in srcs:
                try:
                    cand = [os.path.join(dp, fn) for dp, _, fs in os.walk(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))) if fn in fs]
                    srcs[fn] = open(cand[0]).read().split("\n") if cand else []
                except Exception:
                    srcs[fn] = []
            text = srcs[fn][ln - 1].strip()[:90] if 0 < ln <= len(srcs[fn]) else ""
            print("%-22s %4d inst %5.1f%% smp %5.1f%% | bar %3.0f%% ssb %3.0f%% wait %3.0f%% noinst %3.0f%% sel %3.0f%% | %s" % (
                fn, ln, 100 * v["inst"] / tot_i, 100 * s / max(tot_s, 1), 100 * v["barrier"] / max(s, 1), 100 * v["short_scoreboard"] / max(s, 1),
                100 * v["wait"] / max(s, 1), 100 * v["no_instructions"] / max(s, 1), 100 * v["selected"] / max(s, 1), text))
        # per-function aggregation (function = nearest preceding line that looks like a CPDP_* function header)
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
        heads = {}
        for (fn, ln), v in by.items():
            if fn not in heads:
                cand = [os.path.join(dp, fn) for dp, _, fs in os.walk(root) if fn in fs]
                hs = []
                if cand:
                    for i, line in enumerate(open(cand[0]).read().split("\n"), 1):
                        m = re.match(r"^\s*(?:template.*>\s*)?CPDP_(?:D|HD|GLOBAL|D_NOINLINE)\b.*?(\w+)\s*\(", line)
                        if m:
                            hs.append((i, m.group(1)))
                heads[fn] = hs
            name = fn
            for i, nm in heads[fn]:
                if i <= ln:
                    name = nm
            a = agg[name]
            a[0] += v["inst"]; a[1] += sum(v[nm] for nm in names); a[2] += v["barrier"]
        print("--- by function")
        for name, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            print("%-28s inst %5.1f%%  samples %5.1f%%  (barrier part %5.1f%%)" % (name, 100 * a[0] / tot_i, 100 * a[1] / max(tot_s, 1), 100 * a[2] / max(tot_s, 1)))
        return


if __name__ == "__main__":
    main()
