#!/usr/bin/env python3
"""Developer tool: condenses one kernel of an .ncu-rep (`ncu --set full`) into the small JSON that profiles/ keeps and
bench.py reads for `roofline.traffic`.  usage: ncu_summary.py report.ncu-rep kernel-substring batch n_grid out.json"""
import json
import sys

sys.path.insert(0, "/opt/nvidia/nsight-compute/2025.2.1/extras/python")
import ncu_report  # noqa: E402

rep = ncu_report.load_report(sys.argv[1])
pat, batch, n_grid, out = sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
rng = rep.range_by_idx(0)
for ai in range(rng.num_actions()):
    act = rng.action_by_idx(ai)
    if pat not in act.name():
        continue

    def m(name):
        x = act.metric_by_name(name)
        return None if x is None else x.as_double()
    d = {
        "kernel": act.name(), "batch": batch, "n_grid": n_grid, "report": sys.argv[1].split("/")[-1],
        "duration_ms_under_ncu": m("gpu__time_duration.sum") / 1e6,
        "dram_bytes_read": m("dram__bytes_read.sum"), "dram_bytes_write": m("dram__bytes_write.sum"),
        "registers_per_thread": m("launch__registers_per_thread"), "block_size": m("launch__block_size"), "grid_size": m("launch__grid_size"),
        "shared_mem_per_block_bytes": m("launch__shared_mem_per_block"),
        "occupancy_limit_blocks_regs": m("launch__occupancy_limit_registers"), "occupancy_limit_blocks_smem": m("launch__occupancy_limit_shared_mem"),
        "warp_instructions": m("smsp__inst_executed.sum"), "ipc_per_sm": m("sm__inst_executed.avg.per_cycle_elapsed"),
        "issue_active_pct": m("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "fp64_pipe_active_pct": m("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        "tensor_pipe_pct": m("sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active"),
        "dram_throughput_pct": m("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "warps_active_pct": m("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "stall_per_issue": {k: m("smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % k) for k in
                            ("barrier", "wait", "short_scoreboard", "long_scoreboard", "no_instruction", "branch_resolving", "math_pipe_throttle", "not_selected")},
        "local_load_inst": m("sass__inst_executed_local_loads"), "local_store_inst": m("sass__inst_executed_local_stores"),
        "shared_load_inst": m("smsp__inst_executed_op_shared_ld.sum"), "shared_store_inst": m("smsp__inst_executed_op_shared_st.sum"),
        "shared_wavefronts": m("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
        "shared_bank_conflicts": m("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
        "lsu_pipe_pct": m("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        "l1tex_data_pipe_pct": m("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        "mio_throttle_per_issue": m("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"),
        "lg_throttle_per_issue": m("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
        "dispatch_stall_per_issue": m("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"),
        "elapsed_cycles_sm": m("sm__cycles_elapsed.avg"),
    }
    json.dump(d, open(out, "w"), indent=1)
    print(json.dumps(d))
    break
