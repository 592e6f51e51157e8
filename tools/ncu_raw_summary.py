#!/usr/bin/env python3
"""Print the headline metrics of every kernel in an ncu report (ncu -i X --page raw --csv piped to a file)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"), ("smsp__inst_executed.sum", "warp_inst"),
        ("dram__bytes_read.sum", "dramR"), ("dram__bytes_write.sum", "dramW"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"), ("l1tex__t_sector_hit_rate.pct", "l1hit%"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"), ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall_branch"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
        ("l1tex__data_bank_conflicts_pipe_lsu.sum", "bankconf"),
        ("l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_ld.ratio", "sect/req_ld"),
        ("l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_st.ratio", "sect/req_st")]
for d in data:
    out = []
    for key, name in want:
        if key in idx:
            v = d[idx[key]]
            if key == "Kernel Name":
                v = v.split("(")[0][-28:]
            else:
                try:
                    v = f"{float(v.replace(',', '')):.4g}{units[idx[key]] if name in ('time', 'dramR', 'dramW') else ''}"
                except ValueError:
                    pass
            out.append(f"{name}={v}")
    print("  ".join(out))
