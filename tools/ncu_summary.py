#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, without a GPU) into a small JSON of the metrics the design
decisions rest on. Usage: python tools/ncu_summary.py <file.ncu-rep> [out.json]"""
import csv, io, json, re, subprocess, sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_op_read.sum", "lts__t_bytes.sum",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_xu.sum",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]

def main():
    rep = sys.argv[1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        o = {"kernel": d.get("Kernel Name")}
        for k in KEYS:
            if k in d:
                u = units[hdr.index(k)]
                try:
                    v = float(d[k].replace(",", ""))
                except ValueError:
                    v = d[k]
                o[k] = [v, u] if u else v
        res.append(o)
    s = json.dumps(res if len(res) != 1 else res[0], indent=1)
    s = re.sub(r"\[\s+([^\[\]]+?),\s+(\"[^\"]*\")\s+\]", r"[\1, \2]", s)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(s + "\n")
    print(s)

if __name__ == "__main__":
    main()
