#!/usr/bin/env python3
"""Summarise .ncu-rep captures (ncu --set full) into a markdown table for profiles/.
Usage: python tools/ncu_summary.py out.md title=file.ncu-rep [title=file.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput, % of HW peak"),
    ("lts__t_sector_hit_rate.pct", "L2 sector hit rate"),
    ("l1tex__t_sector_hit_rate.pct", "L1 sector hit rate"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global load sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global load requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "global store sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "global store requests"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid size"),
    ("launch__block_size", "block size"),
    ("launch__waves_per_multiprocessor", "waves per SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / scheduler / cycle"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall: LG throttle"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock"),
]


def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    return hdr, units, data[-1]


def main():
    out_md = sys.argv[1]
    cols = []
    for arg in sys.argv[2:]:
        title, path = arg.split("=", 1)
        hdr, units, row = load(path)
        cols.append((title, hdr, units, row))
    with open(out_md, "w") as fh:
        fh.write("| metric | " + " | ".join(c[0] for c in cols) + " |\n")
        fh.write("|---|" + "---|" * len(cols) + "\n")
        name = []
        for _, hdr, _, row in cols:
            name.append(row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
        fh.write("| kernel | " + " | ".join("`" + n.replace("|", "/") + "`" for n in name) + " |\n")
        for key, label in METRICS:
            cells = []
            for _, hdr, units, row in cols:
                if key in hdr:
                    i = hdr.index(key)
                    v = row[i]
                    try:
                        fv = float(v.replace(",", ""))
                        v = f"{fv:,.3f}".rstrip("0").rstrip(".") if abs(fv) < 1e6 else f"{fv:,.0f}"
                    except ValueError:
                        pass
                    cells.append(f"{v} {units[i]}".strip())
                else:
                    cells.append("n/a")
            fh.write(f"| {label} (`{key}`) | " + " | ".join(cells) + " |\n")


if __name__ == "__main__":
    main()
