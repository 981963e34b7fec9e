"""Summarise an `ncu --page raw --csv` export: one block per profiled launch with the metrics the roofline
discussion in DESIGN.md uses.    python scripts/summarize_ncu_raw.py gpurun_out/x_raw.csv > profiles/x.md"""
import csv
import re
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__occupancy_limit_registers", "occ limit regs (blocks/SM)"),
    ("launch__occupancy_limit_shared_mem", "occ limit smem (blocks/SM)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/TEX throughput %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor-pipe instructions"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__pipe_tensor_subpipe_tf32_cycles_active.avg.pct_of_peak_sustained_active", "tensor tf32 subpipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard_not_issued.ratio", "stall long_scoreboard (warp-cycles/issue)"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall membar"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "stall sleeping"),
    ("lts__t_sectors_op_atom.sum", "L2 atom sectors"),
    ("lts__t_sectors_op_red.sum", "L2 red sectors"),
]


def main(path, pattern=None):
    with open(path, newline="") as f:
        rows = [r for r in csv.reader(f) if r]
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units, data = rows[hdr_i], rows[hdr_i + 1], rows[hdr_i + 2:]
    col = {h: i for i, h in enumerate(hdr)}
    for r in data:
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]])
        if pattern and not re.search(pattern, name):
            continue
        print(f"## `{name}`  grid {r[col['Grid Size']]} block {r[col['Block Size']]}")
        for k, label in KEYS:
            if k in col and r[col[k]] != "":
                print(f"- {label}: {r[col[k]]} {units[col[k]]}")
        print()


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
