"""Turn `ncu -i X.ncu-rep --page raw --csv` output into the per-kernel summary table kept in this directory.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv
    python profiles/summarize.py /tmp/raw.csv > profiles/rNN_kernels.md
"""
import csv
import sys

COLS = [
    ("gpu__time_duration.sum", "time_us"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_%"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads/inst"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_pipe_%"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("dram__bytes_read.sum", "dram_rd_MB"),
    ("dram__bytes_write.sum", "dram_wr_MB"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_%"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_%"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall_no_inst"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    seen = {}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d["Kernel Name"].split("(")[0][:44]
        seen.setdefault(name, d)  # first launch of each kernel
    print("| kernel | " + " | ".join(c[1] for c in COLS) + " |")
    print("|---|" + "---|" * len(COLS))
    u = dict(zip(hdr, units))
    for name, d in seen.items():
        vals = []
        for k, label in COLS:
            v = d.get(k, "")
            try:
                f = float(v)
                if label == "time_us" and u.get(k, "") in ("ns", "nsecond"):
                    f /= 1000.0
                if label == "time_us" and u.get(k, "") in ("ms", "msecond"):
                    f *= 1000.0
                if label.endswith("_MB") and u.get(k, "") in ("byte", "B"):
                    f /= 1e6
                if label.endswith("_MB") and u.get(k, "").lower() in ("kbyte", "kb"):
                    f /= 1e3
                v = f"{f:.2f}" if abs(f) < 1000 else f"{f:.0f}"
            except ValueError:
                pass
            vals.append(v)
        print(f"| {name} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
