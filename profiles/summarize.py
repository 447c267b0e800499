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


if __name__ == "__main__" and not (len(sys.argv) > 2 and sys.argv[2] == "--traffic"):
    main(sys.argv[1])


STAGE_OF = [  # kernel-name prefix -> bench.py stage (b2c_stage_name); first match wins
    ("k_aabb", "aabb"), ("k_bounds", "bounds_keys"), ("k_keys", "bounds_keys"),
    ("k_row_offsets", "sort_proxies"), ("k_row_place", "sort_proxies"), ("k_row_order", "gather"), ("k_gather", "gather"),
    ("k_sweep", "sweep"), ("k_large", "large"),
    ("k_row_scan", "sort_pairs"), ("k_row_scatter", "sort_pairs"), ("k_row_sort_big", "sort_pairs"), ("k_row_rank", "sort_pairs"),
    ("k_pairs_unpack", "unpack_carry"), ("k_carry", "unpack_carry"), ("k_pair_first", "unpack_carry"), ("k_pair_delta", "unpack_carry"),
    ("k_clear_np", "classify_bin"), ("k_classify", "classify_bin"), ("k_partition16", "classify_bin"),
    ("k_sphere_sphere", "closed_form"), ("k_convex_plane", "closed_form"),
    ("k_gjk", "gjk_mesh"), ("k_mesh_query", "gjk_mesh"), ("k_epa", "epa_fold_count"),
    ("k_manifold", "epa_fold_count"), ("k_mesh_manifold", "epa_fold_count"), ("k_compact", "epa_fold_count"), ("k_count", "epa_fold_count"),
]


def traffic(path):
    """Sum dram__bytes_read+write per bench stage over ONE profiled step -> profiles/traffic.json (bytes per launch group)."""
    import json
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    u = dict(zip(hdr, units))
    out = {}
    sort_no = 0
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d["Kernel Name"]
        short = name.replace("void ", "").replace("<unnamed>::", "")
        stage = None
        if short.startswith("rs_"):
            if short.startswith("rs_reset"):
                sort_no += 1
            stage = {1: "sort_proxies", 2: "sort_pairs"}.get(sort_no, "classify_bin")
        else:
            for pre, st in STAGE_OF:
                if short.startswith(pre):
                    stage = st
                    break
        if stage is None:
            continue
        b = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            f = float(d.get(k, 0) or 0)
            unit = u.get(k, "").lower()
            f *= {"byte": 1, "b": 1, "kbyte": 1e3, "kb": 1e3, "mbyte": 1e6, "mb": 1e6, "gbyte": 1e9, "gb": 1e9}.get(unit, 1)
            b += f
        out[stage] = out.get(stage, 0.0) + b
    print(json.dumps({k: round(v) for k, v in out.items()}, indent=1))


if __name__ == "__main__" and len(sys.argv) > 2 and sys.argv[2] == "--traffic":
    traffic(sys.argv[1])
