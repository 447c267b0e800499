"""Per-kernel and per-stage metrics of ONE profiled step from an `ncu --set full` capture -> profiles/kernel_metrics.json
(read by bench.py for the fp32-issue roofline of the dominant kernel) and profiles/traffic.json (DRAM bytes per stage).

    ncu -i gpurun_out/prof_c2_TAG.ncu-rep --page raw --csv > /tmp/raw.csv
    python profiles/kernel_metrics.py /tmp/raw.csv TAG

The counts used by bench.py (warp instructions, thread instructions, DRAM bytes per launch) depend only on the workload —
the same seeded C2 snapshot the bench times — not on the clock; times under ncu are cold-cache and serialised and are kept
only to show each kernel's share of the step."""
import csv
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from summarize import STAGE_OF  # noqa: E402

SCALE = {"byte": 1, "b": 1, "kbyte": 1e3, "kb": 1e3, "mbyte": 1e6, "mb": 1e6, "gbyte": 1e9, "gb": 1e9,
         "ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}


def num(d, u, k):
    try:
        return float(d.get(k, "") .replace(",", "")) * SCALE.get(u.get(k, "").lower(), 1.0)
    except ValueError:
        return 0.0


def stage_of(short):
    for pre, st in STAGE_OF:
        if short.startswith(pre):
            return st
    return "other"


def main(path, tag):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    u = dict(zip(hdr, units))
    kern, stages = {}, {}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        short = d["Kernel Name"].replace("void ", "").replace("<unnamed>::", "").replace("b2c::", "").split("(")[0]
        t = num(d, u, "gpu__time_duration.sum")
        inst = num(d, u, "smsp__inst_executed.sum")
        tpi = num(d, u, "smsp__thread_inst_executed_per_inst_executed.ratio")
        rec = kern.setdefault(short, {"launches": 0, "time_us": 0.0, "warp_inst": 0.0, "thread_inst": 0.0, "dram_bytes": 0.0,
                                      "issue_active_x_time": 0.0, "warps_active_x_time": 0.0, "fma_pipe_x_time": 0.0})
        rec["launches"] += 1
        rec["time_us"] += t
        rec["warp_inst"] += inst
        rec["thread_inst"] += inst * tpi
        rec["dram_bytes"] += num(d, u, "dram__bytes_read.sum") + num(d, u, "dram__bytes_write.sum")
        rec["issue_active_x_time"] += num(d, u, "smsp__issue_active.avg.pct_of_peak_sustained_active") * t
        rec["warps_active_x_time"] += num(d, u, "sm__warps_active.avg.pct_of_peak_sustained_active") * t
        rec["fma_pipe_x_time"] += num(d, u, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active") * t
        rec["regs"] = int(num(d, u, "launch__registers_per_thread"))
        rec["grid"] = int(num(d, u, "launch__grid_size"))
        rec["block"] = int(num(d, u, "launch__block_size"))
        rec["stage"] = stage_of(short)
    out = {}
    for k, r in kern.items():
        t = max(r["time_us"], 1e-9)
        out[k] = {"stage": r["stage"], "launches_per_step": r["launches"], "ncu_time_us": round(r["time_us"], 2),
                  "warp_inst_per_step": round(r["warp_inst"]), "thread_inst_per_step": round(r["thread_inst"]),
                  "threads_per_inst": round(r["thread_inst"] / max(r["warp_inst"], 1.0), 2),
                  "issue_active_pct": round(r["issue_active_x_time"] / t, 2), "warps_active_pct": round(r["warps_active_x_time"] / t, 2),
                  "fma_pipe_pct": round(r["fma_pipe_x_time"] / t, 2), "dram_bytes_per_step": round(r["dram_bytes"]),
                  "regs": r["regs"], "grid": r["grid"], "block": r["block"]}
        s = stages.setdefault(r["stage"], {"ncu_time_us": 0.0, "dram_bytes": 0.0, "thread_inst": 0.0, "warp_inst": 0.0})
        s["ncu_time_us"] += r["time_us"]; s["dram_bytes"] += r["dram_bytes"]; s["thread_inst"] += r["thread_inst"]; s["warp_inst"] += r["warp_inst"]
    total = sum(v["ncu_time_us"] for v in out.values())
    for v in out.values():
        v["share_of_step"] = round(v["ncu_time_us"] / max(total, 1e-9), 4)
    doc = {"source": f"ncu --set full --clock-control none, one profiled C2 step ({tag}); see profiles/README.md",
           "step_ncu_time_us": round(total, 1), "kernels": out,
           "stages": {k: {kk: round(vv, 1) for kk, vv in v.items()} for k, v in stages.items()}}
    json.dump(doc, open(os.path.join(HERE, "kernel_metrics.json"), "w"), indent=1)
    bp = ("aabb", "bounds_keys", "sort_proxies", "gather", "sweep", "large", "sort_pairs")
    tr = {k: round(v["dram_bytes"]) for k, v in stages.items()}
    tr["broadphase"] = round(sum(stages[k]["dram_bytes"] for k in bp if k in stages))
    tr["source"] = doc["source"]
    json.dump(tr, open(os.path.join(HERE, "traffic.json"), "w"), indent=1)
    print("| kernel | stage | n | ncu µs | share | regs | issue-active % | threads/inst | warps-active % | fma pipe % | DRAM MB |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    for k, v in sorted(out.items(), key=lambda kv: -kv[1]["ncu_time_us"]):
        print(f"| `{k}` | {v['stage']} | {v['launches_per_step']} | {v['ncu_time_us']:.1f} | {100 * v['share_of_step']:.1f} % | {v['regs']} | "
              f"{v['issue_active_pct']:.1f} | {v['threads_per_inst']:.1f} | {v['warps_active_pct']:.1f} | {v['fma_pipe_pct']:.1f} | {v['dram_bytes_per_step'] / 1e6:.1f} |")
    print(f"\ntotal {total:.1f} µs under ncu (cold caches, serialised launches)")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "r02")
