"""Per-source-line stall samples of one kernel from an ncu report.

    ncu -i X.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:k_gjk --launch-count 1 > /tmp/src.csv
    python profiles/hotlines.py /tmp/src.csv [top]
"""
import csv
import sys
from collections import defaultdict


def main(path, top=40):
    csv.field_size_limit(1 << 30)
    per = defaultdict(lambda: [0, 0, 0, ""])   # samples, inst, thread inst, text
    fpath = ""
    hdr = None
    for r in csv.reader(open(path, errors="replace")):
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed"); ti = hdr.index("Thread Instructions Executed")
            continue
        if hdr is None or len(r) <= ti or not r[0].isdigit():
            continue
        try:
            s, i, t = int(r[si] or 0), int(r[ii] or 0), int(r[ti] or 0)
        except ValueError:
            continue
        k = (fpath, int(r[0]))
        per[k][0] += s; per[k][1] += i; per[k][2] += t
        if r[1].strip():
            per[k][3] = r[1].strip()[:100]
    tot = sum(v[0] for v in per.values()) or 1
    toti = sum(v[1] for v in per.values()) or 1
    print(f"total samples {tot}, warp instructions {toti}")
    print("| file:line | samples % | inst % | threads/inst | source |\n|---|---|---|---|---|")
    for k, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"| {k[0]}:{k[1]} | {100.0 * v[0] / tot:.1f} | {100.0 * v[1] / toti:.1f} | {v[2] / max(v[1], 1):.1f} | `{v[3]}` |")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
