"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list."""
import collections
import csv
import sys


def summarize(path, limit=None):
    with open(path) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        k = row["Kernel Name"].split("(")[0]
        m, v = row["Metric Name"], float(row["Metric Value"].replace(",", ""))
        a = agg.setdefault(k, {"n": 0, "t": 0.0, "rd": 0.0, "wr": 0.0})
        if m == "gpu__time_duration.sum":
            a["t"] += v; a["n"] += 1
        elif m == "dram__bytes_read.sum":
            a["rd"] += v
        elif m == "dram__bytes_write.sum":
            a["wr"] += v
    tot = 0.0
    for k, a in list(agg.items())[:limit]:
        print(f"  {k[:44]:44s} n={a['n']} t={a['t'] / 1e3:8.1f}us rd={a['rd'] / 1e6:7.1f}MB wr={a['wr'] / 1e6:7.1f}MB")
        tot += a["t"]
    print(f"  total {tot / 1e3:.1f} us")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print(p)
        summarize(p)
