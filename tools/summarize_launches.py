"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/<name>.md"""
import csv
import re
import sys
from collections import OrderedDict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}.get(unit, 1e-6)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        rows.append((name, v * scale))
    agg = OrderedDict()
    for n, ms in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    print(f"source: {path}; {len(rows)} launches, {tot:.3f} ms total device time (ncu-serialised, cold cache: compare shares)\n")
    print("| kernel | launches | total ms | share |")
    print("|---|---:|---:|---:|")
    for n, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n}` | {c} | {ms:.3f} | {100 * ms / tot:.1f}% |")


if __name__ == "__main__":
    main(sys.argv[1])
