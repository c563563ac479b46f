"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel:
launch count, total device time and share of the captured region.
Usage: python tools/summarize_launches.py gpurun_out/launches.csv [steps] > profiles/launches_summary.md"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "second": 1e9, "s": 1e9}.get(unit, 1)
        name = re.sub(r"\(.*$", "", r["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
        rows.append((name, ns, r["Grid Size"], r["Block Size"]))
    agg = OrderedDict()
    for name, ns, grid, block in rows:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    total = sum(a[1] for a in agg.values())
    print(f"source: {path}; {len(rows)} launches captured ({steps} step(s)); cold-cache, serialised per-launch times: compare SHARES\n")
    print("| kernel | launches | total ms | ms / step | share |")
    print("|---|---:|---:|---:|---:|")
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {n} | {ns / 1e6:.3f} | {ns / 1e6 / steps:.3f} | {100 * ns / total:.1f} % |")
    print(f"| total | {len(rows)} | {total / 1e6:.3f} | {total / 1e6 / steps:.3f} | 100 % |")


if __name__ == "__main__":
    main()
