"""Top stall sites of an `ncu --page source --csv` export.  Usage: ncu_source_top.py file.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
# one section per profiled kernel instance: keep the one with the most samples
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
best = None
for si, hi in enumerate(starts):
    end = starts[si + 1] if si + 1 < len(starts) else len(rows)
    hdr = rows[hi]
    col = {k: i for i, k in enumerate(hdr)}
    body = [r for r in rows[hi + 1:end] if len(r) == len(hdr) and r[0] != "Address"]
    tot = sum(int(r[col["# Samples"]] or 0) for r in body)
    if best is None or tot > best[0]:
        best = (tot, hdr, col, body)
_, hdr, col, body = best
tot = sum(int(r[col["# Samples"]] or 0) for r in body)
stall_cols = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
print("total samples", tot, "instructions", len(body))
for rank, r in enumerate(sorted(body, key=lambda r: -int(r[col["# Samples"]] or 0))[:n]):
    s = int(r[col["# Samples"]] or 0)
    top = sorted(((int(r[col[k]] or 0), k[6:]) for k in stall_cols), reverse=True)[:3]
    print(f"{100 * s / tot:5.1f}%  #{body.index(r):4d} {r[col['Source']][:70]:70s} exec={r[col['Instructions Executed']]:>10s} " +
          " ".join(f"{k}:{v}" for v, k in top if v))
