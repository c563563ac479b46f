#!/usr/bin/env bash
# Per-instantiation blur kernel times (ncu launch list, cold-cache) of a 64-frame 1080p quick_bench; env passes through.
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file $O/bt_$1.csv python tools/quick_bench.py 1920 1080 64 > /dev/null 2>&1
python - "$O/bt_$1.csv" <<'P'
import csv, sys, collections, re
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; k = hdr.index("Kernel Name"); v = hdr.index("Metric Value"); u = hdr.index("Metric Unit")
t = collections.defaultdict(list)
for r in rows[1:]:
    if "blur" in r[k]:
        x = float(r[v].replace(",", "")); x = x / 1000 if r[u] in ("ns", "nsecond") else x
        t[re.sub(r"\(.*", "", r[k])].append(x)
for n, xs in sorted(t.items()):
    print(f"{n:50s} n={len(xs):3d} big={max(xs):8.1f} us  sum/iter={sum(xs) / max(1, len(xs)) :8.1f}")
P
