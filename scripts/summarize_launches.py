"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel:
    python scripts/summarize_launches.py gpurun_out/launches_bench.csv profiles/out.csv
Writes one row per kernel (launches, average / total duration, share of the captured GPU time)
followed by the raw launch list (id, kernel, block, grid, ns)."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if r]
hi = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui, bi, gi = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Metric Unit", "Block Size", "Grid Size"))
agg = collections.OrderedDict(); raw = []
for r in data:
    name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    v = float(r[vi].replace(",", "")); u = r[ui]
    v *= {"ns": 1, "nsecond": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9}[u]
    agg.setdefault(name, []).append(v); raw.append((r[0], name, r[bi], r[gi], int(v)))
tot = sum(sum(v) for v in agg.values())
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "launches", "avg_us", "total_us", "share_pct"])
    for k, v in agg.items():
        w.writerow([k, len(v), "%.1f" % (sum(v) / len(v) / 1e3), "%.1f" % (sum(v) / 1e3), "%.2f" % (100 * sum(v) / tot)])
    w.writerow([]); w.writerow(["id", "kernel", "block", "grid", "duration_ns"])
    w.writerows(raw)
print("wrote", sys.argv[2])
