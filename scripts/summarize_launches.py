"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv)."""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
d = collections.defaultdict(list)
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
    d[row["Kernel Name"][:70]].append((v, row.get("Grid Size")))
tot = sum(sum(x[0] for x in v) for v in d.values())
print(f"total {tot:.1f} us over {sum(len(v) for v in d.values())} launches")
for k, v in sorted(d.items(), key=lambda kv: -sum(x[0] for x in kv[1])):
    ts = [x[0] for x in v]
    print(f"{sum(ts):8.1f} n={len(ts):3d} avg={sum(ts) / len(ts):7.1f} min={min(ts):7.1f} max={max(ts):7.1f} grid={v[0][1]} {k}")
