"""Summarise an ncu launch list (gpu__time_duration.sum CSV): per-kernel count, mean us, share."""
import csv, collections, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    k = row["Kernel Name"]; v = float(row["Metric Value"].replace(",", ""))
    agg.setdefault(k, []).append(v)
tot = sum(sum(v) for v in agg.values())
print(f"{'kernel':70s} {'n':>4s} {'mean_us':>10s} {'share':>7s}")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k[:70]:70s} {len(v):4d} {sum(v)/len(v)/1e3:10.1f} {sum(v)/tot*100:6.1f}%")
