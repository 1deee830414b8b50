"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, re, sys
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", ""))))
agg = collections.OrderedDict()
for k, v in rows:
    agg.setdefault(re.sub(r"\(.*", "", k)[:100], []).append(v)
tot = sum(v for _, v in rows)
print(f"{len(rows)} launches, total {tot/1e6:.3f} ms (ncu: cold-cache, serialised -- compare shares, not absolutes)")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{sum(v)/1e3:10.1f} us  n={len(v):3d}  avg {sum(v)/len(v)/1e3:8.1f} us  {100*sum(v)/tot:5.1f}%  {k}")
