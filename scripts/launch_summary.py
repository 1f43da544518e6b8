"""Summarise an ncu launch list (gpu__time_duration.sum CSV): time per kernel name, optionally for a launch-id window."""
import csv, sys, collections, re
path = sys.argv[1]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10**9
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum": continue
    i = int(r["ID"])
    v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
    rows.append((i, r["Kernel Name"], v))
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for i, k, v in rows:
    if lo <= i < hi:
        k = re.sub(r"\(.*$", "", k)[:110]
        agg[k][0] += 1; agg[k][1] += v; tot += v
print(f"launches {sum(a[0] for a in agg.values())}  total {tot/1e3:.2f} ms  (ids {lo}..{hi})")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{v/1e3:9.3f} ms {100*v/tot:5.1f}% n={n:5d} avg {v/n:8.1f} us  {k}")
