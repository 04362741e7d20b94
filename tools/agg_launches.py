"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.  python tools/agg_launches.py file.csv [skip_fraction]"""
import collections, csv, re, sys
with open(sys.argv[1]) as f:
    rows = list(csv.DictReader([l for l in f if not l.startswith("==")]))
skip = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
rows = rows[int(len(rows) * skip):]
agg = collections.OrderedDict()
for r in rows:
    k = re.sub(r"void |<unnamed>::|\(anonymous namespace\)::", "", r["Kernel Name"]).split("(")[0][:60]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r["Metric Value"].replace(",", ""))
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot/1e6:.3f} ms total\n\n| kernel | launches | us | share |\n|---|---|---|---|")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {c} | {t/1e3:.1f} | {100*t/tot:.1f}% |")
