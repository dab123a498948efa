"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, total, mean, share."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as fh:
    lines = [l for l in fh if l.startswith('"')]
rd = csv.DictReader(lines)
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    tot[name][0] += 1; tot[name][1] += us
s = sum(v[1] for v in tot.values()) or 1.0
print(f"{'kernel':60s} {'launches':>8s} {'total_us':>12s} {'mean_us':>10s} {'share':>7s}")
for k, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:60]:60s} {c:8d} {t:12.1f} {t / c:10.2f} {t / s:7.3f}")
