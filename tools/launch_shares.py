"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel: python tools/launch_shares.py file.csv"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0
for row in csv.DictReader(lines):
    val = float(row["Metric Value"].replace(",", ""))
    ns = val * {"ns": 1, "us": 1e3, "ms": 1e6}.get(row["Metric Unit"], 1)
    short = re.sub(r"\(.*", "", row["Kernel Name"])[:84]
    agg[short][0] += 1
    agg[short][1] += ns
    tot += ns
print(f"# total {tot / 1e6:.3f} ms over {sum(v[0] for v in agg.values())} launches (cold-cache, serialised: compare shares)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:86s} n={v[0]:4d} ms={v[1] / 1e6:8.3f} avg_us={v[1] / v[0] / 1e3:8.1f} share={v[1] / tot:.3f}")
