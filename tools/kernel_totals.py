"""Time per kernel name of an ncu launch list (csv with gpu__time_duration.sum):  python tools/kernel_totals.py launches.csv [top]"""
import collections
import csv
import re
import sys

rows = [l for l in open(sys.argv[1]) if not l.startswith("==")]
tot, cnt = collections.Counter(), collections.Counter()
for r in csv.DictReader(rows):
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void ", "", name)
    t = float(r["Metric Value"].replace(",", "")) / 1e3
    tot[name] += t
    cnt[name] += 1
allt = sum(tot.values())
print(f"total {allt:.0f} us over {sum(cnt.values())} launches")
for name, t in tot.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print(f"{t:10.1f} us {100 * t / allt:5.1f}%  n={cnt[name]:5d} avg={t / cnt[name]:8.1f}  {name[:110]}")
