"""Per-kernel summary of an ncu `--metrics gpu__time_duration.sum --csv` launch list (last complete keyframe round).
usage: python tools/launch_summary.py launches.csv [--seq]"""
import collections
import csv
import re
import sys


def load(fn):
    with open(fn) as f:
        lines = [l for l in f if not l.startswith("==")]
    return [(int(x["ID"]), x["Kernel Name"], float(x["Metric Value"].replace(",", "")) / 1000, x["Grid Size"])
            for x in csv.DictReader(lines)]


def short(n):
    return re.sub(r"\(.*", "", n)[:52]


rows = load(sys.argv[1])
starts = [i for i, r in enumerate(rows) if r[1].startswith("k_store_begin")]
ends = [i for i, r in enumerate(rows) if r[1].startswith("k_lg_extract")]
s = starts[1] if len(starts) > 1 else starts[0]
e = [x for x in ends if x > s][0]
rnd = rows[s:e + 1]
print("round launches %d, total %.1f us" % (len(rnd), sum(r[2] for r in rnd)))
if "--seq" in sys.argv:
    for r in rnd:
        print(r[0], short(r[1]), r[3], round(r[2], 1))
agg = collections.OrderedDict()
for r in rnd:
    k = short(r[1])
    agg.setdefault(k, [0, 0.0])
    agg[k][0] += 1
    agg[k][1] += r[2]
for k, v in agg.items():
    print("%-54s n=%3d %9.1f us" % (k, v[0], v[1]))
