"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5 and r[0].isdigit()]
tot = defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[4].replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"[<(].*", "", name).split("::")[-1]
    ns = float(r[-1])
    tot[name][0] += 1
    tot[name][1] += ns
total = sum(v[1] for v in tot.values())
print("launches %d total %.3f ms" % (len(rows), total / 1e6))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-46s n=%5d  %9.3f ms  %5.1f%%  avg %8.1f us" % (k[:46], v[0], v[1] / 1e6, 100 * v[1] / total, v[1] / v[0] / 1e3))
