"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (markdown table)."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
idx = {n: i for i, n in enumerate(h)}
agg = collections.defaultdict(lambda: [0.0, 0])
tot = 0.0
for r in rows[hi + 1:]:
    if len(r) != len(h) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"<.*", "", re.sub(r"\(.*", "", r[idx["Kernel Name"]])).replace("void ", "").strip()
    v = float(r[idx["Metric Value"]].replace(",", ""))
    unit = r[idx["Metric Unit"]]
    v = v / 1000 if unit in ("ns", "nsecond") else v * 1000 if unit in ("ms", "msecond") else v
    agg[name][0] += v
    agg[name][1] += 1
    tot += v
print(f"Total {tot / 1000:.2f} ms over {sum(v[1] for v in agg.values())} launches.\n")
print("| kernel | ms | share | launches | us/launch |\n|---|---:|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"| `{k[:70]}` | {v[0] / 1000:.2f} | {100 * v[0] / tot:.1f}% | {v[1]} | {v[0] / v[1]:.1f} |")
