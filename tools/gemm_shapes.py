"""Join the GEMM launches of an ncu launch list (one graph replay, OF_WGRAD_SIDE=0: capture order = issue order) with the shape tags
of tools/dump_gemm_tags.py -> per-shape table: launches, total ncu time, algorithmic TFLOP/s.   usage: gemm_shapes.py launches.csv tags.json"""
import collections
import csv
import json
import sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
idx = {n: i for i, n in enumerate(h)}
times = []
for r in rows[hi + 1:]:
    if len(r) == len(h) and r[idx["Metric Name"]] == "gpu__time_duration.sum" and "gemm_kernel" in r[idx["Kernel Name"]]:
        v = float(r[idx["Metric Value"]].replace(",", ""))
        u = r[idx["Metric Unit"]]
        times.append(v / 1000 if u.startswith("n") else v * 1000 if u.startswith("m") else v)
tags = [t for t in json.load(open(sys.argv[2])) if t["family"].startswith("gemm_kernel")]
assert len(times) == len(tags), (len(times), len(tags))
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for us, t in zip(times, tags):
    a = agg[t["tag"]]
    a[0] += 1; a[1] += us; a[2] += t["flops"]; a[3] += t["ms"] * 1e3
tot = sum(a[1] for a in agg.values())
print(f"{len(times)} GEMM launches, {tot / 1e3:.2f} ms under ncu, {sum(a[2] for a in agg.values()) / tot / 1e6:.0f} TFLOP/s\n")
print("| shape | launches | ncu ms | share | us/launch | TFLOP/s | event-timed us/launch |\n|---|---:|---:|---:|---:|---:|---:|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k} | {a[0]} | {a[1] / 1e3:.2f} | {100 * a[1] / tot:.1f}% | {a[1] / a[0]:.1f} | {a[2] / a[1] / 1e6:.0f} | {a[3] / a[0]:.1f} |")
