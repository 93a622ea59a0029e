"""Summarise an `ncu --page source --csv` export (SASS view): stall-reason totals and the hottest instructions.
usage: python tools/ncu_src_summary.py file.csv [top_n]"""
import csv
import sys
from collections import Counter

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
allrows = list(csv.reader(open(path)))
# the export holds one section per profiled launch: ["Kernel Name", name] / header / instructions...; take the FIRST section whose
# kernel name contains argv[3] (default: first section)
want = sys.argv[3] if len(sys.argv) > 3 else ""
starts = [i for i, r in enumerate(allrows) if r and r[0] == "Kernel Name"]
sec = next((i for i in starts if want in allrows[i][1]), starts[0])
end = next((j for j in starts if j > sec), len(allrows))
rows = allrows[sec:end]
print("kernel:", rows[0][1][:80])
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = Counter()
body = [r for r in rows[2:] if len(r) >= len(hdr)]
samples = 0
for r in body:
    for s in stalls:
        tot[s] += int(r[idx[s]] or 0)
    samples += int(r[idx["# Samples"]] or 0)
print("total samples", samples)
for s, v in tot.most_common(12):
    print(f"  {s:28s} {v:8d} {100.0 * v / max(1, samples):5.1f}%")
ops = Counter()
inst = Counter()
for r in body:
    op = r[idx["Source"]].split()[0] if r[idx["Source"]].split() else "?"
    if op.startswith("@"):
        op = r[idx["Source"]].split()[1]
    ops[op.split(".")[0]] += int(r[idx["# Samples"]] or 0)
    inst[op.split(".")[0]] += int(r[idx["Instructions Executed"]] or 0)
print("samples by opcode:")
for o, v in ops.most_common(18):
    print(f"  {o:14s} samples {v:8d} ({100.0 * v / max(1, samples):5.1f}%)  executed {inst[o]}")
print("hottest instructions:")
order = sorted(range(len(body)), key=lambda i: -int(body[i][idx["# Samples"]] or 0))[:top]
for i in sorted(order):
    r = body[i]
    main = max(stalls, key=lambda s: int(r[idx[s]] or 0))
    print(f"  {i:5d} {int(r[idx['# Samples']]):7d} {main:22s} exec {r[idx['Instructions Executed']]:>9s}  {r[idx['Source']].strip()[:90]}")
