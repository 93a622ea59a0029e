"""Per-kernel HBM figures of ONE graph replay of the training micro-step, from an ncu launch list with
`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` (cold cache, serialised launches):

    python tools/hbm_kernels.py launches.csv > profiles/rNN_hbm_kernels.json

For every bandwidth-bound kernel: launches, total time, DRAM bytes moved, achieved GB/s over all launches and for the single
largest launch (the level-0 shape B4 x L4096 x C512), both against the measured copy bandwidth (MEASURED_PEAKS.json), and the
algorithmic bytes per element from the header comments in include/osufusion_b200.h."""
import collections
import csv
import json
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
peak = 6551.0
mp = ROOT / "MEASURED_PEAKS.json"
if mp.exists():
    peak = json.loads(mp.read_text()).get("hbm_gbs", peak)

# algorithmic MB (reads + writes) of the LARGEST launch of each kernel in the CFG-L B4 N4096 step (level 0: 16384 rows; residual blocks
# C = 512, attention q|k|v = 1024|64|64 columns), from the header comments in include/osufusion_b200.h
E0 = 4 * 4096 * 512 / 1e6          # 8.39 M elements of a (B, L, C) level-0 activation
ROWS = 4 * 4096 / 1e6
ALGO = {
    "layernorm_fwd_kernel": (4 + 4 + 2) * E0, "layernorm_bwd_kernel": (4 + 4 + 4) * E0, "rb_apply_fwd_kernel": (2 + 2) * E0,
    "rb_logit_pool_kernel": 2 * E0, "rb_gate_fwd_kernel": (2 + 4 + 4 + 2) * E0, "rb_gate_bwd_reduce_kernel": (4 + 2) * E0,
    "rb_bwd_pass1_kernel": (2 + 4 + 2) * E0, "rb_bwd_apply_kernel": (2 + 2 + 2) * E0, "rb_rowdot_kernel": 2 * E0,
    "rope_fwd_kernel": (2 + 2) * ROWS * 1088, "rope_bwd_kernel": (4 + 2) * ROWS * 1152, "colsum_bf16_kernel": 2 * ROWS * 1024,
    "cast_copy_kernel": (2 + 2) * ROWS * 1024,
    "attn_delta_kernel": (2 + 2 + 4) * ROWS * 1024,      # out + dout read, dq zero-filled (dk / dv: + 1 %)
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
idx = {n: i for i, n in enumerate(h)}
per_id = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) != len(h):
        continue
    d = per_id.setdefault(r[idx["ID"]], {"name": re.sub(r"<.*", "", re.sub(r"\(.*", "", r[idx["Kernel Name"]])).replace("void ", "").replace("ofx::", "").strip()})
    v = float(r[idx["Metric Value"]].replace(",", "")) * UNIT.get(r[idx["Metric Unit"]], 1.0)
    d[r[idx["Metric Name"]]] = v
agg = collections.defaultdict(lambda: {"launches": 0, "us": 0.0, "bytes": 0.0, "largest": None})
for d in per_id.values():
    t = d.get("gpu__time_duration.sum", 0.0)
    b = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    a = agg[d["name"]]
    a["launches"] += 1
    a["us"] += t
    a["bytes"] += b
    if a["largest"] is None or b > a["largest"][1]:
        a["largest"] = (t, b, d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0))
out = []
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    if any(s in name for s in ("gemm_kernel", "attn_fwd", "attn_bwd")) or a["us"] <= 0:
        continue
    t, b, br, bw = a["largest"]
    out.append({
        "kernel": name, "launches": a["launches"], "total_us": round(a["us"], 1), "total_dram_mb": round(a["bytes"] / 1e6, 1),
        "gbs_all_launches": round(a["bytes"] / a["us"] / 1e3, 1), "frac_of_peak_all": round(a["bytes"] / a["us"] / 1e3 / peak, 3),
        "largest_launch": {"us": round(t, 2), "dram_read_mb": round(br / 1e6, 2), "dram_write_mb": round(bw / 1e6, 2),
                           "gbs": round(b / t / 1e3, 1) if t > 0 else None, "frac_of_peak": round(b / t / 1e3 / peak, 3) if t > 0 else None},
        "algorithmic_mb_largest_launch": round(ALGO[name], 1) if name in ALGO else None,
        "algorithmic_gbs_largest_launch": round(ALGO[name] / t * 1e3, 1) if name in ALGO and t > 0 else None,
        "algorithmic_frac_of_peak": round(ALGO[name] / t * 1e3 / peak, 3) if name in ALGO and t > 0 else None,
    })
tot_us = sum(o["total_us"] for o in out)
print(json.dumps({"source": Path(sys.argv[1]).name, "hbm_peak_gbs": peak, "how": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
                  "dram__bytes_write.sum --clock-control none over one CUDA-graph replay of the CFG-L B4 N4096 micro-step (cold cache per launch)",
                  "caveat": "ncu flushes the caches before each launch and stops the clock at kernel exit: reads are cold, writes still in the 126 MB "
                            "L2 are not in dram__bytes_write; algorithmic_* (header bytes / measured time of the largest launch) is the fairer figure",
                  "non_tensor_total_ms": round(tot_us / 1e3, 2), "kernels": out}, indent=1))
