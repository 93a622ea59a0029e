"""Summarise an `ncu --set full` capture (.ncu-rep) into a small JSON for profiles/: per launch the duration, tensor-pipe activity,
issue activity, DRAM bytes, L2 / shared-memory pressure, registers and the top stall reasons.

    python tools/ncu_rep_summary.py capture.ncu-rep [note] > profiles/rNN_xxx.json
"""
import csv
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_cycles_active_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_cycles_active_pct_of_active",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed": "issue_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "lts__t_sectors.avg.pct_of_peak_sustained_elapsed": "lts_throughput_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "shared_wavefronts",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__registers_per_thread": "registers",
    "launch__occupancy_limit_shared_mem": "ctas_per_sm_by_smem",
    "smsp__warps_active.avg.per_cycle_active": "warps_active_per_scheduler",
}


def main() -> None:
    rep = sys.argv[1]
    note = sys.argv[2] if len(sys.argv) > 2 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    head, units = rows[0], rows[1]
    idx = {n: i for i, n in enumerate(head)}
    stall_cols = [(n, i) for n, i in idx.items() if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio")]
    out = []
    for r in rows[2:]:
        if len(r) != len(head):
            continue
        d = {"kernel": r[idx["Kernel Name"]].split("(")[0].replace("ofx::", "").replace("void ", "").strip()[:80]}
        for k, name in KEYS.items():
            if k in idx:
                try:
                    d[name] = float(r[idx[k]].replace(",", ""))
                except ValueError:
                    continue
                u = units[idx[k]]
                if name in ("duration", "dram_read", "dram_write"):
                    d[name + "_unit"] = u
        stalls = []
        for n, i in stall_cols:
            try:
                stalls.append((float(r[i].replace(",", "")), n[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
        d["top_stalls_per_issue"] = {n: round(v, 2) for v, n in sorted(stalls, reverse=True)[:4]}
        out.append(d)
    print(json.dumps({"capture": rep.split("/")[-1], "note": note, "launches": out}, indent=1))


if __name__ == "__main__":
    main()
