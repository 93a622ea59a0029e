"""bench.py's reference arm (the reference's CPU implementation = the oracle port, timed on the host cores) runs without a GPU and
prints ONE JSON line with the keys the driver's contract names (a bounded sample: size S, 256 frames, 1 step)."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_contract_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--size", "S", "--ref-frames", "256", "--steps", "1",
                        "--warmup", "1", "--ref-validate-frames", "512"], capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["metric"] == "denoiser fwd+bwd samples/s" and d["unit"] == "samples/s" and d["value"] > 0
    assert d["steps"] == 1 and d["warmup"] == 1 and d["ms_per_step"] > 0 and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "sample" in d["config"]
    assert d["validation"]["frames"] == 512 and d["validation"]["value"] > 0
