"""ORACLE TOOLING (test infrastructure): golden `.osu` texts of the reference's OWN beatmap decoder.

Runs `/root/reference/osu_fusion/library/osu/data/decode.py` (`decode_beatmap`, :133-237, with `fit_bezier.py`, `hit.py`) on seeded
synthetic `(6, N)` signals and stores signals + texts in tests/golden/decode_ref.npz, so the pin of osufusion_b200/decode.py also holds
where /root/reference is absent.  The reference imports the `bezier` package, which this image lacks: tests/bezier_stub.py supplies the
three calls it uses (Bernstein evaluation, quadrature arc length) — everything else that runs is the reference's code.

    python -m oracle.make_golden_decode
"""
from __future__ import annotations

import contextlib
import io
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))

FRAME_MS = 1000 * 176 / 22050          # hop 176 at 22 050 Hz (scripts/dataset_creator.py:17-19)
META = dict(audio_filename="audio.mp3", title="Title", artist="Artist", version="Insane", cs=4.0, ar=9.0, od=8.0, hp=5.0)
CASES = [  # name, seed, frames, hits, rough cursor, bpm, allow_beat_snap
    ("smooth_auto_bpm", 4, 1500, 40, False, None, True),
    ("degenerate_slider_path", 1, 1500, 40, False, None, True),      # the cursor sits on the playfield border during a slider
    ("smooth_no_snap", 2, 1500, 40, False, None, False),
    ("smooth_given_bpm", 3, 1200, 36, False, 173.0, True),
    ("rough_cursor_splits", 28, 1200, 30, True, None, True),
]


def synth_signal(seed: int, n: int, hits: int, rough: bool) -> np.ndarray:
    """A plausible denoiser output: noisy +-0.8 flip / hold rows, a smoothed random-walk cursor (or white noise: many curve splits)."""
    rng = np.random.default_rng(seed)
    sig = rng.normal(0, 0.05, (6, n)) - 0.8
    onsets = np.sort(rng.choice(np.arange(5, n - 80), hits, replace=False))
    state = -1.0
    for h in onsets:
        state = -state
        sig[0, h:] = state * 0.8 + rng.normal(0, 0.05, n - h)
    state = -1.0
    for h in onsets[rng.random(hits) < 0.3]:
        state = -state
        sig[3, h:] = state * 0.7
    for h in onsets[rng.random(hits) < 0.4]:
        length = int(rng.integers(2, 70))
        sig[1, h:h + length] = 0.9
        if rng.random() < 0.75:
            sig[2, h:h + max(1, int(length / rng.integers(1, 4)))] = 0.9
    walk = np.cumsum(rng.normal(0, 0.02, (2, n)), axis=1)
    kern = np.ones(25) / 25
    sig[4:6] = np.clip(np.stack([np.convolve(walk[i], kern, "same") for i in range(2)]), -1, 1)
    if rough:
        sig[4:6] = np.clip(rng.normal(0, 0.3, (2, n)), -1, 1)
    return sig


def main() -> None:
    import bezier_stub
    bezier_stub.install()
    sys.path.insert(0, "/root/reference")
    from osu_fusion.library.osu.data import decode as ref

    out = {"cases": json.dumps([c[0] for c in CASES])}
    for name, seed, n, hits, rough, bpm, snap in CASES:
        sig = synth_signal(seed, n, hits, rough)
        ft = np.arange(n) * FRAME_MS
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                text = ref.decode_beatmap(ref.Metadata(**META), sig, ft, bpm, snap, True)
        except RecursionError:
            text = "!RecursionError"        # fit_bezier.py:96-98 recurses without bound on a zero-length polyline: error behaviour is pinned too
        out[f"{name}.signal"] = sig
        out[f"{name}.params"] = json.dumps({"bpm": bpm, "allow_beat_snap": snap, "frames": n})
        out[f"{name}.text"] = text
        print(name, len(text.splitlines()), "lines,", text.count(",B|"), "sliders,", sum(1 for l in text.splitlines() if l.startswith("256,192,")), "spinners")
    path = ROOT / "tests" / "golden" / "decode_ref.npz"
    np.savez_compressed(path, **out)
    print("wrote", path, path.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
