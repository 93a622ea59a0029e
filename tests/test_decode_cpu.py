"""Beatmap decode after sampling: osufusion_b200/decode.py against the REFERENCE's own decoder.

Two arms, like tests/test_lora_reference_pin.py:
  * golden `.osu` texts produced by `/root/reference/osu_fusion/library/osu/data/decode.py` (oracle/make_golden_decode.py ->
    tests/golden/decode_ref.npz) — run anywhere; the comparison is STRING equality (hit objects, control points, slider lengths
    and velocities with every digit), and the reference's failure on a degenerate slider path (RecursionError) is pinned too;
  * the live reference on fresh random signals — in the build container only.
The reference needs the `bezier` package (absent): tests/bezier_stub.py supplies its three calls, so Bezier evaluation / arc length
themselves are "parity unpinned"; flip / extent decoding, timing estimation, the curve-fit control flow and the file format are the
reference's code.
"""
import contextlib
import io
import json
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from conftest import HAVE_REFERENCE
from oracle.make_golden_decode import FRAME_MS, META, synth_signal
from osufusion_b200 import decode as D

GOLD = Path(__file__).parent / "golden" / "decode_ref.npz"


def run(mod, sig, ft, bpm, snap):
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            return mod.decode_beatmap(mod.Metadata(**META), sig, ft, bpm, snap, True)
    except RecursionError:
        return "!RecursionError"


def test_golden_texts_of_the_reference_decoder():
    g = np.load(GOLD, allow_pickle=False)
    names = json.loads(str(g["cases"]))
    assert len(names) >= 5
    seen_slider = seen_spinner = seen_error = False
    for name in names:
        p = json.loads(str(g[f"{name}.params"]))
        sig, want = g[f"{name}.signal"], str(g[f"{name}.text"])
        got = run(D, sig, np.arange(p["frames"]) * FRAME_MS, p["bpm"], p["allow_beat_snap"])
        assert got == want, (name, [(a, b) for a, b in zip(got.splitlines(), want.splitlines()) if a != b][:3])
        seen_slider |= ",B|" in want
        seen_spinner |= any(line.startswith("256,192,") for line in want.splitlines())
        seen_error |= want == "!RecursionError"
    assert seen_slider and seen_spinner and seen_error


@pytest.fixture(scope="module")
def ref():
    if not HAVE_REFERENCE:
        pytest.skip("/root/reference is not mounted")
    import bezier_stub
    bezier_stub.install()
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
    from osu_fusion.library.osu.data import decode as ref_decode
    return ref_decode


def test_file_template_is_the_reference_format(ref):
    assert D._OSU == ref.map_template
    assert [f.name for f in D.Metadata.__dataclass_fields__.values()] == [f.name for f in ref.Metadata.__dataclass_fields__.values()]


@pytest.mark.parametrize("seed", range(100, 112))
def test_live_reference_on_random_signals(ref, seed):
    rough = seed % 4 == 0
    sig = synth_signal(seed, 1000, 28, rough)
    ft = np.arange(1000) * FRAME_MS
    for bpm, snap in ((None, True), (None, False), (150.0 + seed, True)):
        assert run(D, sig, ft, bpm, snap) == run(ref, sig, ft, bpm, snap), (seed, bpm, snap)


def test_index_decoders_match_reference_on_adversarial_rows(ref):
    """Adjacent flips, runs touching the borders, ends before starts: the hold pairing (hit.py:50-68) and the flip detector."""
    from osu_fusion.library.osu.data.hit import decode_extents, decode_flips
    rng = np.random.default_rng(7)
    for _ in range(300):
        n = int(rng.integers(4, 40))
        row = np.where(rng.random(n) < 0.5, 1.0, -1.0)
        starts, ends = decode_extents(row.copy())
        assert D._runs(row) == list(zip(starts, ends))
        assert D._edges(row) == decode_flips(row)


def test_bezier_arithmetic():
    ctrl = np.array([[0.0, 0.0], [1.0, 2.0], [3.0, 2.0], [4.0, 0.0]])
    t = np.array([0.0, 0.5, 1.0])
    np.testing.assert_allclose(D.bezier_points(ctrl, t), [[0, 0], [2.0, 1.5], [4, 0]], atol=1e-15)
    line = np.array([[1.0, 1.0], [4.0, 5.0]])
    assert D.bezier_length(line) == pytest.approx(5.0, abs=1e-12)
    # a cubic whose control points are collinear and evenly spaced is a straight segment traversed at constant speed
    assert D.bezier_length(np.array([[0.0, 0.0], [1.0, 0.0], [2.0, 0.0], [3.0, 0.0]])) == pytest.approx(3.0, abs=1e-12)
    # quarter circle approximated by the standard cubic (kappa = 0.5523): length within 3e-4 of pi/2
    k = 0.5522847498
    arc = np.array([[1.0, 0.0], [1.0, k], [k, 1.0], [0.0, 1.0]])
    assert D.bezier_length(arc) == pytest.approx(np.pi / 2, rel=3e-4)


def test_curve_fit_reproduces_a_cubic_and_a_line():
    ctrl = np.array([[10.0, 10.0], [120.0, 300.0], [380.0, 290.0], [500.0, 40.0]])
    pts = D.bezier_points(ctrl, np.linspace(0, 1, 40))
    fit = D.fit_curve(np.ascontiguousarray(pts))
    assert len(fit) == 1 and fit[0].shape == (4, 2)
    err = np.sqrt(((D.bezier_points(fit[0], np.linspace(0, 1, 400))[:, None, :] - pts[None, :, :]) ** 2).sum(-1)).min(0).max()
    assert err < 7.5          # max_err = 50 is a SQUARED distance: every sample lies within ~7 px of the fitted curve
    straight = np.stack([np.linspace(0, 300, 30), np.linspace(50, 200, 30)], 1)
    fit = D.fit_curve(straight)
    assert len(fit) == 1 and fit[0].shape == (2, 2)          # a cubic that a line explains as well collapses to its end points


def test_accepts_sampler_output_tensors_and_batches():
    sig = synth_signal(5, 800, 20, False)
    ft = np.arange(800) * FRAME_MS
    a = run(D, sig, ft, 180.0, True)
    b = run(D, torch.from_numpy(sig), torch.from_numpy(ft), 180.0, True)
    assert a == b
    batch = torch.from_numpy(np.stack([sig, synth_signal(6, 800, 20, False)]))
    out = D.decode_batch(D.Metadata(**META), batch, ft, 180.0)
    assert [v for v, _ in out] == ["Insane (1/2)", "Insane (2/2)"]
    assert out[0][1] == a.replace("Version: Insane", "Version: Insane (1/2)")


def test_parallel_slider_fits_give_the_same_text():
    """`workers > 1` spreads the curve fits of a song over host processes; every fit runs the same arithmetic wherever it runs, so the
    text is identical — including the failure on a degenerate slider path."""
    for seed, rough in ((21, False), (24, True)):
        sig = synth_signal(seed, 1500, 60, rough)
        ft = np.arange(1500) * FRAME_MS
        serial = run(D, sig, ft, None, True)
        assert serial.count(",B|") >= 4
        with contextlib.redirect_stdout(io.StringIO()):
            assert D.decode_beatmap(D.Metadata(**META), sig, ft, None, True, True, workers=2) == serial
    g = np.load(GOLD, allow_pickle=False)
    sig = g["degenerate_slider_path.signal"]
    with pytest.raises(RecursionError):
        D.decode_beatmap(D.Metadata(**META), sig, np.arange(sig.shape[1]) * FRAME_MS, None, True, False, workers=2)
