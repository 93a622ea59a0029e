"""Beatmap decode after sampling — the host step that follows `model.sample()` in `inference_gradio.py:128-165`.

    text = decode_beatmap(Metadata(...), signal, frame_times, bpm, allow_beat_snap)      # one `.osu` file as a string

Drop-in for `osu_fusion.library.osu.data.decode.decode_beatmap` (`decode.py:133-237`): same signature, same `Metadata` fields, same
`.osu` text.  `signal` is one sample of the denoiser output, `(6, N)` = HIT, SUSTAIN, SLIDER, COMBO, CURSOR_X, CURSOR_Y
(`encode.py:9-26`); a torch tensor on any device is accepted (the reference takes `generated.cpu().numpy()`).

This is host post-processing in the reference and stays host code here (numpy + the two scipy.signal primitives the reference itself
uses); what it replaces is the reference's dependency on the `bezier` package (absent from this image): cubic Bezier evaluation is
the Bernstein form with powers by repeated multiplication, arc length a 64-point Gauss-Legendre quadrature summed exactly (`math.fsum`)
— both independent of SIMD width, so the text is reproducible across hosts.

Structure (restated, not copied): the four hit signals are reduced to index lists first (`_edges`, `_runs`), objects are classified by a
small decision function, sliders are fitted by an explicit work-list version of Schneider's curve fit (`fit_bezier.py:50-99` recurses) and
the file is assembled at the end.  Quirks of the reference that change the text are kept and marked `# ref quirk`.
"""
from __future__ import annotations

import math
from dataclasses import asdict, dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
from scipy import signal as _sig

HIT, SUSTAIN, SLIDER, COMBO, CURSOR_X, CURSOR_Y = range(6)       # encode.py:9-22
BEAT_DIVISOR = 16                                                  # decode.py:13
PLAYFIELD = (512.0, 384.0)                                         # decode.py:142
MIN_OBJECT_FRAMES = 4                                              # decode.py:196,206
FIT_MAX_ERR = 50.0                                                 # decode.py:74
NEWTON_ROUNDS = 32                                                 # fit_bezier.py:80
MAX_SPLIT_DEPTH = 900                                              # just under CPython's default recursion limit (what bounds the reference)


@dataclass
class Metadata:                                                    # decode.py:19-28
    audio_filename: str
    title: str
    artist: str
    version: str
    cs: float
    ar: float
    od: float
    hp: float


# the file format is the contract (decode.py:31-61)
_OSU = "\n".join([
    "osu file format v14", "",
    "[General]", "AudioFilename: {audio_filename}", "AudioLeadIn: 0", "Mode: 0", "",
    "[Metadata]", "Title: {title}", "TitleUnicode: {title}", "Artist: {artist}", "ArtistUnicode: {artist}", "Creator: OsuFusion",
    "Version: {version}", "Tags: OsuFusion", "",
    "[Difficulty]", "HPDrainRate: {hp}", "CircleSize: {cs}", "OverallDifficulty: {od}", "ApproachRate: {ar}", "SliderMultiplier: 1",
    "SliderTickRate: 1", "",
    "[TimingPoints]", "{timing_points}", "",
    "[HitObjects]", "{hit_objects}", "",
])


# ------------------------------------------------------------------------------------------------ Bezier arithmetic
def _powers(x: np.ndarray, n: int) -> List[np.ndarray]:
    out = [np.ones_like(x)]
    for _ in range(n):
        out.append(out[-1] * x)
    return out


class _Basis:
    """Powers of t and 1 - t up to the cubic, shared by every evaluation at the same parameters (one fit iteration evaluates a cubic
    twice, a quadratic and a linear curve at the same `u`): the products are the ones `bezier_points` would form, so results are
    bit-identical — only the repeated work goes."""
    __slots__ = ("up", "down")

    def __init__(self, t: np.ndarray, n: int = 3) -> None:
        t = np.asarray(t, dtype=float)
        self.up, self.down = _powers(t, n), _powers(1.0 - t, n)


def bezier_points(ctrl: np.ndarray, t, basis: Optional[_Basis] = None) -> np.ndarray:
    """Points of the Bezier curve with control points `ctrl` (n + 1, dim) at parameters `t` (m,) -> (m, dim); Bernstein form, terms
    added in index order.  Computed as (dim, m) and returned transposed: downstream einsum reductions pick their summation order from
    the memory layout, and the reference's evaluation (`bezier.Curve(...).evaluate_multi(t).T`) hands them exactly this layout."""
    n = ctrl.shape[0] - 1
    if basis is None:
        basis = _Basis(t, n)
    up, down = basis.up, basis.down
    acc = None
    for i in range(n + 1):
        term = ctrl[i][:, None] * (math.comb(n, i) * down[n - i] * up[i])[None, :]
        acc = term if acc is None else acc + term
    return acc.T


def _hodograph(ctrl: np.ndarray) -> np.ndarray:
    return ctrl.shape[0] * (ctrl[1:] - ctrl[:-1])                  # fit_bezier.py:10-11 (scaled by the point count, as there)


_GL_X, _GL_W = np.polynomial.legendre.leggauss(64)


def bezier_length(ctrl: np.ndarray) -> float:
    """Arc length: integral of |B'(t)| over [0, 1] (B' = degree * forward differences), 64-point Gauss-Legendre."""
    n = ctrl.shape[0] - 1
    if n < 1:
        return 0.0
    d = bezier_points(n * (ctrl[1:] - ctrl[:-1]), 0.5 * (_GL_X + 1.0))
    speed = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])
    return 0.5 * math.fsum((_GL_W * speed).tolist())


# ------------------------------------------------------------------------------------------------ curve fitting (fit_bezier.py)
def _unit(v: np.ndarray) -> np.ndarray:
    m = math.sqrt(float(np.dot(v, v)))
    return v if m < np.finfo(float).eps else v / m


def _end_tangent(pts: np.ndarray, left: bool) -> np.ndarray:
    """Weighted direction of the first / last few chords (geometric weights 2^-1 .. 2^-n, n <= 5, normalised; fit_bezier.py:62-72)."""
    n = min(5, len(pts) - 2)
    w = (2.0 ** -np.arange(1, n + 1)) / (1 - 2.0 ** -n) * (2.0 - 1)
    vecs = (pts[2:2 + n] - pts[1]) if left else (pts[-3:-3 - n:-1] - pts[-2])
    return _unit(np.einsum("np,n->p", vecs, w))


def _max_sq_error(ctrl: np.ndarray, pts: np.ndarray, u: np.ndarray, basis: Optional[_Basis] = None, on_curve=None) -> Tuple[float, int]:
    q = bezier_points(ctrl, u, basis) if on_curve is None else on_curve
    e = ((q - pts) ** 2).sum(-1)
    k = int(e.argmax())
    return float(e[k]), k


def _least_squares_cubic(pts: np.ndarray, u: np.ndarray, tl: np.ndarray, tr: np.ndarray, basis: Optional[_Basis] = None) -> np.ndarray:
    """Cubic with end points on the data and inner control points on the end tangents, distances from the 2x2 normal equations
    (fit_bezier.py:102-150)."""
    ctrl = np.array([pts[0], pts[0], pts[-1], pts[-1]])
    a = (3 * (1 - u) * u * np.array([1 - u, u])).T[..., None] * np.array([tl, tr])
    c = np.einsum("lix,ljx->ij", a, a)
    x = np.einsum("lix,lx->i", a, pts - bezier_points(ctrl, u, basis))
    det = c[0][0] * c[1][1] - c[1][0] * c[0][1]
    al = 0.0 if abs(det) < 1e-5 else (x[0] * c[1][1] - x[1] * c[0][1]) / det
    ar = 0.0 if abs(det) < 1e-5 else (c[0][0] * x[1] - c[1][0] * x[0]) / det
    chord = np.linalg.norm(pts[0] - pts[-1])
    if al < 1e-6 * chord or ar < 1e-6 * chord:       # Wu / Barsky fallback
        al = ar = chord / 3.0
    ctrl[1] += tl * al
    ctrl[2] += tr * ar
    return ctrl


def _reparameterise(ctrl: np.ndarray, pts: np.ndarray, u: np.ndarray, basis: Optional[_Basis] = None, on_curve=None) -> np.ndarray:
    """One Newton step per point towards its foot point on the curve (fit_bezier.py:153-173)."""
    d = (bezier_points(ctrl, u, basis) if on_curve is None else on_curve) - pts
    h1 = _hodograph(ctrl)
    v = bezier_points(h1, u, basis)
    num = (d * v).sum(-1)
    den = (v ** 2 + d * bezier_points(_hodograph(h1), u, basis)).sum(-1)
    return u - np.divide(num, den, out=np.zeros_like(num), where=den != 0)


def fit_curve(points: np.ndarray, max_err: float = FIT_MAX_ERR) -> List[np.ndarray]:
    """Piecewise Bezier fit of a polyline: list of control-point arrays (2 points = a line, 4 = a cubic), in path order."""
    done: List[np.ndarray] = []
    work = [(points, None, None, 0)]                 # depth-first, left piece first: same order as the reference's recursion
    while work:
        pts, tl, tr, depth = work.pop()
        if depth > MAX_SPLIT_DEPTH:
            # a degenerate polyline (e.g. a cursor that does not move during a slider: zero path length -> NaN parameters) never
            # converges; the reference's recursive fit dies with RecursionError there — same exception type, instead of spinning
            raise RecursionError("curve fit does not converge (degenerate slider path); fit_bezier.py:96-98 recurses without bound here")
        if len(pts) < 2:
            continue
        if tl is None:
            tl = _end_tangent(pts, True)
        if tr is None:
            tr = _end_tangent(pts, False)
        if len(pts) == 2:
            done.append(pts)
            continue
        u = np.cumsum(np.linalg.norm(pts[1:] - pts[:-1], axis=1))
        u = np.pad(u, (1, 0)) / u[-1]
        piece = None
        for _ in range(NEWTON_ROUNDS):
            basis = _Basis(u)                        # powers of u, 1 - u: shared by the five curve evaluations of this round
            ctrl = _least_squares_cubic(pts, u, tl, tr, basis)
            on_curve = bezier_points(ctrl, u, basis)
            err, worst = _max_sq_error(ctrl, pts, u, on_curve=on_curve)
            if err < max_err:
                ends = ctrl[[0, -1]]
                piece = ends if _max_sq_error(ends, pts, u, basis)[0] < max_err else ctrl
                break
            u = _reparameterise(ctrl, pts, u, basis, on_curve)
        if piece is not None:
            done.append(piece)
            continue
        mid = _unit(pts[worst - 1] - pts[worst + 1])
        work.append((pts[worst:], -mid, tr, depth + 1))
        work.append((pts[:worst + 1], tl, mid, depth + 1))
    return done


# ------------------------------------------------------------------------------------------------ signals -> index lists
def _edges(level: np.ndarray) -> List[int]:
    """Frames where a flip signal changes state (hit.py:23-27: peaks of +-gradient above 0.5)."""
    g = np.gradient(level)
    up = _sig.find_peaks(g, height=0.5)[0]
    down = _sig.find_peaks(-g, height=0.5)[0]
    return sorted(up.tolist() + down.tolist())


def _runs(level: np.ndarray) -> List[Tuple[int, int]]:
    """(start, end) frame pairs of the positive runs of a hold signal (hit.py:50-68): a run starts at the last non-positive frame
    before it and ends at its last positive frame; ends that do not come after their start are skipped."""
    pos = level > 0
    starts = np.flatnonzero(~pos[:-1] & pos[1:]).tolist()
    ends = np.flatnonzero(pos[:-1] & ~pos[1:]).tolist()
    pairs, j = [], 0
    for s in starts:
        while j < len(ends) and ends[j] <= s:
            j += 1
        if j == len(ends):
            break
        pairs.append((s, ends[j]))
        j += 1
    return pairs


# ------------------------------------------------------------------------------------------------ timing
def _phase_timing(hit_times: np.ndarray, beat_len: float) -> Tuple[float, float]:
    """(offset, beat length): the most populated of 100 phase bins (decode.py:81-85)."""
    hist, edges = np.histogram(hit_times % beat_len, bins=100, range=(0, beat_len))
    return edges[np.argmax(hist)], beat_len


def estimate_timing(hit_times: np.ndarray, allow_beat_snap: bool, verbose: bool = True) -> Tuple[bool, float, float]:
    """(snap?, offset, beat length) from the hit times alone (decode.py:88-123): autocorrelation peak of the inter-onset intervals
    inside 1..300 BPM, refined over +-5 % by the sharpest phase histogram."""
    fallback = (False, 0, 60000 / 200)
    if not allow_beat_snap:
        return fallback
    gaps = np.diff(hit_times)
    ac = _sig.correlate(gaps, gaps, mode="full")
    ac = ac[len(ac) // 2:]
    periods = 60000 / np.arange(1, 300 + 1, 1)
    peaks, _ = _sig.find_peaks(ac, distance=periods.min())
    peaks = peaks[(periods.min() * 0.95 <= peaks) & (peaks <= periods.max() * 1.05)]
    if len(peaks) == 0:
        if verbose:
            print("Warning: no valid BPM found within the range, disabling beat snap")
        return fallback
    bpm0 = 60000 / peaks[np.argmax(ac[peaks])]
    grid = np.linspace(bpm0 * 0.95, bpm0 * 1.05, 1000)
    score = np.zeros_like(grid)
    for i, bpm in enumerate(grid):
        bl = 60000 / bpm
        score[i] = np.max(np.histogram(hit_times % bl, bins=100, range=(0, bl))[0])
    off, bl = _phase_timing(hit_times, 60000 / grid[np.argmax(score)])
    return True, off, bl


def _snap(t: float, offset: float, beat_len: float) -> float:
    step = beat_len / BEAT_DIVISOR
    return round((t - offset) / step) * step + offset


# ------------------------------------------------------------------------------------------------ slider paths, optionally in parallel
def slider_path(points: np.ndarray) -> Tuple[List[np.ndarray], float]:
    """(control points rounded to whole osu! pixels, path length) of one slider's cursor samples (decode.py:64-78)."""
    ctrl_pts: List[np.ndarray] = []
    length = 0.0
    for seg in fit_curve(points):
        seg = seg.round()
        ctrl_pts.extend(seg)
        length += bezier_length(seg)
    return ctrl_pts, length


_POOLS: dict = {}


def _fit_many(jobs: Sequence[np.ndarray], workers: int) -> List[Tuple[List[np.ndarray], float]]:
    """Curve fits of all sliders of a song.  The fits are independent and are ~all of the decode time (tens of thousands of tiny
    numpy calls: 8.8 s for a 4-minute song with 230 sliders against 3.3 s for sampling it on a B200), so `workers > 1` spreads them
    over a persistent pool of host processes (spawned, numpy-only: safe next to an initialised CUDA context).  Each fit runs the same
    arithmetic wherever it runs: the text is identical for every `workers`."""
    if workers <= 1 or len(jobs) < 2 * workers:
        return [slider_path(j) for j in jobs]
    pool = _POOLS.get(workers)
    if pool is None:
        import multiprocessing
        from concurrent.futures import ProcessPoolExecutor
        pool = _POOLS[workers] = ProcessPoolExecutor(max_workers=workers, mp_context=multiprocessing.get_context("spawn"))
    # cursor.T slices are strided views: ship them in the reference's memory layout (the fit's reductions depend on it)
    return list(pool.map(_slider_path_job, [np.asarray(j).T.copy() for j in jobs], chunksize=max(1, len(jobs) // (4 * workers))))


def _slider_path_job(points_t: np.ndarray) -> Tuple[List[np.ndarray], float]:
    return slider_path(points_t.T)          # (2, m) C-contiguous -> the (m, 2) transposed view the serial path sees


def _object_kind(f: int, he: int, se: int) -> str:
    """circle / spinner / slider by the hold and slide extents that start on onset frame f (decode.py:191-209)."""
    if he == -1 or he - f < MIN_OBJECT_FRAMES:
        return "circle"
    if se == -1:
        return "spinner"
    if se - f < MIN_OBJECT_FRAMES:
        return "circle"
    return "slider"


# ------------------------------------------------------------------------------------------------ decode
def _as_numpy(x) -> np.ndarray:
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.asarray(x)


def decode_beatmap(metadata: Metadata, encoded_beatmap, frame_times, bpm: Optional[float], allow_beat_snap: bool = True,
                   verbose: bool = True, workers: int = 0) -> str:
    enc = _as_numpy(encoded_beatmap)
    frame_times = _as_numpy(frame_times)
    level = np.where(enc[[HIT, SUSTAIN, SLIDER, COMBO]] > 0, 1.0, -1.0)
    cursor = ((enc[[CURSOR_X, CURSOR_Y]] + 1) / 2) * np.array([[PLAYFIELD[0]], [PLAYFIELD[1]]])

    onsets = _edges(level[HIT])
    slot = np.full_like(frame_times, -1, dtype=int)           # frame -> index of the object that starts there
    for i, f in enumerate(onsets):
        slot[f] = i
    n_obj = len(onsets)
    combo = [False] * n_obj
    for f in _edges(level[COMBO]):
        combo[slot[f]] = True                                 # ref quirk: slot -1 (no onset on that frame) marks the LAST object
    hold_end, slide_end = [-1] * n_obj, [-1] * n_obj
    for dst, row in ((hold_end, SUSTAIN), (slide_end, SLIDER)):
        for s, e in _runs(level[row]):
            if slot[s] != -1:
                dst[slot[s]] = e

    hit_times = frame_times[onsets]
    if bpm is not None:
        snap = True
        offset, beat_len = _phase_timing(hit_times, 60000 / bpm)
    else:
        snap, offset, beat_len = estimate_timing(hit_times, allow_beat_snap, verbose)
    base_vel = 1.0 * 100 / beat_len
    timing_lines = [f"{offset},{beat_len},4,0,0,50,1,0"]
    object_lines: List[str] = []

    # every slider's curve fit first (optionally on `workers` host processes), then the objects in order
    spans = []
    for f, he, se in zip(onsets, hold_end, slide_end):
        if _object_kind(f, he, se) == "slider":
            slides = max(1, round((he - f) / (se - f)))
            spans.append((slides, round(f + (he - f) / slides)))
        else:
            spans.append(None)
    fits = iter(_fit_many([cursor.T[f:sp[1] + 1] for f, sp in zip(onsets, spans) if sp is not None], workers))

    for f, nc, he, se, span in zip(onsets, combo, hold_end, slide_end, spans):
        x, y = cursor[:, f].round().astype(int)
        t, u = frame_times[f], frame_times[he]                # he == -1 indexes the last frame: u is unused in that case
        if snap:
            t, u = _snap(t, offset, beat_len), _snap(u, offset, beat_len)
        cb = 4 if nc else 0
        circle = f"{x},{y},{t},{1 + cb},0,0:0:0:0:"
        kind = _object_kind(f, he, se)
        if kind == "circle":
            object_lines.append(circle)
            continue
        if kind == "spinner":
            object_lines.append(f"256,192,{t},{8 + cb},0,{u}")
            continue
        slides = span[0]
        ctrl_pts, length = next(fits)
        if length == 0:
            object_lines.append(circle)                       # ref quirk: the slider line below is emitted as well
        x1, y1 = ctrl_pts[0]
        path = "|".join(f"{px}:{py}" for px, py in ctrl_pts[1:])
        object_lines.append(f"{x1},{y1},{t},{2 + cb},0,B|{path},{slides},{length}")
        vel = length * slides / (u - t) / base_vel
        vel = 1 if vel == 0 else vel
        if (vel > 10 or vel < 0.1) and verbose:
            print(f"Warning: slider velocity {vel} is out of bounds, slider will not be good")
        timing_lines.append(f"{t},{-100 / vel},4,0,0,50,0,0")

    return _OSU.format(**asdict(metadata), timing_points="\n".join(timing_lines), hit_objects="\n".join(object_lines))


def decode_batch(metadata: Metadata, samples, frame_times, bpm: Optional[float], allow_beat_snap: bool = True,
                 verbose: bool = False, version_template: str = "{version_name} ({batch_number}/{batch_size})",
                 workers: int = 0) -> List[Tuple[str, str]]:
    """All samples of one `model.sample()` call -> [(version name, .osu text)] (the loop of inference_gradio.py:152-163); one
    device -> host copy for the whole batch."""
    arr = _as_numpy(samples)
    out = []
    base = metadata.version
    for i, sig in enumerate(arr):
        md = Metadata(**{**asdict(metadata), "version": version_template.format(version_name=base, batch_number=i + 1, batch_size=len(arr))})
        out.append((md.version, decode_beatmap(md, sig, frame_times, bpm, allow_beat_snap, verbose, workers)))
    return out
