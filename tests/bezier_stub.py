"""Stand-in for the `bezier` package (absent from this image) so that the REFERENCE's own decode path
(`osu_fusion/library/osu/data/{decode,fit_bezier,hit}.py`, `library/osu/{beatmap,sliders}.py`) imports and runs here: the reference uses
exactly three things of it — `Curve.from_nodes(nodes)`, `.evaluate_multi(s)` and `.length` (fit_bezier.py:16,21,26,47; sliders.py).
Test infrastructure only (tests/test_decode_cpu.py, oracle/make_golden_decode.py).  Numerics: Bernstein-basis evaluation (powers by repeated multiplication, terms added in index
order) and a 64-point Gauss-Legendre arc length summed with math.fsum — the published definitions, written so that the result does not
depend on SIMD width; the real package's floating-point path is NOT reproduced ("parity unpinned" for those two calls), everything around
them is the reference's own code."""
import math
import sys
import types

import numpy as np

_GL_X, _GL_W = np.polynomial.legendre.leggauss(64)


class Curve:
    def __init__(self, nodes, degree=None):
        self.nodes = np.asarray(nodes, dtype=float)           # (dim, n_points)
        self.degree = self.nodes.shape[1] - 1 if degree is None else degree

    @classmethod
    def from_nodes(cls, nodes):
        return cls(nodes)

    def evaluate_multi(self, s):
        s = np.asarray(s, dtype=float)
        n = self.degree
        up, down = [np.ones_like(s)], [np.ones_like(s)]
        for _ in range(n):
            up.append(up[-1] * s)
            down.append(down[-1] * (1.0 - s))
        out = None
        for i in range(n + 1):
            term = self.nodes[:, i][:, None] * (math.comb(n, i) * down[n - i] * up[i])[None, :]
            out = term if out is None else out + term
        return out                                                                                 # (dim, len(s))

    @property
    def length(self):
        n = self.degree
        if n < 1:
            return 0.0
        hod = n * (self.nodes[:, 1:] - self.nodes[:, :-1])
        d = Curve(hod).evaluate_multi(0.5 * (_GL_X + 1.0))
        speed = np.sqrt(d[0] * d[0] + d[1] * d[1])
        return 0.5 * math.fsum((_GL_W * speed).tolist())


def install():
    if "bezier" not in sys.modules:
        m = types.ModuleType("bezier")
        m.Curve = Curve
        sys.modules["bezier"] = m
