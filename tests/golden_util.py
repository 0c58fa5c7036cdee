"""Loader for tests/golden/*.npz (written by oracle/make_golden.py from the
unmodified reference run in the build container)."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EXACT = ("apix_col", "apix_row", "rmax", "object_edge", "status")
FLOATS = ("r20", "r80", "C", "A", "Abgr", "S", "As", "As90", "gini", "m20")


class GoldenSet:
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.fields = [str(f) for f in self.z["fields"]]
        self.rows = self.z["rows"]
        self.n = len(self.rows)

    def image(self, i):
        return self.z[f"img_{i}"]

    def segmap(self, i):
        n = int(self.z["sizes"][i])
        return np.unpackbits(self.z[f"seg_{i}"])[:n * n].reshape(n, n)

    def sky(self, i):
        return float(self.z["sky"][i]), float(self.z["sky_err"][i])

    def fs(self, i):
        return int(self.z["filter_size"][i])

    def row(self, i):
        return dict(zip(self.fields, self.rows[i]))


ATOL = 1e-13      # skimage's bilinear rotation leaks ~1e-16 into exactly-zero asymmetries (symmetric maps)


def compare_row(got, want, rtol, label=""):
    """got/want: dicts over the golden FIELDS.  Exact on EXACT, rtol (+ a 1e-13 absolute floor) on FLOATS."""
    errs = []
    for k in EXACT:
        if float(got[k]) != float(want[k]):
            errs.append(f"{label}{k}: got {got[k]!r} want {want[k]!r}")
    for k in FLOATS:
        g, w = float(got[k]), float(want[k])
        if g == w or (g != g and w != w):          # identical (incl. +-inf), or both NaN (e.g. log10 of a negative moment ratio, casgm.py:271)
            continue
        if w == -99.0 or g == -99.0:
            if g != w:
                errs.append(f"{label}{k}: got {g!r} want {w!r}")
        elif not (abs(g - w) <= rtol * max(abs(w), 1e-300) + ATOL):
            errs.append(f"{label}{k}: got {g!r} want {w!r} rel {abs(g - w) / max(abs(w), 1e-300):.3e}")
    return errs
