"""Parity assertions shared by the CPU-emulator and the GPU test modules.  `run` is one of
tests/runners.py.  The oracle (oracle/pm_oracle.py) and the golden vectors made by the unmodified
reference (tests/golden/*.npz) are the checkers."""
import numpy as np

import synth
from oracle import pm_oracle as O
from tests.golden_util import GoldenSet, compare_row

FIELDS = ("apix_col", "apix_row", "rmax", "r20", "r80", "C", "A", "Abgr", "S", "As", "As90", "gini", "m20",
          "object_edge", "status", "n_candidates")
# north star: floating-point statistics within 1e-6 relative of the reference; the kernels hold ~1e-12,
# the tests ask for 1e-9 so that a regression shows long before the contract is at risk.
RTOL = 1e-9
_sets = {}


def golden(name):
    if name not in _sets:
        _sets[name] = GoldenSet(name)
    return _sets[name]


def rowd(rows, i):
    return dict(zip(FIELDS, rows[i]))


def check_golden_case(run, name, i, flags=15, rtol=RTOL):
    g = golden(name)
    sky, err = g.sky(i)
    img = g.image(i)
    img32 = img.astype(np.float32)
    assert np.array_equal(img32.astype(np.float64), img.astype(np.float64)), "fixture is not float32-exact"
    res = run(img32[None], sky, err, g.fs(i), flags)
    assert np.array_equal(res["segmap"][0], g.segmap(i)), f"{name}[{i}]: pixelmap differs from the reference"
    errs = compare_row(rowd(res["rows"], 0), g.row(i), rtol, f"{name}[{i}] ")
    assert not errs, "\n".join(errs)


def check_against_oracle(run, images, sky, err, fs=3, flags=15, rtol=RTOL, dtype=np.float32, check_sorted=True):
    """images [B,n,n]; compares every output the ABI offers with the oracle, cutout by cutout."""
    x = np.ascontiguousarray(images.astype(dtype))
    res = run(x, sky, err, fs, flags)
    B, n = x.shape[0], x.shape[1]
    for i in range(B):
        want, seg, dbg = O.analyse(x[i].astype(np.float64), sky, err, fs, flags)
        assert np.array_equal(seg, res["segmap"][i]), f"[{i}] pixelmap differs"
        got = rowd(res["rows"], i)
        errs = compare_row(got, want, rtol, f"[{i}] ")
        if got["n_candidates"] != want["n_candidates"]:
            errs.append(f"[{i}] n_candidates {got['n_candidates']} != {want['n_candidates']}")
        assert not errs, "\n".join(errs)
        if want["status"] != 0 or not check_sorted:
            continue
        # sorted indices: bit-exact on the positive prefix (canonical order: value desc, index desc)
        m = (x[i].astype(np.float64) - sky) * seg
        vals, order = O.sorted_desc(m.ravel())
        npos = int((vals > 0).sum())
        assert int(res["n_positive"][i]) == npos
        assert np.array_equal(res["sorted_idx"][i][:npos], order[:npos].astype(np.int32)), f"[{i}] sort order differs"
        assert np.all(res["sorted_idx"][i][npos:] == -1)
        # aperture mask and the per-candidate asymmetries
        ap = O.aperpixmap(n, want["rmax"], 9, 0.1)
        assert np.array_equal(res["apermask"][i], ap.astype(np.uint8)), f"[{i}] aperture mask differs"
        a = np.asarray(dbg["a"])
        k = min(len(a), res["cand_a"].shape[1])
        np.testing.assert_allclose(res["cand_a"][i][:k], a[:k], rtol=1e-9, atol=0, err_msg=f"[{i}] candidate asymmetries")


def synth_batch(B, n, seed):
    return synth.make_batch(B, n, seed).numpy()


def check_small_radius_growth_curves(run, rtol=RTOL):
    """calcR20_R80 (casgm.py:124-155) for apertures of a few pixels: the exact circle/pixel overlap then sees
    chords that are long against the radius (the closed-form segment area instead of the series), pixels that
    straddle both axes, and circles inside one pixel."""
    n = 41
    yy, xx = np.mgrid[0:n, 0:n].astype(np.float64)
    rng = np.random.default_rng(7)
    img = (200.0 * np.exp(-((xx - 20.3) ** 2 + (yy - 19.6) ** 2) / (2 * 2.2 ** 2)) + rng.normal(0, 0.5, (n, n))).astype(np.float32)
    seg = np.ones((n, n), np.uint8)
    radii = [1.2, 1.9, 2.6, 4.3, 9.7, 17.0]
    B = len(radii)
    x = np.repeat(img[None], B, axis=0)
    res = run(x, 0.0, 0.0, 3, 4, segmap_in=np.repeat(seg[None], B, axis=0), centroid_in=np.tile([20.0, 20.0], (B, 1)),
              radius_in=np.asarray(radii))
    for i, rad in enumerate(radii):
        got = rowd(res["rows"], i)
        r20, r80 = O.calcR20_R80(img.astype(np.float64), seg.astype(np.float64), [20, 20], rad)
        for name, want in (("r20", r20), ("r80", r80)):
            assert abs(got[name] - want) <= rtol * abs(want) + 1e-13, f"radius {rad}: {name} {got[name]!r} != {want!r}"
