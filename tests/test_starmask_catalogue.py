"""The star-catalogue helpers (imageutils.maskstarsPSF and friends, objectMasker.objectOccluded) against vectors made by
executing the reference's own function source (oracle/make_starmask_golden.py; the sky -> pixel step is this repository's
TAN restatement on both sides)."""
import json
import os

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(REPO, "tests", "golden", "starmask_catalogue.npz"))
M = json.load(open(os.path.join(REPO, "tests", "golden", "starmask_catalogue.json")))


def test_circle_masks_and_radii():
    import pawlikmorph_lsst_b200 as pm
    for k, (shape, centre, r) in enumerate(M["circles"]):
        assert np.array_equal(pm.imageutils._circle_mask(tuple(shape), tuple(centre), r), G[f"circle{k}"])
    got = np.array([pm.imageutils._calculateRadius(*a) for a in M["radii"]])
    assert np.array_equal(got, G["radii"])


def test_adaptive_radius():
    import pawlikmorph_lsst_b200 as pm
    import sys
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    from oracle.make_starmask_golden import inputs
    _, _, _, image, _ = inputs()
    n = image.shape[0]
    img2 = image.copy()
    yy, xx = np.mgrid[0:n, 0:n]
    for (x, y) in M["stars"]:
        img2 += 4000.0 * np.exp(-((xx - x) ** 2 + (yy - y) ** 2) / (2 * 6.0 ** 2))
    mask = np.zeros((n, n), dtype=bool)
    for (x, y) in M["stars"]:
        mask |= pm.imageutils._circle_mask((n, n), (y, x), 4.0)
    got = [pm.imageutils._calculateAdaptiveRadius(mask, list(p), img2, 100.0, 5.0) for p in M["stars"]]
    want = [int(v) for v in G["adaptive"]]
    assert all(g == w or (w == -1 and g == 0) for g, w in zip(got, want)), (got, want)     # -1: upstream's UnboundLocalError


def test_catalogue_search_occlusion_and_mask(tmp_path):
    import pawlikmorph_lsst_b200 as pm
    from oracle.make_starmask_golden import inputs
    header, radec, cat, image, segmap = inputs()
    path = tmp_path / "cat.csv"
    path.write_text("\n".join(",".join(str(v) for v in row) for row in cat) + "\n")
    stars = pm.objectMasker._findStars(str(path), radec[0], radec[1], False, False, False)
    every = pm.objectMasker._findStars(str(path), radec[0], radec[1], True, True, True)
    assert stars == [[r[0], r[1], r[2], r[3]] for r in M["findStars"]["stars"]]
    assert every == [[r[0], r[1], r[2], r[3]] for r in M["findStars"]["all"]]              # the zip of :197 stops at the stars
    flag, occ = pm.objectMasker.objectOccluded(segmap, radec, str(path), header)
    assert flag == M["occluded"][0] and occ == [[r[0], r[1], r[2], r[3]] for r in M["occluded"][1]]
    assert pm.objectMasker.objectOccluded(np.zeros_like(segmap), radec, str(path), header) == (False, [])
    mask = pm.imageutils.maskstarsPSF(image, stars, header, 1200.0, 5.0, adaptive=False)
    assert mask.dtype == G["starmask"].dtype and np.array_equal(mask, G["starmask"])
    none = pm.imageutils.maskstarsPSF(image[:8, :8], [], header, 1200.0, 5.0, adaptive=False)
    assert np.array_equal(none, G["starmask_none"]) and none.dtype == G["starmask_none"].dtype
    bad = dict(header)
    del bad["PSF_S1"]
    with pytest.raises(KeyError):                                                        # main.py:389-391 catches it
        pm.imageutils.maskstarsPSF(image, stars, bad, 1200.0)
    # adaptive=True (never used by main.py): the mask can only grow
    grown = pm.imageutils.maskstarsPSF(image, stars, header, 1200.0, 5.0, adaptive=True, sky_err=5.0)
    assert not np.any(grown & ~mask)


def test_subpixel_disc():
    import pawlikmorph_lsst_b200 as pm
    d = pm.imageutils._subpixel_disc((20, 20), (9.3, 10.6), 3.0)
    yy, xx = np.mgrid[0:20, 0:20]
    r = np.hypot(xx - 9.3, yy - 10.6)
    assert np.all(d[r <= 3.0 - 0.71] == 1) and np.all(d[r >= 3.0 + 0.71] == 0) and set(np.unique(d)) == {0.0, 1.0}
    assert pm.imageutils._subpixel_disc((20, 20), (-50.0, 3.0), 3.0).sum() == 0
