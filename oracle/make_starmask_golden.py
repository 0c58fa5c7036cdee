"""TEST INFRASTRUCTURE — golden vectors of the star-catalogue helpers, made by EXECUTING THE REFERENCE'S OWN FUNCTION SOURCE
(imageutils.py: _circle_mask, _calculateRadius, _calculateAdaptiveRadius, maskstarsPSF; objectMasker.py: _getObject,
_findStars, objectOccluded) as it lies under /root/reference, in a namespace whose third-party names are real where the
library is installed (numpy, pandas) and stand-ins where it is not:
  * astropy's WCS / SkyCoord / skycoord_to_pixel -> image.sky_to_pixel of this repository (TAN headers): the sky -> pixel
    step of these vectors is therefore NOT pinned against astropy (the restatement is checked elsewhere against the
    positions the reference's sample cut-outs were made at);
  * photutils' CircularAperture(...).to_mask(method="subpixel") -> not used: maskstarsPSF is run with adaptive=False, the only
    way main.py:387 calls it.
Run in the build container:  python oracle/make_starmask_golden.py   ->  tests/golden/starmask_catalogue.npz / .json"""
import ast
import json
import os
import sys
import tempfile
import warnings
from typing import List, Tuple  # noqa: F401  (annotations of the executed source)

import numpy as np
import pandas as pd

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PAWLIK_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)


def functions_of(path, names):
    src = open(path).read()
    tree = ast.parse(src)
    return "\n\n".join(ast.get_source_segment(src, node) for node in tree.body if isinstance(node, ast.FunctionDef) and node.name in names)


def namespace():
    import pawlikmorph_lsst_b200 as pm

    class _Pos:
        def __init__(self, ra, dec, unit=None):
            self.ra, self.dec = float(ra), float(dec)

    class _Utils:
        @staticmethod
        def skycoord_to_pixel(pos, wcs=None):
            return pm.image.sky_to_pixel(wcs, pos.ra, pos.dec)

    class _Wcs:
        utils = _Utils

        @staticmethod
        def WCS(header):
            return header

    class _Units:
        deg = 1.0

    ns = dict(np=np, pd=pd, sys=sys, warnings=warnings, AstropyWarning=Warning, wcs=_Wcs, SkyCoord=_Pos, units=_Units, List=List, Tuple=Tuple)
    exec(functions_of(os.path.join(REF, "pawlikMorphLSST", "imageutils.py"), {"_circle_mask", "_calculateRadius", "_calculateAdaptiveRadius", "maskstarsPSF"}), ns)
    exec(functions_of(os.path.join(REF, "pawlikMorphLSST", "objectMasker.py"), {"_getObject", "_findStars", "objectOccluded"}), ns)
    return ns


def inputs():
    """The seeded inputs (also rebuilt by the tests from the stored scalars)."""
    wcs = json.load(open(os.path.join(REPO, "tests", "golden", "sdss_wcs.json")))[0]
    header = dict(wcs["header"])
    header.update(PSF_S1=1.31, PSF_S2=2.94, EXPTIME=53.907456, PHT_AA=-24.05, PHT_KK=0.09, AIRMASS=1.19, SOFTBIAS=1000, PHT_B=1.2e-10)
    ra0, dec0 = wcs["ra"], wcs["dec"]
    rng = np.random.default_rng(42)
    n = header["NAXIS1"]
    cat = [("objID", "ra", "dec", "psfMag_r", "type")]
    kinds = ["STAR"] * 9 + ["GALAXY"] * 3 + ["COSMIC_RAY", "UNKNOWN"]
    for k, kind in enumerate(kinds):
        spread = 0.012 if k % 4 == 3 else 0.006
        cat.append((1000 + k, ra0 + rng.uniform(-spread, spread), dec0 + rng.uniform(-spread, spread), rng.uniform(14.0, 21.0), kind))
    image = (1100.0 + rng.normal(0, 5.0, (n, n))).astype(np.float64)
    segmap = np.zeros((n, n))
    segmap[n // 2 - 30:n // 2 + 30, n // 2 - 35:n // 2 + 35] = 1
    return header, (ra0, dec0), cat, image, segmap


def main():
    ns = namespace()
    header, radec, cat, image, segmap = inputs()
    out, meta = {}, {}
    # _circle_mask / _calculateRadius
    circles = [((24, 24), (10.0, 12.0), 5.0), ((24, 30), (3.2, 27.9), 6.5), ((16, 16), (-2.0, 8.0), 4.0), ((16, 16), (8, 8), 0.0)]
    for k, (shape, centre, r) in enumerate(circles):
        out[f"circle{k}"] = np.asarray(ns["_circle_mask"](shape, centre, r))
    meta["circles"] = [[list(s), list(c), r] for s, c, r in circles]
    radii = [(20.3, 15.2, 1.31), (21.9, 18.0, 2.94), (19.0, 19.0, 1.0)]
    out["radii"] = np.array([ns["_calculateRadius"](*a) for a in radii])
    meta["radii"] = [list(a) for a in radii]
    # _calculateAdaptiveRadius: a star-like column profile
    rng = np.random.default_rng(7)
    n = image.shape[0]
    img2 = image.copy()
    stars = [(60.4, 40.2), (150.7, 200.1), (20.0, 229.5), (100.0, 0.5)]
    for (x, y) in stars:
        yy, xx = np.mgrid[0:n, 0:n]
        img2 += 4000.0 * np.exp(-((xx - x) ** 2 + (yy - y) ** 2) / (2 * 6.0 ** 2))
    mask = np.zeros((n, n), dtype=bool)
    for (x, y) in stars:
        mask |= ns["_circle_mask"]((n, n), (y, x), 4.0)
    adaptive = []
    for pos in stars:
        try:
            adaptive.append(int(ns["_calculateAdaptiveRadius"](mask, list(pos), img2, 100.0, 5.0)))
        except UnboundLocalError:
            adaptive.append(-1)
    out["adaptive"] = np.array(adaptive)
    meta["stars"] = [list(p) for p in stars]
    # the catalogue
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "cat.csv")
        with open(path, "w") as fh:
            for row in cat:
                fh.write(",".join(str(v) for v in row) + "\n")
        found = {}
        for name, kw in (("stars", dict(galaxy=False, cosmicray=False, unknown=False)), ("all", dict(galaxy=True, cosmicray=True, unknown=True))):
            found[name] = ns["_findStars"](path, radec[0], radec[1], **kw)
        flag, occ = ns["objectOccluded"](segmap, radec, path, header)
    meta["catalogue"] = [list(r) for r in cat]
    meta["findStars"] = found
    meta["occluded"] = [bool(flag), occ]
    # maskstarsPSF the way main.py:387 calls it, on the stars found, and with no objects
    out["starmask"] = np.asarray(ns["maskstarsPSF"](image, found["stars"], header, 1100.0 + 100.0, 5.0, adaptive=False))
    out["starmask_none"] = np.asarray(ns["maskstarsPSF"](image[:8, :8], [], header, 1200.0, 5.0, adaptive=False))
    meta["header"] = header
    meta["radec"] = list(radec)
    np.savez_compressed(os.path.join(REPO, "tests", "golden", "starmask_catalogue.npz"), **out)
    json.dump(meta, open(os.path.join(REPO, "tests", "golden", "starmask_catalogue.json"), "w"), indent=1)
    print({k: (v.shape, int(np.asarray(v).sum())) for k, v in out.items()}, meta["occluded"][0], len(found["stars"]), len(found["all"]))


if __name__ == "__main__":
    main()
