"""Drop-in for `pawlikMorphLSST/casgm.py` (citations into that file)."""
import numpy as np

from . import _abi, _single

__all__ = ["gini", "m20", "concentration", "smoothness", "calcR20_R80"]


def calcR20_R80(image, segmap, centroid, radius):
    """casgm.py:124-155 — radii of the circles about `centroid` = [col, row] that enclose 20 % and
    80 % of the exact-overlap aperture flux of image*segmap within `radius`."""
    cen = _single.int_centroid(centroid)
    row, _ = _single.run(image, flags=_abi.PM_DO_CAS, segmap_in=np.asarray(segmap)[None], centroid_in=cen[None],
                         radius_in=[float(radius)])
    if int(row["status"]) != _abi.PM_ST_OK:
        raise ValueError("no bracket for the enclosed-flux radius (casgm.py:55-76)")
    return float(row["r20"]), float(row["r80"])


def concentration(r20, r80):
    """casgm.py:158-184 — C = 5 log10(r80 / r20) (one scalar expression; evaluated on the host)."""
    return 5.0 * np.log10(r80 / r20)


def _cas_row(image, segmap):
    n = np.asarray(image).shape[0]
    row, _ = _single.run(image, flags=_abi.PM_DO_CAS, segmap_in=np.asarray(segmap)[None],
                         centroid_in=np.array([[n // 2, n // 2]], dtype=np.float64), radius_in=[1.0])
    if int(row["status"]) == _abi.PM_ST_EMPTY_SEGMAP:
        raise ValueError("empty segmap")
    return row


def gini(image, segmap):
    """casgm.py:187-218 — Gini index of the pixels inside the map (photutils.morphology.gini)."""
    return float(_cas_row(image, segmap)["gini"])


def m20(image, segmap):
    """casgm.py:221-273 — log10 of the normalised second moment of the brightest pixels holding 20 % of the flux."""
    return float(_cas_row(image, segmap)["m20"])


def smoothness(image, segmap, centroid, Rmax, r20, sky):
    """casgm.py:276-338 (+ :341-406) — clumpiness S inside the annulus r20..Rmax, corrected for the
    background smoothness.  `image` is the sky-subtracted image, `sky` the value used by the
    background-box test (casgm.py:385)."""
    cen = _single.int_centroid(centroid)
    if int(r20) < 1:
        raise RuntimeError("filter size must be positive")       # scipy.ndimage.uniform_filter(size=0) upstream
    row, _ = _single.run(image, flags=_abi.PM_DO_S, segmap_in=np.asarray(segmap)[None], centroid_in=cen[None],
                         radius_in=[float(Rmax)], r20_in=[float(r20)], smooth_sky_in=[float(sky)])
    return float(row["S"])
