"""Drop-in for `pawlikMorphLSST/asymmetry.py` (citations into that file)."""
import numpy as np
import torch

from . import _abi, _single

__all__ = ["minapix", "calcA", "calculateAsymmetries"]


def minapix(image, segmap, apermask, starMask=None):
    """asymmetry.py:50-129 — the pixel [col, row], among the brightest pixels holding 20 % of the
    flux, about which the 180-degree asymmetry inside the re-centred aperture is smallest.
    Ties in the brightness order are canonical (value descending, flat index descending)."""
    star = _single.star_tensor(starMask, np.shape(image))
    if star is not None:                      # asymmetry.py:82-86: imageMask = image * segmap * starMask
        image = (torch.from_numpy(np.asarray(image, dtype=np.float64)).to(star.device) * star).cpu().numpy()
    row, _ = _single.run(image, flags=0, segmap_in=np.asarray(segmap)[None], apermask_in=np.asarray(apermask)[None])
    st = int(row["status"])
    if st == _abi.PM_ST_NO_CANDIDATE or st == _abi.PM_ST_EMPTY_SEGMAP:
        raise ValueError("attempt to get argmin of an empty sequence")   # asymmetry.py:127 upstream
    return [int(row["apix_col"]), int(row["apix_row"])]


def calcA(img, segmap, apermask, centroid, angle, starMask=None, noisecorrect=False):
    """asymmetry.py:132-257 — [A, Abgr] for a rotation by `angle` (180 or 90 degrees) about `centroid`
    = [col, row], inside the (un-recentred) aperture; with `noisecorrect` the background asymmetry of
    asymmetry.py:210-255 is subtracted.  As upstream, shape asymmetry is calcA(segmap, segmap, ...)."""
    if float(angle) not in (180.0, 90.0):
        raise NotImplementedError("only the angles the pipeline uses (180 and 90 degrees, main.py:470-478) are supported")
    flags = _abi.PM_DO_A
    if float(angle) == 90.0:
        flags |= _abi.PM_A_ANGLE90
    if not noisecorrect:
        flags |= _abi.PM_NO_NOISECORRECT
    cen = _single.int_centroid(centroid)
    star = _single.star_tensor(starMask, np.shape(img))
    seg_used = np.asarray(segmap)
    if star is not None:
        # asymmetry.py:186-196: starMask *= rotate(starMask, angle, cval=1); img *= starMask; with noisecorrect the
        # caller's segmap is multiplied in place too (:219) — reproduced, the kernel then sees masked inputs
        sm = _single.rotated_mask(star, angle, int(cen[0]), int(cen[1]))
        img = (torch.from_numpy(np.asarray(img, dtype=np.float64)).to(sm.device) * sm).cpu().numpy()
        if noisecorrect:
            segmap *= sm.cpu().numpy()
            seg_used = segmap
    row, _ = _single.run(img, flags=flags, segmap_in=np.asarray(seg_used)[None], apermask_in=np.asarray(apermask)[None],
                         centroid_in=cen[None])
    if int(row["status"]) == _abi.PM_ST_EMPTY_SEGMAP:
        raise ValueError("empty segmap")
    A, Abgr = float(row["A"]), float(row["Abgr"])
    if not noisecorrect:
        Abgr = 0
    elif Abgr == -99.0:
        Abgr = -99
    return [A, Abgr]


def calculateAsymmetries(image, segmap):
    """asymmetry.py:13-47 — (A, As, As90) with rmax, aperture and centre computed on the way."""
    row, _ = _single.run(image, flags=_abi.PM_DO_A | _abi.PM_DO_AS, segmap_in=np.asarray(segmap)[None])
    if int(row["status"]) != _abi.PM_ST_OK:
        raise ValueError(f"calculateAsymmetries failed with status {int(row['status'])}")
    return float(row["A"]), float(row["As"]), float(row["As90"])
