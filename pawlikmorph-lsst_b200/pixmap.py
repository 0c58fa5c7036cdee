"""Drop-in for `pawlikMorphLSST/pixmap.py`: same names, arguments and error behaviour; the work is
done by the CUDA library.  Citations are into /root/reference/pawlikMorphLSST/pixmap.py."""
import numpy as np
import torch

from . import _abi, _single

__all__ = ["pixelmap", "calcMaskedFraction", "calcRmax", "checkPixelmapEdges"]


class _ImageError(AttributeError):
    """pixmap.py:17-23 — the reference's error class builds and raises an AttributeError."""

    def __init__(self, value="ERROR! Central pixel too faint"):
        super().__init__(value)
        self.value = value

    def __str__(self):
        return repr(self.value)


class _FilterSizeError(NameError, ValueError):
    """pixmap.py:160-162 — an even filter size reaches `sys.exit()` with `sys` not imported, i.e. a
    NameError upstream; kept catchable as both NameError and ValueError."""


def pixelmap(image, threshold, filterSize, starMask=None):
    """pixmap.py:130-213 — 8-connected object mask grown from the central pixel of the
    filterSize-mean-filtered, filterSize-shrunk image, nearest-neighbour upsampled.
    Returns a float64 array of 0./1.  Raises AttributeError when the central pixel is below
    `threshold` (pixmap.py:175-179)."""
    if filterSize % 2 == 0:
        raise _FilterSizeError("ERROR! Filter can not be of even size.")
    star = _single.star_tensor(starMask, np.shape(image))
    if star is not None:                      # pixmap.py:156-158: imageTmp = image * starMask, then the same path
        image = (torch.from_numpy(np.asarray(image, dtype=np.float64)).to(star.device) * star).cpu().numpy()
    row, res = _single.run(image, filter_size=int(filterSize), flags=_abi.PM_STOP_AFTER_PIXELMAP,
                           threshold_in=[float(threshold)], want_segmap=True)
    if int(row["status"]) == _abi.PM_ST_FAINT_CENTRE:
        raise _ImageError()
    return res.segmap[0].cpu().numpy().astype(np.float64)


def calcRmax(image, segmap):
    """pixmap.py:26-55 — distance from the brightest pixel of image*segmap to the farthest map pixel."""
    row, _ = _single.run(image, flags=_abi.PM_STOP_AFTER_RMAX, segmap_in=np.asarray(segmap)[None])
    if int(row["status"]) == _abi.PM_ST_EMPTY_SEGMAP:
        raise ValueError("zero-size array to reduction operation maximum which has no identity")  # pixmap.py:53
    return float(row["rmax"])


def checkPixelmapEdges(segmap):
    """pixmap.py:91-127 — True when the map touches any border of the image."""
    seg = np.asarray(segmap)
    row, _ = _single.run(np.zeros(seg.shape, np.float32), flags=_abi.PM_STOP_AFTER_PIXELMAP, segmap_in=seg[None])
    return bool(row["object_edge"])


def calcMaskedFraction(oldMask, starMask, cenpix):
    """pixmap.py:58-88 — fraction of the object's pixelmap covered by the (180-degree symmetrised) star mask:
    1 - sum(oldMask * starMask * rotate(starMask, 180, center=cenpix, cval=1)) / sum(oldMask).
    A masked pixel count: evaluated with device tensor ops (there is no kernel worth writing for it)."""
    dev = _single.device()
    star = torch.from_numpy(np.ascontiguousarray(np.asarray(starMask, dtype=np.float64))).to(dev)
    old = torch.from_numpy(np.ascontiguousarray(np.asarray(oldMask, dtype=np.float64))).to(dev)
    cen = _single.int_centroid(cenpix)
    sm = _single.rotated_mask(star, 180.0, int(cen[0]), int(cen[1]))
    return float(1.0 - (old * sm).sum() / old.sum())
