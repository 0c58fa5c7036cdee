"""Shared plumbing of the per-cutout (reference-signature) functions: one cutout = a batch of 1."""
import numpy as np
import torch

from . import _abi
from .batch import analyse_batch


def device():
    if not torch.cuda.is_available():
        raise RuntimeError("pawlikmorph-lsst_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def image_tensor(image):
    """float64 ndarray -> CUDA tensor [1, n, n]; float32 when that is lossless (the kernels then
    sort 32-bit keys), float64 otherwise."""
    a = np.asarray(image)
    if a.ndim != 2 or a.shape[0] != a.shape[1]:
        raise ValueError("expected a square 2-D image")
    if a.dtype != np.float32:
        a64 = a.astype(np.float64, copy=False)
        a32 = a64.astype(np.float32)
        a = a32 if np.array_equal(a32.astype(np.float64), a64) else a64
    return torch.from_numpy(np.ascontiguousarray(a)).to(device())[None]


def star_tensor(starMask, shape):
    """None / all-ones -> None; otherwise a float64 CUDA tensor of the mask."""
    if starMask is None:
        return None
    m = np.asarray(starMask)
    if m.shape != tuple(shape):
        raise ValueError("starMask must have the shape of the image")
    if np.all(m == 1):
        return None
    return torch.from_numpy(np.ascontiguousarray(m.astype(np.float64))).to(device())


def rotated_mask(star, angle, cx, cy):
    """starMask * rotate(starMask, angle, center=(cx, cy), cval=1) (asymmetry.py:186-191, pixmap.py:81-83) for
    the integer centres and the angles the pipeline uses: skimage's inverse-map rotation is then an index
    permutation (out-of-range sources take cval = 1).  Device tensor ops (plumbing): exact 0/1 arithmetic."""
    n = star.shape[0]
    r = torch.arange(n, device=star.device)[:, None].expand(n, n)
    c = torch.arange(n, device=star.device)[None, :].expand(n, n)
    if float(angle) == 180.0:
        sr, sc = 2 * cy - r, 2 * cx - c
    elif float(angle) == 90.0:
        sr, sc = cy + (c - cx), cx - (r - cy)
    else:
        raise NotImplementedError("only 180 and 90 degrees")
    ok = (sr >= 0) & (sr < n) & (sc >= 0) & (sc < n)
    rot = torch.ones_like(star)
    rot[ok] = star[sr[ok], sc[ok]]
    return star * rot


def int_centroid(centroid):
    c = np.asarray(centroid, dtype=np.float64).reshape(2)
    if not np.all(c == np.rint(c)):
        raise ValueError("centroid must be integer pixel indices [col, row] (what asymmetry.minapix returns)")
    return c


def run(image, *, sky=0.0, sky_err=0.0, filter_size=3, flags=0, **kw):
    res = analyse_batch(image_tensor(image), float(sky), float(sky_err), filter_size, flags, **kw)
    torch.cuda.current_stream().synchronize()
    row = res.row_dict(0)
    if int(row["status"]) == _abi.PM_ST_WORKSPACE:
        raise MemoryError("pawlik_b200: per-CTA scratch exhausted")
    return row, res
