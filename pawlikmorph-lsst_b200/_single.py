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


def check_starmask(starMask):
    if starMask is None:
        return
    if not np.all(np.asarray(starMask) == 1):
        raise NotImplementedError("star masks other than all-ones are not supported yet (the catalogue path, "
                                  "main.py:385-401, is out of scope)")


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
