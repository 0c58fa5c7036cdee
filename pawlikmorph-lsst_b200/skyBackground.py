"""Sky background from the sky region of a cutout — the statistics half of `pawlikMorphLSST/skyBackground.py`.

`skybgr` upstream = (1) a source-detection sanity check, (2) a 2-D Gaussian fit of the object (scipy curve_fit or
astropy's LevMar fitter, skyBackground.py:167-262, :306-395) giving `fwhms, theta`, (3) the statistics of the pixels
outside the ellipse of semi-axes 2 min(fwhms), 2 max(fwhms) (:134-164, :265-303).  Step (3) is the per-pixel work and
runs on the GPU here for a whole batch (`pm_sky_stats_batch`); (1) and (2) are not implemented (SURVEY §8f rank 1,
DESIGN §7) — the caller supplies `fwhms, theta` per cutout.
"""
import math

import numpy as np
import torch

from . import _abi
from ._lib import lib


class _SkyError(Exception):
    """skyBackground.py:20-27 (raised for a sky region of fewer than 100 pixels, :147-148)."""


def sky_stats_batch(images, fwhms, theta, stream=None):
    """images: CUDA tensor [B, n, n] float32 / float64 (sky still in); fwhms: [B, 2]; theta: [B] (radians).
    Returns a CUDA float64 tensor [B, 12] with columns `_abi.SKY_FIELDS` (sky = column 9, sky_err = column 10)."""
    if not torch.is_tensor(images) or not images.is_cuda:
        raise TypeError("sky_stats_batch needs a CUDA tensor (there is no CPU path)")
    if images.dtype not in (torch.float32, torch.float64) or images.ndim != 3 or images.shape[1] != images.shape[2]:
        raise ValueError("images must be float32 / float64 [B, n, n]")
    images = images.contiguous()
    B, n = int(images.shape[0]), int(images.shape[1])
    f = np.asarray(fwhms, dtype=np.float64).reshape(B, 2)
    t = np.broadcast_to(np.asarray(theta, dtype=np.float64), (B,))
    if B and not (np.all(f > 0) and np.all(np.isfinite(f)) and np.all(np.isfinite(t))):
        raise ValueError("fwhms must be positive and finite, theta finite")
    ell = np.empty((B, 4))
    ell[:, 0], ell[:, 1] = 2.0 * f.min(axis=1), 2.0 * f.max(axis=1)                      # skyBackground.py:292-293
    ell[:, 2], ell[:, 3] = [math.cos(v) for v in t], [math.sin(v) for v in t]             # libm, like photutils
    dev = images.device
    L = lib()
    with torch.cuda.device(dev):
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        with torch.cuda.stream(st):
            ell_d = torch.from_numpy(ell).to(dev)
            out = torch.empty((B, _abi.PM_SKY_COLS), dtype=torch.float64, device=dev)
            rc = L.pm_sky_stats_batch(images.data_ptr(), _abi.PM_F32 if images.dtype == torch.float32 else _abi.PM_F64, B, n,
                                      ell_d.data_ptr(), out.data_ptr(), st.cuda_stream)
            _abi.check(L, rc, "pm_sky_stats_batch")
            images.record_stream(st)
            ell_d.record_stream(st)
    return out


def skyFromRegion(img, fwhms, theta, device="cuda:0"):
    """(sky, sky_err) of one image given the fitted Gaussian: the tail of `_calcSkybgr` (skyBackground.py:134-164)."""
    x = torch.from_numpy(np.ascontiguousarray(img, dtype=np.float64))[None].to(device)
    row = sky_stats_batch(x, [fwhms], [theta])[0].cpu().numpy()
    if row[11] != 0:
        raise _SkyError(f"Error! Sky region too small {int(row[0])}")
    return float(row[9]), float(row[10])
