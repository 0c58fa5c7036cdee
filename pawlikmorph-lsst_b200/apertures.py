"""Drop-in for `pawlikMorphLSST/apertures.py` (citations into that file)."""
import ctypes as C

import numpy as np
import torch

from . import _abi, _single
from ._lib import lib

__all__ = ["aperpixmap", "distarr", "subdistarr", "apercentre"]


def aperpixmap(npix, rad, nsubpix, frac):
    """apertures.py:42-128 — binary disc centred on (npix//2, npix//2): a pixel is set when at least
    `frac` of its nsubpix**2 sub-samples lie within `rad`.  float64 array of 0./1."""
    npix = int(npix)
    dev = _single.device()
    r = torch.tensor([float(rad)], dtype=torch.float64, device=dev)
    out = torch.empty((1, npix, npix), dtype=torch.uint8, device=dev)
    L = lib()
    st = torch.cuda.current_stream(dev)
    rc = L.pm_aperpixmap_batch(npix, r.data_ptr(), 1, int(nsubpix), float(frac), out.data_ptr(), C.c_void_p(st.cuda_stream))
    _abi.check(L, rc, "pm_aperpixmap_batch")
    return out[0].cpu().numpy().astype(np.float64)


def distarr(npixx, npixy, cenpix):
    """apertures.py:131-165 — dist[i, j] = sqrt((i - c0)^2 + (j - c1)^2).  The kernels never
    materialise this array (it is folded into the rmax reduction); this helper builds it with
    device tensor ops for API completeness."""
    dev = _single.device()
    i = torch.arange(int(npixx), device=dev, dtype=torch.float64)[:, None] - float(cenpix[0])
    j = torch.arange(int(npixy), device=dev, dtype=torch.float64)[None, :] - float(cenpix[1])
    return torch.sqrt(i * i + j * j).cpu().numpy()


def subdistarr(npix, nsubpix, cenpix):
    """apertures.py:168-209 — oversampled distance grid with the reference's coordinate quirks
    (the y vector reuses the x one, :200).  Folded into the aperture kernel; device tensor ops here."""
    dev = _single.device()
    half = (int(npix) // 2) * int(nsubpix)
    # the coordinate vector on the host (a few thousand correctly rounded divisions: torch divides by a scalar through its
    # reciprocal), the (npix * nsubpix)^2 distance grid on the device
    neg = np.arange(half, dtype=np.float64) / float(nsubpix) - float(cenpix[0])
    v = torch.from_numpy(np.concatenate((neg, np.zeros(int(nsubpix)), -neg[::-1]))).to(dev)
    v2 = v * v
    return torch.sqrt(v2[:, None] + v2[None, :]).cpu().numpy()


def apercentre(apermask, pix):
    """apertures.py:212-240 — cyclic shift of the mask by pix - (n//2 + 1) on both axes."""
    dev = _single.device()
    m = torch.as_tensor(np.asarray(apermask), device=dev)
    n = m.shape[0]
    d0, d1 = int(pix[0]) - (n // 2 + 1), int(pix[1]) - (n // 2 + 1)
    return torch.roll(torch.roll(m, d0, 0), d1, 1).cpu().numpy()
