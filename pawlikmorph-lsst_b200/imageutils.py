"""Image cleaning — `pawlikMorphLSST/imageutils.py:42-91` (`maskstarsSEG`) on the device, for batches.

maskstarsSEG upstream: sigma-clipped mean / std of the image (:62), `detect_threshold(image, 1.5)` (:67), `detect_sources`
with a 3 x 3 Gaussian kernel and npixels = 8 (:68-71); every segment whose bounding box does not contain the centre pixel is
an external source (:78-82) and is painted over with `np.random.normal(mean, std)` (:85-89).  The reference's generator is
UNSEEDED, which makes its output irreproducible; here the normal deviate of (seed, cutout index, pixel index) comes from a
counter-based generator (Philox4x32-10 + Box-Muller), the same on the device and in the CPU checker.
`maskstarsPSF` and the catalogue helpers (:94-333) belong to the star-catalogue path and are not covered.
"""
import numpy as np
import torch

from . import _abi
from ._lib import lib


def _prep(images):
    if not torch.is_tensor(images) or not images.is_cuda:
        raise TypeError("needs a CUDA tensor (there is no CPU path)")
    if images.dtype not in (torch.float32, torch.float64) or images.ndim != 3 or images.shape[1] != images.shape[2]:
        raise ValueError("images must be float32 / float64 [B, n, n]")
    return images.contiguous()


def _run(images, threshold, npixels, mean, std, seed, want_labels, want_clean, stream):
    B, n = int(images.shape[0]), int(images.shape[1])
    dev = images.device
    L = lib()
    with torch.cuda.device(dev):
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        with torch.cuda.stream(st):
            labels = torch.empty((B, n, n), dtype=torch.int32, device=dev) if want_labels else None
            clean = torch.empty_like(images) if want_clean else None
            info = torch.empty((B, _abi.PM_STAR_COLS), dtype=torch.float64, device=dev)
            wsb = L.pm_maskstars_workspace_bytes(B, n)
            ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev)
            p = lambda t: None if t is None else t.data_ptr()
            rc = L.pm_maskstars_batch(images.data_ptr(), _abi.PM_F32 if images.dtype == torch.float32 else _abi.PM_F64, B, n,
                                      threshold.data_ptr(), int(npixels), p(mean), p(std), int(seed) & 0xFFFFFFFFFFFFFFFF, p(labels), p(clean),
                                      info.data_ptr(), ws.data_ptr(), wsb, st.cuda_stream)
            _abi.check(L, rc, "pm_maskstars_batch")
            for t in (images, threshold, mean, std, ws):
                if t is not None:
                    t.record_stream(st)
    return dict(labels=labels, clean=clean, info=info)


def detect_sources_batch(images, nsigma=1.5, npixels=8, want_labels=False, stream=None):
    """photutils.detect_threshold + detect_sources as skyBackground.py:104-110 / imageutils.py:66-71 call them.
    Returns dict(labels: CUDA int32 [B, n, n] or None, info: CUDA float64 [B, 4] — column 0 = number of segments)."""
    from .skyBackground import clipped_stats_batch
    images = _prep(images)
    cs = clipped_stats_batch(images, 3.0, None, stream=stream)
    thr = (cs[:, 6] + float(nsigma) * cs[:, 8]).contiguous()
    return _run(images, thr, npixels, None, None, 0, want_labels, False, stream)


def maskstars_batch(images, seed=0, nsigma=1.5, npixels=8, want_labels=False, stream=None):
    """imageutils.maskstarsSEG for a batch: dict(clean: CUDA tensor like `images`, labels, info [B, 4] = segments kept,
    segments cleaned, pixels replaced, status)."""
    from .skyBackground import clipped_stats_batch
    images = _prep(images)
    c5 = clipped_stats_batch(images, 3.0, 5, stream=stream)              # imageutils.py:62
    cn = clipped_stats_batch(images, 3.0, None, stream=stream)           # detect_threshold
    thr = (cn[:, 6] + float(nsigma) * cn[:, 8]).contiguous()
    return _run(images, thr, npixels, c5[:, 6].contiguous(), c5[:, 8].contiguous(), seed, want_labels, True, stream)


def maskstarsSEG(image, seed=0, device="cuda:0"):
    """imageutils.maskstarsSEG (imageutils.py:42-91) for one image: the cleaned image as a float64 numpy array."""
    x = torch.from_numpy(np.ascontiguousarray(image, dtype=np.float64))[None].to(device)
    return maskstars_batch(x, seed=seed)["clean"][0].cpu().numpy()
