"""Image cleaning — `pawlikMorphLSST/imageutils.py:42-91` (`maskstarsSEG`) on the device, for batches.

maskstarsSEG upstream: sigma-clipped mean / std of the image (:62), `detect_threshold(image, 1.5)` (:67), `detect_sources`
with a 3 x 3 Gaussian kernel and npixels = 8 (:68-71); every segment whose bounding box does not contain the centre pixel is
an external source (:78-82) and is painted over with `np.random.normal(mean, std)` (:85-89).  The reference's generator is
UNSEEDED, which makes its output irreproducible; here the normal deviate of (seed, cutout index, pixel index) comes from a
counter-based generator (Philox4x32-10 + Box-Muller), the same on the device and in the CPU checker.
`maskstarsPSF` and its helpers (:94-333, the star-catalogue path) are kept as host functions: a handful of circles per frame
rasterised from header scalars and catalogue rows — there is no per-pixel work in them worth a kernel.
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


def detect_sources_batch(images, nsigma=1.5, npixels=8, want_labels=False, stream=None, stats=None):
    """photutils.detect_threshold + detect_sources as skyBackground.py:104-110 / imageutils.py:66-71 call them.
    stats: the rows of clipped_stats_batch(images, 3.0, None) when the caller has them already.
    Returns dict(labels: CUDA int32 [B, n, n] or None, info: CUDA float64 [B, 4] — column 0 = number of segments)."""
    from .skyBackground import clipped_stats_batch
    images = _prep(images)
    cs = stats if stats is not None else clipped_stats_batch(images, 3.0, None, stream=stream)
    thr = (cs[:, 6] + float(nsigma) * cs[:, 8]).contiguous()
    return _run(images, thr, npixels, None, None, 0, want_labels, False, stream)


def maskstars_batch(images, seed=0, nsigma=1.5, npixels=8, want_labels=False, stream=None):
    """imageutils.maskstarsSEG for a batch: dict(clean: CUDA tensor like `images`, labels, info [B, 4] = segments kept,
    segments cleaned, pixels replaced, status)."""
    from .skyBackground import clipped_stats_both
    images = _prep(images)
    c5, cn = clipped_stats_both(images, 3.0, 5, stream=stream)           # imageutils.py:62 (maxiters 5) and detect_threshold (:67, converged)
    thr = (cn[:, 6] + float(nsigma) * cn[:, 8]).contiguous()
    return _run(images, thr, npixels, c5[:, 6].contiguous(), c5[:, 8].contiguous(), seed, want_labels, True, stream)


def maskstarsSEG(image, seed=0, device="cuda:0"):
    """imageutils.maskstarsSEG (imageutils.py:42-91) for one image: the cleaned image as a float64 numpy array."""
    x = torch.from_numpy(np.ascontiguousarray(image, dtype=np.float64))[None].to(device)
    return maskstars_batch(x, seed=seed)["clean"][0].cpu().numpy()


# ---- the star-catalogue path (imageutils.py:94-333): host functions ------------------------------------------------------
def _circle_mask(shape, centre, radius):
    """imageutils.py:94-132: boolean disc about `centre` = (position along axis 0, position along axis 1), boundary included.
    (Upstream multiplies by an angular mask `theta % 2 pi <= 2 pi`, which holds for every pixel.)"""
    d0 = np.arange(shape[0], dtype=np.float64)[:, None] - centre[0]
    d1 = np.arange(shape[1], dtype=np.float64)[None, :] - centre[1]
    return d0 * d0 + d1 * d1 <= radius * radius


def _calculateRadius(y, A, sigma):
    """imageutils.py:135-165: the x at which A exp(-x^2 / (2 sigma^2)) = y, for either order of A and y."""
    return sigma * np.sqrt(np.abs(2.0 * np.log(A / y)))


def _subpixel_disc(shape, pixelPos, radius, subpixels=5):
    """np.where(CircularAperture(pixelPos, r).to_mask(method="subpixel").to_image(shape) > 0, 1., 0.) (imageutils.py:243-245):
    a pixel counts when any of its subpixels x subpixels sample points lies inside the circle (photutils' circular_overlap_grid
    with use_exact = 0; pixel (row, col) is the unit square about (x, y) = (col, row))."""
    out = np.zeros(shape)
    px, py = float(pixelPos[0]), float(pixelPos[1])
    if not (radius > 0) or not np.isfinite([px, py, radius]).all():
        return out
    c0, c1 = max(0, int(np.floor(px - radius - 0.5))), min(shape[1] - 1, int(np.ceil(px + radius + 0.5)))
    r0, r1 = max(0, int(np.floor(py - radius - 0.5))), min(shape[0] - 1, int(np.ceil(py + radius + 0.5)))
    if c1 < c0 or r1 < r0:
        return out
    off = (np.arange(subpixels) + 0.5) / subpixels - 0.5
    ys = (np.arange(r0, r1 + 1)[:, None] + off[None, :]).ravel() - py
    xs = (np.arange(c0, c1 + 1)[:, None] + off[None, :]).ravel() - px
    inside = (ys[:, None] ** 2 + xs[None, :] ** 2) < radius * radius
    hit = inside.reshape(r1 - r0 + 1, subpixels, c1 - c0 + 1, subpixels).any(axis=(1, 3))
    out[r0:r1 + 1, c0:c1 + 1] = hit
    return out


def _calculateAdaptiveRadius(mask, pixelPos, image, skyCount, sky_err):
    """imageutils.py:262-333: the column through the star, read from the star's row away from the image centre for at most half the
    image (masked pixels count as skyCount + 1000); the length of the first run of consecutive samples above mean + sky_err, minus
    one, is the extra radius.  0 for a star on the first / last row (and, where upstream fails on an unbound variable because no
    sample exceeds the threshold, 0 as well)."""
    side = image.shape[0]
    x, y = int(pixelPos[0]), int(pixelPos[1])
    if pixelPos[1] >= side - 1 or pixelPos[1] <= 0:
        return 0
    direction = 1 if pixelPos[1] > int(side / 2) else -1
    rows = y + direction * np.arange(0, int(side / 2) + 1)
    rows = rows[(rows >= 0) & (rows < side)]
    profile = np.where(mask[rows, x], skyCount + 1000, image[rows, x])
    above = np.flatnonzero(profile > profile.mean() + sky_err)
    if above.size == 0:
        return 0
    gaps = np.flatnonzero(np.diff(above) != 1)
    return int(gaps[0]) if gaps.size else int(above.size - 1)


def _sky_magnitude(header, skyCount):
    """The asinh magnitude of the sky level from the SDSS photometric header cards (imageutils.py:199-214;
    https://classic.sdss.org/dr7/algorithms/fluxcal.html#counts2mag).  KeyError when a card is missing."""
    counts = skyCount - header["SOFTBIAS"]
    flux_ratio = (counts / header["EXPTIME"]) * 10 ** (0.4 * (header["PHT_AA"] + header["PHT_KK"] * header["AIRMASS"]))
    b = header["PHT_B"]
    return -(2.5 / np.log(10.)) * (np.arcsinh(flux_ratio / (2 * b)) + np.log(b)), counts


def maskstarsPSF(image, objs, header, skyCount, numSigmas=5., adaptive=True, sky_err=0.0):
    """imageutils.py:168-259: mask (True = keep) of the catalogue stars `objs` ([ra, dec, type, psfMag_r] rows): for each a disc of
    numSigmas * sigma_PSF * sqrt(|2 ln(psfMag / skyMag)|) pixels about its position, the sky magnitude from the SDSS photometric
    header cards.  header: mapping with PSF_S1, PSF_S2, EXPTIME, PHT_AA, PHT_KK, AIRMASS, SOFTBIAS, PHT_B (KeyError when one is
    missing, which main.py:389-391 catches) and the WCS cards `image.sky_to_pixel` understands (TAN).  With no objects the mask is
    all ones (of the image's type), like upstream.  adaptive: grow each disc by `_calculateAdaptiveRadius` (a subpixel-sampled disc;
    main.py:387 never asks for it)."""
    from .image import sky_to_pixel
    psf_sigma = max(header["PSF_S1"], header["PSF_S2"])
    sky_mag, counts = _sky_magnitude(header, skyCount)
    image = np.asarray(image)
    if len(objs) == 0:
        return np.ones_like(image)
    covered = np.zeros(image.shape, dtype=bool)
    for ra, dec, _kind, mag in objs:
        x, y = sky_to_pixel(header, ra, dec)
        radius = numSigmas * _calculateRadius(sky_mag, mag, psf_sigma)
        covered |= _circle_mask(image.shape, (y, x), radius)
        if adaptive:
            extra = _calculateAdaptiveRadius(covered, (x, y), image, counts, sky_err)
            covered |= _subpixel_disc(image.shape, (x, y), radius + extra) > 0
    return ~covered
