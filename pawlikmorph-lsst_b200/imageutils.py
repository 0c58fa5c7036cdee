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
    """imageutils.py:94-132: boolean disc (x - cx)^2 + (y - cy)^2 <= radius^2 with x the FIRST array axis (the angular factor of
    the original, theta % 2 pi <= 2 pi, is true everywhere except where theta is NaN)."""
    x, y = np.ogrid[:shape[0], :shape[1]]
    cx, cy = centre
    r2 = (x - cx) * (x - cx) + (y - cy) * (y - cy)
    theta = np.arctan2(x - cx, y - cy) % (2 * np.pi)
    return (r2 <= radius * radius) * (theta <= 2. * np.pi)


def _calculateRadius(y, A, sigma):
    """imageutils.py:135-165: sigma * sqrt(|2 ln(A / y)|)"""
    return sigma * np.sqrt(abs(2. * np.log(A / y)))


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
    """imageutils.py:262-333: walk away from the image centre along the star's column; the run of consecutive profile points
    above mean + sky_err gives the extra radius."""
    maskedImage = image * ~mask + ((skyCount + 1000) * mask)
    yinit = maskedImage[int(pixelPos[1]), int(pixelPos[0])]
    imgsize = image.shape[0]
    delta = 1 if pixelPos[1] > int(imgsize / 2) else -1
    if pixelPos[1] >= imgsize - 1 or pixelPos[1] <= 0:
        return 0
    ypos = int(pixelPos[1]) + delta
    ys = [yinit]
    for _ in range(int(imgsize / 2)):
        ys.append(maskedImage[ypos, int(pixelPos[0])])
        ypos += delta
        if ypos >= imgsize or ypos < 0:
            break
    thres = np.mean(ys) + sky_err
    arr = np.where(ys > thres)[0]
    i = 0                                   # (upstream leaves `i` unbound when nothing exceeds the threshold: UnboundLocalError)
    for i, val in enumerate(arr):
        if i + 1 > arr.shape[0] - 1:
            break
        if val + 1 != arr[i + 1]:
            break
    return i


def maskstarsPSF(image, objs, header, skyCount, numSigmas=5., adaptive=True, sky_err=0.0):
    """imageutils.py:168-259: mask (True = keep) of the catalogue stars `objs` ([ra, dec, type, psfMag_r] rows): for each a disc of
    numSigmas * sigma_PSF * sqrt(|2 ln(psfMag / skyMag)|) pixels about its position, the sky magnitude from the SDSS photometric
    header cards (asinh magnitudes).  header: mapping with PSF_S1, PSF_S2, EXPTIME, PHT_AA, PHT_KK, AIRMASS, SOFTBIAS, PHT_B
    (KeyError when one is missing, which main.py:389-391 catches) and the WCS cards `image.sky_to_pixel` understands (TAN).
    With no objects the mask is all ones, like upstream."""
    from .image import sky_to_pixel
    sigma1, sigma2 = header["PSF_S1"], header["PSF_S2"]
    expTime, aa, kk = header["EXPTIME"], header["PHT_AA"], header["PHT_KK"]
    airMass, softwareBias, b = header["AIRMASS"], header["SOFTBIAS"], header["PHT_B"]
    skyCount = skyCount - softwareBias
    factor = 0.4 * (aa + kk * airMass)
    skyFluxRatio = (skyCount / expTime) * 10 ** factor
    skyMag = -(2.5 / np.log(10.)) * (np.arcsinh(skyFluxRatio / (2 * b)) + np.log(b))
    image = np.asarray(image)
    if len(objs) == 0:
        return np.ones_like(image)
    starmask = np.zeros_like(image)
    for obj in objs:
        pixelPos = sky_to_pixel(header, obj[0], obj[1])
        sigma = max(sigma1, sigma2)
        radius = numSigmas * _calculateRadius(skyMag, obj[3], sigma)
        starmask = np.logical_or(starmask, _circle_mask(starmask.shape, (pixelPos[1], pixelPos[0]), radius))
        if adaptive:
            extra = _calculateAdaptiveRadius(starmask, pixelPos, image, skyCount, sky_err)
            starmask = np.logical_or(starmask, _subpixel_disc(image.shape, pixelPos, radius + extra))
    return ~starmask
