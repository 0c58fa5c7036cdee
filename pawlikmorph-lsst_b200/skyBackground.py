"""Sky background of a cutout — `pawlikMorphLSST/skyBackground.py` on the device, for batches.

`skybgr` upstream = (1) a source-detection sanity check (skyBackground.py:104-113), (2) a 2-D Gaussian fit of the object
(scipy curve_fit of gaussfitter.twodgaussian, :167-214, :306-331; astropy's LevMar fitter as the fall-back, :217-262,
:334-395) giving `fwhms, theta`, (3) the statistics of the pixels outside the ellipse of semi-axes 2 min(fwhms),
2 max(fwhms) (:134-164, :265-303); with `largeImage` (the `-li` mode, :71-76) steps (1) and (3) run on the cleaned larger
cutout and (2) on the small one.  All three run on the GPU here (`pm_maskstars_batch`, `pm_clipped_stats_batch`,
`pm_gaussfit_batch`, `pm_sky_stats_batch`); `skybgr_batch` chains them for a batch, `skybgr` is the per-image function with
the reference's signature and error behaviour.

Deviations: the Gaussian fit is a Levenberg-Marquardt iteration with the analytic Jacobian run to the optimum:
it finds the same Gaussian as curve_fit (cost, centre, |sigma| to ~1e-6) but, among the equivalent optima (sign of sigma,
theta + k pi/2 with the sigmas swapped) — between which MINPACK's path chooses unpredictably — it reports the one its own
path reaches, so `theta` can differ from upstream's by such a step.  The fall-back fitter of :334-395 is astropy's
Gaussian2D + LevMarLSQFitter; here the same iteration is restarted on the pre-processed image of :359-374 with the
fall-back's start values (parity unpinned).
"""
import math

import numpy as np
import torch

from . import _abi
from ._lib import lib


class _SkyError(AttributeError):
    """skyBackground.py:20-27: upstream's _SkyError prints its message and raises AttributeError, which main.py:350-368 catches."""

    def __init__(self, value=""):
        super().__init__(value)
        if value:
            print(value)


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


# ---- the whole of skybgr for a batch --------------------------------------------------------------------------------
FWHM_FACTOR = 2.0 * math.sqrt(2.0 * math.log(2.0))


def _as_cuda(images):
    if not torch.is_tensor(images) or not images.is_cuda:
        raise TypeError("needs a CUDA tensor (there is no CPU path)")
    if images.dtype not in (torch.float32, torch.float64) or images.ndim != 3 or images.shape[1] != images.shape[2]:
        raise ValueError("images must be float32 / float64 [B, n, n]")
    return images.contiguous()


def clipped_stats_batch(images, sigma=3.0, maxiters=5, stream=None):
    """astropy.stats.sigma_clipped_stats over every pixel of each cutout (maxiters=None: until convergence).
    CUDA float64 [B, 10], columns `_abi.CLIP_FIELDS`."""
    images = _as_cuda(images)
    B, n = int(images.shape[0]), int(images.shape[1])
    dev = images.device
    L = lib()
    with torch.cuda.device(dev):
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        with torch.cuda.stream(st):
            out = torch.empty((B, _abi.PM_CLIP_COLS), dtype=torch.float64, device=dev)
            rc = L.pm_clipped_stats_batch(images.data_ptr(), _abi.PM_F32 if images.dtype == torch.float32 else _abi.PM_F64, B, n,
                                          float(sigma), 0 if maxiters is None else int(maxiters), out.data_ptr(), st.cuda_stream)
            _abi.check(L, rc, "pm_clipped_stats_batch")
            images.record_stream(st)
    return out


def clipped_stats_both(images, sigma=3.0, maxiters=5, stream=None):
    """(clipped_stats_batch(images, sigma, maxiters), clipped_stats_batch(images, sigma, None)) from one launch: the rounds up to
    `maxiters` are the first rounds of the converged sequence."""
    images = _as_cuda(images)
    B, n = int(images.shape[0]), int(images.shape[1])
    dev = images.device
    L = lib()
    with torch.cuda.device(dev):
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        with torch.cuda.stream(st):
            first = torch.empty((B, _abi.PM_CLIP_COLS), dtype=torch.float64, device=dev)
            conv = torch.empty((B, _abi.PM_CLIP_COLS), dtype=torch.float64, device=dev)
            rc = L.pm_clipped_stats2_batch(images.data_ptr(), _abi.PM_F32 if images.dtype == torch.float32 else _abi.PM_F64, B, n,
                                           float(sigma), int(maxiters), first.data_ptr(), conv.data_ptr(), st.cuda_stream)
            _abi.check(L, rc, "pm_clipped_stats2_batch")
            images.record_stream(st)
    return first, conv


def gaussfit_batch(images, stats=None, p0=None, stream=None):
    """skyBackground._gauss2dfit for a batch: CUDA float64 [B, 10], columns `_abi.GAUSS_FIELDS` (raw optimum: apply abs() for
    the np.abs of skyBackground.py:331).  stats: the rows of clipped_stats_batch (median and max feed gaussfitter.moments)."""
    images = _as_cuda(images)
    B, n = int(images.shape[0]), int(images.shape[1])
    dev = images.device
    L = lib()
    with torch.cuda.device(dev):
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        with torch.cuda.stream(st):
            med = mx = p0_d = None
            if p0 is not None:
                p0_d = torch.as_tensor(np.asarray(p0, dtype=np.float64)).reshape(B, 7).to(dev).contiguous()
            else:
                if stats is None:
                    stats = clipped_stats_batch(images, 3.0, 1, stream=st)
                med, mx = stats[:, 2].contiguous(), stats[:, 9].contiguous()
            out = torch.empty((B, _abi.PM_GAUSS_COLS), dtype=torch.float64, device=dev)
            rc = L.pm_gaussfit_batch(images.data_ptr(), _abi.PM_F32 if images.dtype == torch.float32 else _abi.PM_F64, B, n,
                                     None if med is None else med.data_ptr(), None if mx is None else mx.data_ptr(),
                                     None if p0_d is None else p0_d.data_ptr(), out.data_ptr(), st.cuda_stream)
            _abi.check(L, rc, "pm_gaussfit_batch")
            for t in (images, med, mx, p0_d):
                if t is not None:
                    t.record_stream(st)
    return out


def _theta_quadrant(theta):
    """skyBackground.py:201-211, vectorised."""
    theta = np.asarray(theta, dtype=np.float64)
    frac = theta / (2.0 * np.pi)
    r = frac - np.trunc(frac)
    return np.where(theta > np.pi / 2.0, np.pi - theta, r * 2.0 * np.pi - np.pi / 2.0)


def _alt_start(images):
    """Start values and pre-processed images of skyBackground._altgauss2dfit (:359-390): bright pixels farther than 20 pixels
    from the centre are set to the mean until the maximum lies in the central box, the mean is subtracted; widths from the
    central row / column, height from the median, angle 3.14 / 2."""
    x = images.to(torch.float64).clone()
    B, n = x.shape[0], x.shape[1]
    mid = int(n / 2)
    mean = x.mean(dim=(1, 2), keepdim=True)
    lo, hi = max(0, mid - 20), min(n, mid + 21)
    inner_max = x[:, lo:hi, lo:hi].amax(dim=(1, 2), keepdim=True)
    outside = torch.ones((n, n), dtype=torch.bool, device=x.device)
    outside[lo:hi, lo:hi] = False
    x = torch.where(outside[None] & (x > inner_max), mean, x)
    x = x - x.mean(dim=(1, 2), keepdim=True)
    idx = torch.arange(n, device=x.device, dtype=torch.float64)
    col, row = x[:, mid, :], x[:, :, mid]
    sx = torch.sqrt(((idx - mid) ** 2 * col).abs().sum(dim=1) / col.abs().sum(dim=1))
    sy = torch.sqrt(((idx - mid) ** 2 * row).abs().sum(dim=1) / row.abs().sum(dim=1))
    height = x.reshape(B, -1).median(dim=1).values
    amp = x.amax(dim=(1, 2)) - height
    p0 = torch.stack([torch.zeros_like(amp), amp, torch.full_like(amp, mid), torch.full_like(amp, mid), sx, sy, torch.full_like(amp, 3.14 / 2.0)], dim=1)
    return x, p0


def skybgr_batch(images, large_images=None, seed=0, stream=None):
    """skyBackground.skybgr for a batch of cutouts on the GPU.  images: CUDA [B, n, n]; large_images: CUDA [B, N, N] (the `-li`
    cutouts) or None.  Returns a dict of numpy arrays: sky, sky_err [B], fwhms [B, 2], theta [B], status [B] (0 ok, 1 "No object
    in image", 2 "Sky region too small": the two _SkyError cases, skyBackground.py:112-113, :147-148; 3 NaN start values) and
    the CUDA tensor `sky_rows` [B, 12] of the statistics."""
    from .imageutils import detect_sources_batch, maskstars_batch
    images = _as_cuda(images)
    B, n = int(images.shape[0]), int(images.shape[1])
    work = images
    if large_images is not None:
        work = maskstars_batch(_as_cuda(large_images), seed=seed, stream=stream)["clean"]             # skyBackground.py:74
    N = int(work.shape[1])
    # one pass of clipped statistics serves both: the converged ones are the detection threshold, the unclipped median / maximum of the
    # same call are gaussfitter.moments' start values (when the source check runs on the large cutout the fit takes its own)
    cstats = clipped_stats_batch(work, 3.0, None, stream=stream)
    nseg = detect_sources_batch(work, stream=stream, stats=cstats)["info"][:, 0].cpu().numpy()         # :104-113
    fit = gaussfit_batch(images, stats=cstats if large_images is None else None, stream=stream).cpu().numpy()   # :117 (on the small cutout)
    status = np.where(nseg > 0, 0, 1).astype(np.int64)
    status[(fit[:, 9] == 2) & (status == 0)] = 3

    def shape_of(f):
        fw = np.stack([FWHM_FACTOR * np.abs(f[:, 4]), FWHM_FACTOR * np.abs(f[:, 5])], axis=1)
        return fw, _theta_quadrant(np.abs(f[:, 6])), 2.0 * fw.max(axis=1)

    fwhms, theta, r_in = shape_of(fit)
    failed = (fit[:, 9] != 0)                                                                        # curve_fit's RuntimeError -> _fitAltGauss
    bad = ~np.isfinite(fwhms).all(axis=1) | (fwhms <= 0).any(axis=1)
    fw_use = np.where(bad[:, None], 1.0, np.minimum(fwhms, 1.414 * N))                               # :123-126
    rows = sky_stats_batch(work, fw_use, np.where(bad, 0.0, theta), stream=stream)
    cnt = rows[:, 0].cpu().numpy()
    redo = (failed | bad | (cnt < 300) | (r_in > N * 3)) & (status == 0)                              # :131-143
    if redo.any():
        idx = np.nonzero(redo)[0]
        sel = torch.as_tensor(idx, device=images.device)
        pre, p0 = _alt_start(images[sel])
        alt = gaussfit_batch(pre, p0=p0.cpu().numpy(), stream=stream).cpu().numpy()
        fwhms[idx] = np.stack([FWHM_FACTOR * np.abs(alt[:, 4]), FWHM_FACTOR * np.abs(alt[:, 5])], axis=1)
        theta[idx] = alt[:, 6]                                                                       # :259 (the fitter's angle as it is)
        ok = np.isfinite(fwhms[idx]).all(axis=1) & (fwhms[idx] > 0).all(axis=1)
        fw2 = np.where(ok[:, None], np.minimum(fwhms[idx], 1.414 * N), 1.0)
        rows[sel] = sky_stats_batch(work[sel], fw2, np.where(ok, theta[idx], 0.0), stream=stream)
        cnt = rows[:, 0].cpu().numpy()
    status[(cnt < 100) & (status == 0)] = 2                                                          # :147-148
    r = rows.cpu().numpy()
    return dict(sky=r[:, 9].copy(), sky_err=r[:, 10].copy(), fwhms=np.minimum(fwhms, 1.414 * N), theta=theta, status=status, sky_rows=rows)


def skybgr(img, largeImage=None, file=None, imageSource=None, device="cuda:0", seed=0):
    """skyBackground.skybgr (skyBackground.py:32-78) for one image: (sky, sky_err, fwhms, theta).  Raises AttributeError through
    _SkyError like upstream (no object in the image, sky region too small)."""
    x = torch.from_numpy(np.ascontiguousarray(img, dtype=np.float64))[None].to(device)
    big = None if largeImage is None else torch.from_numpy(np.ascontiguousarray(largeImage, dtype=np.float64))[None].to(device)
    r = skybgr_batch(x, big, seed=seed)
    if r["status"][0] == 1:
        raise _SkyError("No object in image")
    if r["status"][0] == 2:
        raise _SkyError(f"Error! Sky region too small {int(r['sky_rows'][0, 0].item())}")
    if r["status"][0] == 3:
        raise ValueError("something is nan")                                                         # gaussfitter.py:76-77
    return float(r["sky"][0]), float(r["sky_err"][0]), [float(r["fwhms"][0, 0]), float(r["fwhms"][0, 1])], float(r["theta"][0])
