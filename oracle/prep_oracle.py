"""CPU restatement of the pre-processing around the morphology path — TEST INFRASTRUCTURE ONLY (tests/, smoke, bench
cpu_baseline); never imported by the product.

  * gauss2dfit / fit_gauss   skyBackground._gauss2dfit / _fitGauss (skyBackground.py:167-214, :306-331): gaussfitter.moments
    (gaussfitter.py:40-84) and gaussfitter.twodgaussian (:87-148) restated, fitted with the REAL scipy.optimize.curve_fit.
    PINNED: oracle/make_prep_golden.py runs the reference's own unmodified gaussfitter.py (it imports numpy only) through
    curve_fit and tests/test_prep_oracle.py checks this restatement against those vectors bit for bit.
  * detect_threshold / detect_sources   photutils 0.7 (skyBackground.py:104-110, imageutils.py:66-71): sigma_clipped_stats
    with maxiters=None, then scipy.ndimage.convolve(mode="constant") > threshold, scipy.ndimage.label with the 8-connected
    structure, segments of fewer than npixels pixels removed, relabelled consecutively.  The scipy calls are the real
    library; the photutils / astropy glue around them is restated from their published behaviour: PARITY UNPINNED for
    that glue (neither library is installed, the reference has no fixtures).
  * maskstarsSEG   imageutils.py:42-91 with the reference's unseeded np.random.normal replaced by a Philox4x32-10 /
    Box-Muller normal deviate of (seed, cutout index, pixel index) — the same counter-based stream the CUDA kernel uses.
  * calc_skybgr    skyBackground._calcSkybgr (:81-165) on top of the pieces above and oracle/sky_oracle.py.
"""
import math

import numpy as np
from scipy import ndimage, optimize

from oracle import sky_oracle as S


# ---- gaussfitter.py ------------------------------------------------------------------------------------------------
def moments(data, angle_guess=90.0):
    """gaussfitter.py:40-84: [height, amplitude, x, y, width_x, width_y, angle]."""
    y = int(data.shape[1] / 2)
    x = int(data.shape[0] / 2)
    col = data[int(y), :]
    width_x = np.sqrt(np.abs((np.arange(col.size) - y) ** 2 * col).sum() / np.abs(col).sum())
    row = data[:, int(x)]
    width_y = np.sqrt(np.abs((np.arange(row.size) - x) ** 2 * row).sum() / np.abs(row).sum())
    height = np.median(data.ravel())
    amplitude = data.max() - height
    if np.isnan((width_y, width_x, height, amplitude)).any():
        raise ValueError("something is nan")
    return [height, amplitude, x, y, width_x, width_y, angle_guess * np.pi / 180.]


def twodgaussian(xydata, offset, amplitude, xo, yo, sigma_x, sigma_y, theta):
    """gaussfitter.py:87-148 (centre: column yo, row xo — `center_y, center_x = xo, yo`, :132)."""
    center_y, center_x = xo, yo
    rcen_x = center_x * np.cos(theta) - center_y * np.sin(theta)
    rcen_y = center_x * np.sin(theta) + center_y * np.cos(theta)
    x, y = xydata[0], xydata[1]
    xp = x * np.cos(theta) - y * np.sin(theta)
    yp = x * np.sin(theta) + y * np.cos(theta)
    g = offset + amplitude * np.exp(-(((rcen_x - xp) / sigma_x) ** 2 + ((rcen_y - yp) / sigma_y) ** 2) / 2.)
    return g.ravel()


def gauss2dfit(img, raw=False):
    """skyBackground.py:306-331.  raw=True returns popt before the np.abs."""
    n = img.shape[0]
    X, Y = np.meshgrid(np.linspace(0, n - 1, n), np.linspace(0, n - 1, n))
    popt, _ = optimize.curve_fit(twodgaussian, (X, Y), img.ravel(), p0=moments(img))
    return popt if raw else np.abs(popt)


def theta_quadrant(theta):
    """skyBackground.py:201-211."""
    if theta > (np.pi / 2.):
        return np.pi - theta
    thetaDeg = theta / (2. * np.pi)
    r = thetaDeg - int(thetaDeg)
    return (r * 2. * np.pi) - (np.pi / 2.)


def fit_gauss(img):
    """skyBackground._fitGauss (:167-214): r_in, [fwhm_x, fwhm_y], theta."""
    yfit = gauss2dfit(img)
    factor = 2 * np.sqrt(2 * np.log(2))
    fwhm_x, fwhm_y = factor * np.abs(yfit[4]), factor * np.abs(yfit[5])
    return 2. * max(fwhm_x, fwhm_y), [fwhm_x, fwhm_y], theta_quadrant(yfit[6])


# ---- photutils glue -----------------------------------------------------------------------------------------------------
def clipped(img, sigma=3.0, maxiters=5):
    """astropy sigma_clipped_stats over every pixel: mean, median, std (maxiters=None: until a round removes nothing)."""
    m, med, s, _, _ = S.sigma_clipped_stats(np.asarray(img, dtype=np.float64).ravel(), sigma, 10 ** 9 if maxiters is None else maxiters)
    return m, med, s


def detect_threshold(img, nsigma):
    m, _, s = clipped(img, 3.0, None)
    return m + nsigma * s


def gaussian_kernel3():
    sig = 3.0 * (1.0 / (2.0 * math.sqrt(2.0 * math.log(2.0))))
    yy, xx = np.mgrid[-1:2, -1:2]
    k = np.exp(-(xx * xx + yy * yy) / (2.0 * sig * sig))
    return k / k.sum()


def detect_sources(img, threshold, npixels=8):
    """Label image (0 = background; 1.. in scipy.ndimage.label order, small segments removed) and the number of segments."""
    sm = ndimage.convolve(np.asarray(img, dtype=np.float64), gaussian_kernel3(), mode="constant", cval=0.0)
    lab, nl = ndimage.label(sm > threshold, structure=np.ones((3, 3)))
    if nl == 0:
        return lab, 0
    counts = np.bincount(lab.ravel(), minlength=nl + 1)
    keep = np.zeros(nl + 1, dtype=np.int64)
    good = [l for l in range(1, nl + 1) if counts[l] >= npixels]
    keep[good] = np.arange(1, len(good) + 1)
    return keep[lab], len(good)


# ---- Philox4x32-10 + Box-Muller: the stream of pm_prep.cuh ---------------------------------------------------------------------
def philox4x32(c, k):
    c = [np.uint32(v) for v in c]
    k0, k1 = np.uint32(k[0]), np.uint32(k[1])
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * np.uint64(c[0])
        p1 = np.uint64(0xCD9E8D57) * np.uint64(c[2])
        c = [np.uint32(p1 >> np.uint64(32)) ^ c[1] ^ k0, np.uint32(p1 & np.uint64(0xffffffff)),
             np.uint32(p0 >> np.uint64(32)) ^ c[3] ^ k1, np.uint32(p0 & np.uint64(0xffffffff))]
        k0 = np.uint32((int(k0) + 0x9E3779B9) & 0xffffffff)
        k1 = np.uint32((int(k1) + 0xBB67AE85) & 0xffffffff)
    return c


def philox_normal(seed, cutout, pixels):
    """standard normal deviates of (seed, cutout, pixel) for an int array of pixel indices (vectorised over pixels)."""
    pixels = np.asarray(pixels, dtype=np.uint64)
    c = [pixels.astype(np.uint32), np.full(pixels.shape, cutout & 0xffffffff, np.uint32), np.full(pixels.shape, (cutout >> 32) & 0xffffffff, np.uint32),
         np.full(pixels.shape, 0x504d4253, np.uint32)]
    k0 = np.full(pixels.shape, seed & 0xffffffff, np.uint32)
    k1 = np.full(pixels.shape, (seed >> 32) & 0xffffffff, np.uint32)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = np.uint64(0xD2511F53) * c[0].astype(np.uint64)
            p1 = np.uint64(0xCD9E8D57) * c[2].astype(np.uint64)
            c = [(p1 >> np.uint64(32)).astype(np.uint32) ^ c[1] ^ k0, p1.astype(np.uint32),
                 (p0 >> np.uint64(32)).astype(np.uint32) ^ c[3] ^ k1, p0.astype(np.uint32)]
            k0 = k0 + np.uint32(0x9E3779B9)
            k1 = k1 + np.uint32(0xBB67AE85)
    u1 = (c[0].astype(np.float64) + 1.0) * (1.0 / 4294967296.0)
    u2 = (c[1].astype(np.float64) + 1.0) * (1.0 / 4294967296.0)
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(6.283185307179586 * u2)


def maskstarsSEG(image, seed=0, cutout=0, return_info=False):
    """imageutils.py:42-91.  A kept segment whose bounding box does not contain cenpix is a star; inside the bounding box of
    every star, every pixel carrying ANY label is replaced by noise (np.where(data_ma > 0, 1, 0) reads the label cutout whether
    masked or not) and the zeros left in the box are put back from the image (:86-89)."""
    image = np.asarray(image, dtype=np.float64)
    n = image.shape[0]
    cen = int(image.shape[0] / 2) + 1
    mean, _, std = clipped(image, 3.0, 5)
    lab, nseg = detect_sources(image, detect_threshold(image, 1.5), 8)
    clean = image.copy()
    info = dict(n_segments=nseg, n_cleaned=0, n_replaced=0)
    if nseg == 0:
        return (clean, info) if return_info else clean
    hit = np.zeros(image.shape, bool)
    for sl in ndimage.find_objects(lab):
        r0, r1, c0, c1 = sl[0].start, sl[0].stop - 1, sl[1].start, sl[1].stop - 1
        inbox = cen > c0 - 0.5 and cen > r0 - 0.5 and cen < c1 + 0.5 and cen < r1 + 0.5        # _inBbox(extent, cenpix), strict
        if not inbox:
            info["n_cleaned"] += 1
            hit[sl] |= lab[sl] > 0
    rr, cc = np.nonzero(hit)
    v = mean + std * philox_normal(seed, cutout, rr * n + cc)
    v = np.where(v == 0.0, image[rr, cc], v)
    clean[rr, cc] = v
    info["n_replaced"] = int(hit.sum())
    return (clean, info) if return_info else clean


class SkyError(Exception):
    pass


def calc_skybgr(img, large=None, seed=0, cutout=0):
    """skyBackground.skybgr / _calcSkybgr (:32-165) without the astropy fall-back fitter: raises NotImplementedError where
    upstream would switch to _fitAltGauss.  Returns sky, sky_err, fwhms, theta."""
    img = np.asarray(img, dtype=np.float64)
    work = img if large is None else maskstarsSEG(np.asarray(large, dtype=np.float64), seed, cutout)
    n = work.shape[0]
    _, nseg = detect_sources(work, detect_threshold(work, 1.5), 8)
    if nseg == 0:
        raise SkyError("No object in image")
    try:
        r_in, fwhms, theta = fit_gauss(img)
    except RuntimeError:
        raise NotImplementedError("_fitAltGauss")
    fwhms = [min(f, 1.414 * n) for f in fwhms]
    row = S.calc_sky(work, fwhms, theta)
    if row["count"] < 300 or r_in > n * 3:
        raise NotImplementedError("_fitAltGauss")
    if row["count"] < 100:
        raise SkyError(f"Error! Sky region too small {int(row['count'])}")
    return row["sky"], row["sky_err"], fwhms, theta
