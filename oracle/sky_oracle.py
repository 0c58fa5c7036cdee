"""CPU restatement of the sky-region statistics of `skyBackground._calcSkybgr` (skyBackground.py:134-164) — TEST
INFRASTRUCTURE ONLY (tests/, smoke, bench cpu_baseline); never imported by the product.

PARITY UNPINNED: the arithmetic lives in photutils (EllipticalAperture.to_mask(method="center"), photutils 0.7.x) and
astropy (sigma_clipped_stats, astropy.stats.SigmaClip), neither of which is installed here, and the reference has no
tests or fixtures for this function.  Restated from their published behaviour:
  * `_getSkyRegion` (skyBackground.py:265-303): a = 2 min(fwhms), b = 2 max(fwhms); a pixel belongs to the OBJECT when its
    centre (x = column, y = row), relative to cenpix = (int(n/2)+1, int(n/2)+1), satisfies
    (x cos t + y sin t)^2 / a^2 + (y cos t - x sin t)^2 / b^2 < 1  (photutils geometry/elliptical_overlap.pyx, subpixels = 1).
  * np.ma.mean / np.ma.median / np.ma.std of the remaining pixels (:153-155).
  * mean <= median: sky = mean, sky_err = std (:157-160); else sigma_clipped_stats(img, mask=object, sigma=3.) (:162):
    up to 5 rounds of [median - 3 std, median + 3 std] (bounds inclusive) on the surviving values, stopping when a
    round removes nothing; then mean, median, std (ddof 0) of the survivors; sky = 3 median - 2 mean (:163).
"""
import math

import numpy as np


def object_region(n, fwhms, theta):
    a, b = 2.0 * min(fwhms), 2.0 * max(fwhms)
    cen = int(n / 2) + 1
    y, x = np.mgrid[0:n, 0:n].astype(np.float64)
    x -= cen
    y -= cen
    c, s = math.cos(theta), math.sin(theta)
    xt = y * s + x * c
    yt = y * c - x * s
    return xt * xt * (1.0 / (a * a)) + yt * yt * (1.0 / (b * b)) < 1.0


def sigma_clipped_stats(values, sigma=3.0, maxiters=5):
    v = np.asarray(values, dtype=np.float64).ravel()
    it, changed = 0, 1
    while changed != 0 and it < maxiters:
        it += 1
        size = v.size
        centre, std = np.median(v), np.std(v)
        v = v[(v >= centre - std * sigma) & (v <= centre + std * sigma)]
        changed = size - v.size
    return float(np.mean(v)), float(np.median(v)), float(np.std(v)), it, v.size


def calc_sky(img, fwhms, theta):
    """dict over the fields of pm_sky_stats_batch for one image."""
    img = np.asarray(img, dtype=np.float64)
    sky_px = img[~object_region(img.shape[0], fwhms, theta)]
    out = dict(count=float(sky_px.size), status=1.0 if sky_px.size < 100 else 0.0, clip_iterations=0.0)
    out["mean"], out["median"], out["std"] = float(np.mean(sky_px)), float(np.median(sky_px)), float(np.std(sky_px))
    out.update(clipped_count=out["count"], clipped_mean=out["mean"], clipped_median=out["median"], clipped_std=out["std"])
    if out["mean"] <= out["median"]:
        out["sky"], out["sky_err"] = out["mean"], out["std"]
    else:
        m, med, s, it, cnt = sigma_clipped_stats(sky_px)
        out.update(clip_iterations=float(it), clipped_count=float(cnt), clipped_mean=m, clipped_median=med, clipped_std=s)
        out["sky"], out["sky_err"] = 3.0 * med - 2.0 * m, s
    return out
