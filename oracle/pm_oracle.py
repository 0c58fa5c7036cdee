"""CPU restatement ("port") of the PawlikMorph-LSST per-cutout morphology hot path.

TEST INFRASTRUCTURE — the checker the CUDA path is compared with.  Never imported
by the product package; only tests/, __graft_entry__.smoke() and bench.py's CPU
baseline legs use it.

What it is: a numpy/scipy restatement of the reference's L1 functions
(`pawlikMorphLSST/{pixmap,apertures,asymmetry,casgm}.py`), each citing the
file:line it follows, plus `analyse()`, which strings them together in the order
of `main.py:421-495`.  The third-party leaves live in `oracle/leaves.py`.

How it is pinned: the reference ships no tests and no golden outputs,
so the pin is the reference ITSELF run in the build container: `oracle/make_golden.py`
imports the unmodified reference modules (`oracle/ref_import.py`) and stores their
outputs for the 69+69 bundled SDSS cutouts and a synthetic set under
`tests/golden/`; `tests/test_oracle_golden.py` checks this port against those
vectors (and, when `/root/reference` is present, against the live reference).
The scikit-image / photutils leaves under the reference are themselves restated
(not installed here): PARITY UNPINNED for those leaves — see oracle/leaves.py.

Deliberate, documented deviation: `np.argsort` tie order in `minapix`
(`asymmetry.py:95`) is unstable and build-dependent upstream; here (and on the
GPU, and in the goldens) ties are ordered canonically: value descending, then
flat index descending  (== ``np.argsort(kind="stable")[::-1]``).
"""
import math

import numpy as np
from scipy import ndimage as ndi
from scipy.optimize import brentq

from . import leaves

SENTINEL = -99.0

# status codes of analyse() — mirrored by pm_row_t.status in include/pawlik_b200.h
ST_OK = 0
ST_FAINT_CENTRE = 1      # pixmap.py:178-179 -> main.py:420-424 default Result
ST_NO_CANDIDATE = 2      # asymmetry.py:127 np.argmin([]) would raise upstream
ST_BAD_BRACKET = 3       # casgm.py:65-76 loop/brentq would raise upstream

# flags of analyse() — mirrored by PM_DO_* in include/pawlik_b200.h
DO_A, DO_AS, DO_CAS, DO_S = 1, 2, 4, 8


class FaintCentreError(AttributeError):
    """The reference raises a bare AttributeError (pixmap.py:17-23)."""


# ---------------------------------------------------------------------------
# apertures.py
# ---------------------------------------------------------------------------
def distarr(npixx, npixy, cenpix):
    """apertures.py:131-165 — dist[i, j] = sqrt((i-c0)^2 + (j-c1)^2)."""
    i = np.arange(npixx, dtype=np.int64)[:, None] - int(cenpix[0])
    j = np.arange(npixy, dtype=np.int64)[None, :] - int(cenpix[1])
    return np.sqrt((i * i + j * j).astype(np.float64))


def subdistarr(npix, nsubpix, cenpix):
    """apertures.py:168-209 — oversampled distance grid, quirks kept:
    the coordinate vector is [k/ns - cen for k < (npix//2)*ns] ++ zeros(ns) ++ mirrored,
    and the y vector reuses the x one (`y1 = xneg`, apertures.py:200)."""
    half = (int(npix) // 2) * int(nsubpix)
    neg = np.arange(half) / nsubpix - cenpix[0]
    v = np.concatenate((neg, np.zeros(nsubpix), -neg[::-1]))
    return np.sqrt(v[:, None] ** 2 + v[None, :] ** 2)


def aperpixmap(npix, rad, nsubpix, frac):
    """apertures.py:42-128 — binary disc: pixel set iff the fraction of its
    nsubpix^2 sub-samples with subdist <= rad is >= frac."""
    npix, nsubpix = int(npix), int(nsubpix)
    cen = np.array([npix // 2 + 1, npix // 2 + 1])
    sd = subdistarr(npix, nsubpix, cen)[:npix * nsubpix, :npix * nsubpix]
    inside = (sd <= rad).reshape(npix, nsubpix, npix, nsubpix)
    cnt = inside.sum(axis=(1, 3))
    return ((cnt.astype(np.float64) / float(nsubpix ** 2)) >= frac).astype(np.float64)


def aperture_halfwidths(npix, rad2_int):
    """Integer closed form of aperpixmap(npix, sqrt(rad2_int), 9, 0.1):
    hw[d] = largest column offset e with mask[c+-d, c+-e] set, or -1.  Used to
    cross-check the GPU's disc table; bit-identical to aperpixmap for rad^2 integer."""
    c = npix // 2
    lim = 81 * int(rad2_int)

    def sub(d):
        return [0] * 9 if d == 0 else [9 * d + k for k in range(1, 10)]

    hw = np.full(c + 1, -1, dtype=np.int32)
    for d in range(c + 1):
        sd = sub(d)
        for e in range(c, -1, -1):
            se = sub(e)
            cnt = sum(1 for a in sd for b in se if a * a + b * b <= lim)
            if cnt >= 9:
                hw[d] = e
                break
    return hw


def apercentre(apermask, pix):
    """apertures.py:212-240 — cyclic shift by pix - (n//2 + 1) on each axis."""
    n = apermask.shape[0]
    d0 = int(pix[0]) - (n // 2 + 1)
    d1 = int(pix[1]) - (n // 2 + 1)
    return np.roll(np.roll(apermask, d0, axis=0), d1, axis=1)


# ---------------------------------------------------------------------------
# pixmap.py
# ---------------------------------------------------------------------------
def mean_filter_shrink(image, filterSize, starMask=None):
    """pixmap.py:156-173 — fs x fs mean (reflect) then bilinear shrink to m x m."""
    img = np.asarray(image, dtype=np.float64)
    if starMask is not None:
        img = img * starMask
    n = img.shape[0]
    f = ndi.uniform_filter(img, size=filterSize, mode="reflect")
    m = int(n / filterSize)
    return leaves.resize(f, (m, m), anti_aliasing=False, preserve_range=True)


def pixelmap(image, threshold, filterSize, starMask=None):
    """pixmap.py:130-213 — 8-connected component of {Z > thr} ∪ {centre} that
    contains the centre of the shrunk image Z, nearest-neighbour upsampled."""
    if filterSize % 2 == 0:
        raise NameError("sys")          # pixmap.py:160-162: `sys` is not imported upstream
    n = np.asarray(image).shape[0]
    z = mean_filter_shrink(image, filterSize, starMask)
    m = z.shape[0]
    c = m // 2
    if z[c, c] < threshold:
        raise FaintCentreError("ERROR! Central pixel too faint")
    above = z > threshold
    above[c, c] = True                  # pixmap.py:184 — centre always set
    lab, _ = ndi.label(above, structure=np.ones((3, 3), dtype=bool))
    small = (lab == lab[c, c]).astype(np.float64)
    return leaves.resize(small, (n, n), order=0, mode="edge")


def checkPixelmapEdges(segmap):
    """pixmap.py:91-127 (column test indexes with shape[0]: square images)."""
    s = np.asarray(segmap)
    last = s.shape[0] - 1
    return bool(s[0, :].sum() > 0 or s[last, :].sum() > 0 or s[:, 0].sum() > 0 or s[:, last].sum() > 0)


def calcRmax(image, segmap):
    """pixmap.py:26-55 — farthest `segmap == 1` pixel from the brightest masked pixel."""
    masked = np.asarray(image) * segmap
    cen = np.unravel_index(int(np.argmax(masked)), masked.shape)
    rr, cc = np.nonzero(np.asarray(segmap) == 1)
    d2 = (rr - cen[0]) ** 2 + (cc - cen[1]) ** 2
    return float(np.sqrt(np.float64(d2.max())))


def calcMaskedFraction(oldMask, starMask, cenpix):
    """pixmap.py:58-88."""
    sm = starMask * leaves.rotate(starMask, 180.0, center=(cenpix[0], cenpix[1]),
                                  preserve_range=True, cval=1.0)
    return 1.0 - np.sum(oldMask * sm) / np.sum(oldMask)


# ---------------------------------------------------------------------------
# asymmetry.py
# ---------------------------------------------------------------------------
def sorted_desc(values_flat):
    """Canonical descending order: value desc, flat index desc (see module doc)."""
    v = np.where(values_flat == 0.0, 0.0, values_flat)      # -0.0 -> +0.0
    idx = np.argsort(v, kind="stable")[::-1]
    return v[idx], idx


def centroid_candidates(image_mask):
    """asymmetry.py:89-108 — brightest pixels holding 20 % of the flux, as [col, row]."""
    flat = np.ravel(image_mask)
    t20 = 0.2 * np.sum(image_mask)
    vals, idx = sorted_desc(flat)
    ncol = image_mask.shape[1]
    cands, flux = [], 0
    for v, k in zip(vals, idx):
        if v > 0:
            flux += v
            cands.append([int(k % ncol), int(k // ncol)])
        if flux >= t20:
            break
    return cands, idx


def minapix(image, segmap, apermask, starMask=None, debug=None):
    """asymmetry.py:50-129 — candidate with the smallest 180-degree asymmetry.

    Returns [col, row].  `debug`, if a dict, receives the candidate list, their
    asymmetries and the canonical sort order (parity aids; not reference outputs)."""
    m = image * segmap * starMask if starMask is not None else image * segmap
    flat = np.ravel(m)
    cands, order = centroid_candidates(m)
    a = np.zeros(len(cands))
    for i, p in enumerate(cands):
        rot = leaves.rotate(m, 180.0, center=p, preserve_range=True)
        resid = np.ravel(np.abs(m - rot))
        region = np.nonzero(np.ravel(apercentre(apermask, np.asarray(p))) == 1)[0]
        a[i] = np.sum(resid[region]) / (2.0 * np.sum(np.abs(flat[region])))
    if debug is not None:
        debug.update(candidates=cands, a=a, order=order)
    return cands[int(np.argmin(a))]


def calcA(img, segmap, apermask, centroid, angle, starMask=None, noisecorrect=False):
    """asymmetry.py:132-257 — [A, Abgr]; quirks kept: the disc is NOT re-centred
    (:201), the background image is rotated about the swapped centre (:245),
    `segmap` is multiplied in place (:219)."""
    cx, cy = centroid[0], centroid[1]
    if starMask is None:
        sm = np.ones_like(img)
    else:
        sm = starMask.astype(np.float64)
        sm = sm * leaves.rotate(sm, angle, center=(cx, cy), preserve_range=True, cval=1.0)
    im = img * sm
    rot = leaves.rotate(im, angle, center=(cx, cy), preserve_range=True)
    flat = np.ravel(im)
    region = np.nonzero(np.ravel(apermask) == 1)[0]
    denom = 2.0 * np.sum(np.abs(flat[region]))
    A = np.sum(np.ravel(np.abs(im - rot))[region]) / denom
    Abgr = 0
    if noisecorrect:
        segmap *= sm
        grown = ndi.binary_dilation(segmap, structure=np.ones((9, 9)))
        mask_idx = np.nonzero(np.ravel(grown) == 1)[0]
        bgr_idx = np.nonzero(np.ravel(grown) != 1)[0]
        nb, nm = bgr_idx.shape[0], mask_idx.shape[0]
        if nb > nb / 10.0 and nm > 1:
            bgr = flat[bgr_idx]
            fill = bgr[np.arange(nm) % nb]          # :227-239 truncate or tile
            b = np.zeros(flat.shape[0])
            b[bgr_idx] = bgr
            b[mask_idx] = fill
            b = b.reshape(im.shape[0], im.shape[0])
            brot = leaves.rotate(b, 180.0, center=(cy, cx), preserve_range=True)
            Abgr = np.sum(np.ravel(np.abs(b - brot))[region]) / denom
            A = A - Abgr
        else:
            Abgr = -99
    return [A, Abgr]


def calculateAsymmetries(image, segmap):
    """asymmetry.py:13-47."""
    rmax = calcRmax(image, segmap)
    ap = aperpixmap(image.shape[0], rmax, 9, 0.1)
    ones = np.ones_like(image)
    apix = minapix(image, segmap, ap, ones)
    A = calcA(image, segmap, ap, apix, 180.0, ones, noisecorrect=True)
    As = calcA(segmap, segmap, ap, apix, 180.0, ones)
    As90 = calcA(segmap, segmap, ap, apix, 90.0, ones)
    return A[0], As[0], As90[0]


# ---------------------------------------------------------------------------
# casgm.py
# ---------------------------------------------------------------------------
def circle_flux(image, centroid, r):
    """casgm.py:51-52 — exact-overlap circular aperture sum at (x, y) = centroid."""
    return leaves.CircularAperture(centroid, r).do_photometry(image, method="exact")[0][0]


def _radius_enclosing(image, centroid, radius, fraction):
    """casgm.py:22-78 — walk down from `radius` in steps of radius/100 until the
    enclosed fraction is <= `fraction`, then brentq inside that step."""
    total = circle_flux(image, centroid, radius)
    delta = (radius / 200.0) * 2.0
    r = radius - delta
    while True:
        if circle_flux(image, centroid, r) / total <= fraction:
            lo, hi = r, r + delta
            break
        r -= delta
    return brentq(lambda x: circle_flux(image, centroid, x) / total - fraction, lo, hi)


def calcR20_R80(image, segmap, centroid, radius):
    """casgm.py:124-155."""
    masked = image * segmap
    return (_radius_enclosing(masked, centroid, radius, 0.2),
            _radius_enclosing(masked, centroid, radius, 0.8))


def concentration(r20, r80):
    """casgm.py:158-184."""
    return 5.0 * np.log10(r80 / r20)


def gini(image, segmap):
    """casgm.py:187-218."""
    return leaves.gini(image[segmap > 0])


def m20(image, segmap):
    """casgm.py:221-273."""
    img = np.where(segmap > 0, image, 0.0)
    M = leaves.moments(img, order=1)
    cen = (M[1, 0] / M[0, 0], M[0, 1] / M[0, 0])
    mc = leaves.moments_central(img, center=cen, order=2)
    tot = mc[2, 0] + mc[0, 2]
    s = np.sort(img.ravel())
    frac = np.cumsum(s) / np.sum(s)
    thresh = s[frac > 0.8][0]
    mc20 = leaves.moments_central(np.where(img >= thresh, img, 0.0), center=cen, order=2)
    return np.log10((mc20[0, 2] + mc20[2, 0]) / tot)


def background_box(image, segmap, sky):
    """casgm.py:367-398 — first all-sky box found by the 2-pixel raster walk.
    Returns (x, y, size).  Quirk kept: `not np.any(box) == -99.` is always True."""
    work = image * ~np.array(segmap, dtype=bool)
    work[work == 0.0] = -99
    size, x, y = 32, 0, 0
    while True:
        box = work[y:y + size, x:x + size]
        if abs(sky) > abs(np.sum(box) / size ** 2):
            break
        x += 2
        if x >= image.shape[1] - size:
            x = 0
            y += 2
        if y >= image.shape[0] - size:
            if size > 8:
                x = y = 0
                size //= 2
            else:
                break
    return x, y, size


def smoothness(image, segmap, centroid, Rmax, r20, sky):
    """casgm.py:276-406."""
    ann = leaves.CircularAnnulus(centroid, r20, Rmax)
    k = int(r20)
    diff = image - ndi.uniform_filter(image, size=k)
    diff[diff < 0.0] = 0.0
    flux = ann.do_photometry(image, method="exact")[0][0]
    dflux = ann.do_photometry(diff, method="exact")[0][0]
    x, y, size = background_box(image, segmap, sky)
    bkg = image[y:y + size, x:x + size]
    bd = bkg - ndi.uniform_filter(bkg, size=k)
    bd[bd < 0.0] = 0.0
    bs = np.sum(bd) / float(bkg.size)
    return (dflux - ann.area * bs) / flux


# ---------------------------------------------------------------------------
# main.py:421-495 — the per-cutout dataflow, as one call
# ---------------------------------------------------------------------------
ROW_FIELDS = ("apix_col", "apix_row", "rmax", "r20", "r80", "C", "A", "Abgr", "S", "As",
              "As90", "gini", "m20", "object_edge", "status", "n_candidates")


def analyse(image_with_sky, sky, sky_err, filterSize=3, flags=DO_A | DO_AS | DO_CAS | DO_S,
            segmap_in=None):
    """One cutout through the hot path in the order of main.py:421-495.

    Returns (row dict over ROW_FIELDS, segmap uint8 [n,n], debug dict).  Image
    cleaning (`maskstarsSEG`, main.py:445) is the identity here — it is out of
    scope and non-deterministic upstream.  S is computed by a direct
    `smoothness` call (commented out upstream, main.py:494)."""
    img = np.asarray(image_with_sky, dtype=np.float64)
    n = img.shape[0]
    row = {k: SENTINEL for k in ROW_FIELDS}
    row.update(object_edge=0.0, status=float(ST_OK), n_candidates=0.0)
    dbg = {}
    if segmap_in is None:
        try:
            seg = pixelmap(img, sky + sky_err, filterSize)
        except FaintCentreError:
            row["status"] = float(ST_FAINT_CENTRE)
            return row, np.zeros((n, n), np.uint8), dbg
        row["object_edge"] = float(checkPixelmapEdges(seg))
    else:
        seg = np.asarray(segmap_in, dtype=np.float64)
    seg_u8 = seg.astype(np.uint8)
    im = img - sky
    ones = np.ones_like(im)
    rmax = calcRmax(im, seg)
    row["rmax"] = rmax
    ap = aperpixmap(n, rmax, 9, 0.1)
    cands, _ = centroid_candidates(im * seg)
    row["n_candidates"] = float(len(cands))
    if not cands:
        row["status"] = float(ST_NO_CANDIDATE)
        return row, seg_u8, dbg
    apix = minapix(im, seg, ap, ones, debug=dbg)
    row["apix_col"], row["apix_row"] = float(apix[0]), float(apix[1])
    if flags & (DO_A | DO_CAS):
        A = calcA(im, seg, ap, apix, 180.0, ones, noisecorrect=True)
        row["A"], row["Abgr"] = float(A[0]), float(A[1])
    if flags & DO_AS:
        row["As"] = float(calcA(seg, seg, ap, apix, 180.0, ones)[0])
        row["As90"] = float(calcA(seg, seg, ap, apix, 90.0, ones)[0])
    if flags & DO_CAS:
        try:
            r20, r80 = calcR20_R80(im, seg, apix, rmax)
            if not (math.isfinite(r20) and math.isfinite(r80)):
                raise ValueError
            row["r20"], row["r80"] = float(r20), float(r80)
            row["C"] = float(concentration(r20, r80))
        except (ValueError, RuntimeError):
            row["status"] = float(ST_BAD_BRACKET)
        row["gini"] = float(gini(im, seg))
        row["m20"] = float(m20(im, seg))
        if flags & DO_S and row["status"] == ST_OK and int(row["r20"]) >= 1:
            # int(r20) == 0 makes scipy's uniform_filter raise upstream (uncaught): S stays -99
            row["S"] = float(smoothness(im, seg, apix, rmax, row["r20"], sky))
    return row, seg_u8, dbg
