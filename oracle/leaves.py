"""CPU restatements of the third-party leaf routines the reference hot path calls.

TEST INFRASTRUCTURE ONLY — nothing under ``oracle/`` is imported by the product
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline legs use it, and only as the checker.

The reference (`/root/reference/pawlikMorphLSST/*.py`) gets its arithmetic from
scikit-image, photutils and scipy.  scipy and numpy are installed here and are
used directly; scikit-image (pinned 0.16.2, `docs/requirements.txt:108`) and
photutils (pinned 0.7.1/0.7.2, `environment.yml:13`, `docs/requirements.txt:76`)
are NOT installed and cannot be fetched, so their published algorithms are
restated below.  PARITY UNPINNED for these leaves: they are checked by
geometric/algebraic property tests (`tests/test_oracle_leaves.py`), not against
the real libraries.

Each function cites the reference call site it serves.
"""
import math

import numpy as np
from scipy import ndimage as _ndi

# --------------------------------------------------------------------------
# skimage.transform.resize            (pixmap.py:171-173 and pixmap.py:211-212)
# --------------------------------------------------------------------------
_SK2NDI = {"constant": "constant", "edge": "nearest", "symmetric": "reflect",
           "reflect": "mirror", "wrap": "wrap"}


def resize(image, output_shape, order=None, mode="reflect", cval=0, clip=True,
           preserve_range=False, anti_aliasing=None, anti_aliasing_sigma=None):
    """skimage.transform.resize for 2-D float input, no anti-aliasing.

    scikit-image >= 0.19 implements this as
    ``scipy.ndimage.zoom(image, out/in, order, mode, grid_mode=True)`` (pixel
    centres aligned: u = (i + 0.5) * n_in / n_out - 0.5); 0.16.2 reached the same
    mapping through an affine ``warp``.  order defaults to 1 (bilinear) for
    non-bool input.  The output is clipped to the input range (a no-op for
    orders 0 and 1).
    """
    image = np.asarray(image)
    if anti_aliasing:
        raise NotImplementedError("oracle resize: anti_aliasing not needed on the path")
    if order is None:
        order = 0 if image.dtype == bool else 1
    img = image.astype(np.float64)
    zoom = [o / i for o, i in zip(output_shape, img.shape)]
    out = _ndi.zoom(img, zoom, order=order, mode=_SK2NDI[mode], cval=cval,
                    grid_mode=True, prefilter=order > 1)
    if clip and order != 0 and img.size:
        np.clip(out, img.min(), img.max(), out=out)
    return out


# --------------------------------------------------------------------------
# skimage.transform.rotate   (asymmetry.py:114, :189-191, :195-196, :245;
#                             pixmap.py:81-83)
# --------------------------------------------------------------------------
def rotate(image, angle, resize=False, center=None, order=None, mode="constant",
           cval=0, clip=True, preserve_range=False):
    """skimage.transform.rotate restated: inverse-map bilinear warp.

    tform = T(center) @ R(angle) @ T(-center) on (x=col, y=row) coordinates is
    used as the map from OUTPUT to INPUT coordinates; bilinear taps with
    floor/ceil, out-of-range taps take ``cval`` (mode="constant"); result
    clipped to [min, max] of the input (extended by cval).
    """
    if resize or mode != "constant":
        raise NotImplementedError
    img = np.asarray(image, dtype=np.float64)
    rows, cols = img.shape
    if order is None:
        order = 1
    if center is None:
        center = np.array((cols, rows)) / 2.0 - 0.5
    center = np.asarray(center, dtype=np.float64)
    th = np.deg2rad(angle)
    t1 = np.array([[1, 0, center[0]], [0, 1, center[1]], [0, 0, 1.0]])
    t2 = np.array([[math.cos(th), -math.sin(th), 0], [math.sin(th), math.cos(th), 0], [0, 0, 1.0]])
    t3 = np.array([[1, 0, -center[0]], [0, 1, -center[1]], [0, 0, 1.0]])
    # skimage: tform3 + tform2 + tform1  ==  t1 @ t2 @ t3
    M = t1 @ (t2 @ t3)
    rr, cc = np.mgrid[0:rows, 0:cols].astype(np.float64)
    x = M[0, 0] * cc + M[0, 1] * rr + M[0, 2]
    y = M[1, 0] * cc + M[1, 1] * rr + M[1, 2]

    def pix(r, c):
        ok = (r >= 0) & (r < rows) & (c >= 0) & (c < cols)
        v = np.full(r.shape, float(cval))
        v[ok] = img[r[ok], c[ok]]
        return v

    if order == 0:
        out = pix(np.rint(y).astype(np.int64), np.rint(x).astype(np.int64))
    elif order == 1:
        minr = np.floor(y).astype(np.int64)
        minc = np.floor(x).astype(np.int64)
        maxr = np.ceil(y).astype(np.int64)
        maxc = np.ceil(x).astype(np.int64)
        dr = y - minr
        dc = x - minc
        top = (1 - dc) * pix(minr, minc) + dc * pix(minr, maxc)
        bottom = (1 - dc) * pix(maxr, minc) + dc * pix(maxr, maxc)
        out = (1 - dr) * top + dr * bottom
    else:
        raise NotImplementedError
    if clip and order != 0 and img.size:
        lo, hi = img.min(), img.max()
        if not (lo <= cval <= hi):
            lo, hi = min(lo, cval), max(hi, cval)
        np.clip(out, lo, hi, out=out)
    return out


# --------------------------------------------------------------------------
# skimage.measure.moments / moments_central        (casgm.py:253, :257, :268)
# --------------------------------------------------------------------------
def moments_central(image, center=None, order=3, **kwargs):
    """M[p, q] = sum_{r,c} (r - center[0])**p (c - center[1])**q image[r, c]."""
    image = np.asarray(image)
    if center is None:
        center = (0,) * image.ndim
    calc = image.astype(np.float64)
    for dim, dim_length in enumerate(image.shape):
        delta = np.arange(dim_length, dtype=np.float64) - center[dim]
        powers_of_delta = delta[:, np.newaxis] ** np.arange(order + 1)
        calc = np.rollaxis(calc, dim, image.ndim)
        calc = np.dot(calc, powers_of_delta)
        calc = np.rollaxis(calc, -1, dim)
    return calc


def moments(image, order=3):
    return moments_central(image, (0,) * np.ndim(image), order=order)


# --------------------------------------------------------------------------
# photutils exact circle / pixel overlap
#   (casgm.py:51-52, :66-67, :116-117, :323, :333-336)
# --------------------------------------------------------------------------
def _fsqrt(v):
    return math.sqrt(v) if v > 0.0 else 0.0


def _area_arc(x1, y1, x2, y2, r):
    a = math.sqrt((x2 - x1) ** 2 + (y2 - y1) ** 2)
    theta = 2.0 * math.asin(0.5 * a / r)
    return 0.5 * r * r * (theta - math.sin(theta))


def _area_triangle(x1, y1, x2, y2, x3, y3):
    return 0.5 * abs(x1 * (y2 - y3) + x2 * (y3 - y1) + x3 * (y1 - y2))


def _overlap_core(xmin, ymin, xmax, ymax, r):
    """Area of circle(0, r) ∩ [xmin,xmax]x[ymin,ymax] for a rectangle in the first quadrant."""
    if xmin * xmin + ymin * ymin > r * r:
        return 0.0
    if xmax * xmax + ymax * ymax < r * r:
        return (xmax - xmin) * (ymax - ymin)
    d1 = _fsqrt(xmax * xmax + ymin * ymin)
    d2 = _fsqrt(xmin * xmin + ymax * ymax)
    if d1 < r and d2 < r:
        x1, y1 = _fsqrt(r * r - ymax * ymax), ymax
        x2, y2 = xmax, _fsqrt(r * r - xmax * xmax)
        return ((xmax - xmin) * (ymax - ymin)
                - _area_triangle(x1, y1, x2, y2, xmax, ymax)
                + _area_arc(x1, y1, x2, y2, r))
    if d1 < r:
        x1, y1 = xmin, _fsqrt(r * r - xmin * xmin)
        x2, y2 = xmax, _fsqrt(r * r - xmax * xmax)
        return (_area_arc(x1, y1, x2, y2, r)
                + _area_triangle(x1, y1, x1, ymin, xmax, ymin)
                + _area_triangle(x1, y1, x2, ymin, x2, y2))
    if d2 < r:
        x1, y1 = _fsqrt(r * r - ymin * ymin), ymin
        x2, y2 = _fsqrt(r * r - ymax * ymax), ymax
        return (_area_arc(x1, y1, x2, y2, r)
                + _area_triangle(x1, y1, xmin, y1, xmin, ymax)
                + _area_triangle(x1, y1, xmin, y2, x2, y2))
    x1, y1 = _fsqrt(r * r - ymin * ymin), ymin
    x2, y2 = xmin, _fsqrt(r * r - xmin * xmin)
    return _area_arc(x1, y1, x2, y2, r) + _area_triangle(x1, y1, x2, y2, xmin, ymin)


def overlap_single_exact(xmin, ymin, xmax, ymax, r):
    """Area of circle(0, r) ∩ rectangle, any quadrant.

    photutils recurses: a rectangle straddling an axis is split at that axis and
    each piece is mirrored into the first quadrant.  Written here without
    recursion (x split outer, y split inner — the same pieces; the order the
    <=4 pieces are added differs from the recursion only at the 1e-16 level).
    """
    total = 0.0
    for ix in range(2):
        if ix == 0:
            if xmin >= 0.0:
                continue
            xa, xb = xmin, min(xmax, 0.0)
        else:
            if xmax <= 0.0:
                continue
            xa, xb = max(xmin, 0.0), xmax
        for iy in range(2):
            if iy == 0:
                if ymin >= 0.0:
                    continue
                ya, yb = ymin, min(ymax, 0.0)
            else:
                if ymax <= 0.0:
                    continue
                ya, yb = max(ymin, 0.0), ymax
            # mirror the piece into the first quadrant exactly as photutils does
            if xa >= 0.0:
                if ya >= 0.0:
                    total += _overlap_core(xa, ya, xb, yb, r)
                else:
                    total += _overlap_core(-yb, xa, -ya, xb, r)
            else:
                if ya >= 0.0:
                    total += _overlap_core(-xb, ya, -xa, yb, r)
                else:
                    total += _overlap_core(-xb, -yb, -xa, -ya, r)
    return total


def circular_overlap_grid(xmin, xmax, ymin, ymax, nx, ny, r):
    """photutils.geometry.circular_overlap_grid(use_exact=1): [ny, nx] weights."""
    frac = np.zeros((ny, nx))
    dx = (xmax - xmin) / nx
    dy = (ymax - ymin) / ny
    pixel_radius = 0.5 * math.sqrt(dx * dx + dy * dy)
    bxmin, bxmax = -r - 0.5 * dx, r + 0.5 * dx
    bymin, bymax = -r - 0.5 * dy, r + 0.5 * dy
    for i in range(nx):
        pxmin = xmin + i * dx
        pxcen = pxmin + dx * 0.5
        pxmax = pxmin + dx
        if pxmax > bxmin and pxmin < bxmax:
            for j in range(ny):
                pymin = ymin + j * dy
                pycen = pymin + dy * 0.5
                pymax = pymin + dy
                if pymax > bymin and pymin < bymax:
                    d = math.sqrt(pxcen * pxcen + pycen * pycen)
                    if d < r - pixel_radius:
                        frac[j, i] = 1.0
                    elif d < r + pixel_radius:
                        frac[j, i] = overlap_single_exact(pxmin, pymin, pxmax, pymax, r) / (dx * dy)
    return frac


try:  # speed only: the same Python source, LLVM-compiled
    import numba as _nb
    _fsqrt = _nb.njit(cache=True)(_fsqrt)
    _area_arc = _nb.njit(cache=True)(_area_arc)
    _area_triangle = _nb.njit(cache=True)(_area_triangle)
    _overlap_core = _nb.njit(cache=True)(_overlap_core)
    overlap_single_exact = _nb.njit(cache=True)(overlap_single_exact)
    circular_overlap_grid = _nb.njit(cache=True)(circular_overlap_grid)
except ImportError:  # pragma: no cover
    pass


class _Mask:
    def __init__(self, data, ixmin, iymin):
        self.data = data
        self.ixmin, self.iymin = ixmin, iymin
        self.iymax, self.ixmax = iymin + data.shape[0], ixmin + data.shape[1]

    def weighted_sum(self, image):
        ny, nx = image.shape
        y0, y1 = max(self.iymin, 0), min(self.iymax, ny)
        x0, x1 = max(self.ixmin, 0), min(self.ixmax, nx)
        if y1 <= y0 or x1 <= x0:
            return float("nan")
        cut = image[y0:y1, x0:x1]
        w = self.data[y0 - self.iymin:y1 - self.iymin, x0 - self.ixmin:x1 - self.ixmin]
        return np.sum(cut * w)


def _circle_mask(x, y, r):
    if not r > 0:
        raise ValueError("'r' must be a positive scalar")
    ixmin = int(math.floor(x - r + 0.5))
    ixmax = int(math.ceil(x + r + 0.5))
    iymin = int(math.floor(y - r + 0.5))
    iymax = int(math.ceil(y + r + 0.5))
    nx, ny = ixmax - ixmin, iymax - iymin
    w = circular_overlap_grid(ixmin - 0.5 - x, ixmax - 0.5 - x,
                              iymin - 0.5 - y, iymax - 0.5 - y, nx, ny, r)
    return _Mask(w, ixmin, iymin)


class CircularAperture:
    """photutils.CircularAperture, single position, method="exact" only."""

    def __init__(self, positions, r):
        self.positions = np.asarray(positions, dtype=np.float64)
        self.r = float(r)

    @property
    def area(self):
        return math.pi * self.r ** 2

    def to_mask(self, method="exact"):
        return _circle_mask(self.positions[0], self.positions[1], self.r)

    def do_photometry(self, data, error=None, mask=None, method="exact", subpixels=5, unit=None):
        if method != "exact":
            raise NotImplementedError
        data = np.asarray(data, dtype=np.float64)
        return (np.array([self.to_mask().weighted_sum(data)]), np.array([]))


class CircularAnnulus(CircularAperture):
    """photutils.CircularAnnulus: outer-circle weights minus inner-circle weights."""

    def __init__(self, positions, r_in, r_out):
        if not (r_out > r_in):
            raise ValueError("r_out must be greater than r_in")
        self.positions = np.asarray(positions, dtype=np.float64)
        self.r_in, self.r_out = float(r_in), float(r_out)

    @property
    def area(self):
        return math.pi * (self.r_out ** 2 - self.r_in ** 2)

    def to_mask(self, method="exact"):
        x, y = self.positions
        outer = _circle_mask(x, y, self.r_out)
        w = outer.data.copy()
        # photutils evaluates the inner circle on the OUTER bounding box grid
        inner = circular_overlap_grid(outer.ixmin - 0.5 - x, outer.ixmax - 0.5 - x,
                                      outer.iymin - 0.5 - y, outer.iymax - 0.5 - y,
                                      w.shape[1], w.shape[0], self.r_in)
        return _Mask(w - inner, outer.ixmin, outer.iymin)


# --------------------------------------------------------------------------
# photutils.morphology.gini                                   (casgm.py:216)
# --------------------------------------------------------------------------
def gini(data):
    flattened = np.sort(np.ravel(data))
    npix = np.size(flattened)
    normalization = np.abs(np.mean(flattened)) * npix * (npix - 1)
    kernel = (2.0 * np.arange(1, npix + 1) - npix - 1) * np.abs(flattened)
    return np.sum(kernel) / normalization
