"""TEST INFRASTRUCTURE — CPU restatement of sersic.fitSersic (/root/reference/pawlikMorphLSST/sersic.py:53-136, called at
main.py:480-488).  Only tests/, __graft_entry__.smoke() and the CPU legs of bench.py may import this.

The reference leans on three third-party pieces that are NOT in this container (astropy 4.x, photutils 0.7 are absent;
scipy is here):
  * photutils EllipticalAperture / EllipticalAnnulus `do_photometry(method="exact")`: the sum of image x (exact area of
    ellipse ∩ pixel).  Restated as the area of (parallelogram ∩ unit circle) after mapping the ellipse to the unit circle,
    summed edge by edge (inside pieces: triangles about the origin; outside pieces: circular sectors) — any exact method
    gives the same number to rounding.  PARITY UNPINNED against photutils itself; pinned here against brute-force
    supersampling (tests/test_sersic.py).
  * scipy.optimize.brentq (sersic.py:114): the real one is called.
  * astropy.modeling.fitting.LevMarLSQFitter on models.Sersic2D (sersic.py:92, :125-134): astropy's fitter is a thin wrapper
    around scipy.optimize.leastsq (MINPACK lmdif: Sersic2D has no analytic derivative) with
    `maxfev=maxiter, epsfcn=sqrt(finfo(float).eps), xtol=acc` and the default ftol; Sersic2D.evaluate is five lines around
    scipy.special.gammaincinv.  Both restated with the real scipy routines, so the fit is pinned to MINPACK as the reference
    drives it, not to astropy's wrapper code (parity unpinned for the wrapper)."""
import numpy as np
from scipy.optimize import brentq, leastsq
from scipy.special import gammaincinv


# ---- exact ellipse / pixel overlap ------------------------------------------------------------------------------------
def _edge_area(px, py, qx, qy):
    """Signed contribution of the edge p -> q to the area of polygon ∩ unit disc (vectorised over edges)."""
    dx, dy = qx - px, qy - py
    a = dx * dx + dy * dy
    b = 2.0 * (px * dx + py * dy)
    c = px * px + py * py - 1.0
    disc = b * b - 4.0 * a * c
    out = np.zeros_like(px)

    def sector(ux, uy, vx, vy):            # half the signed angle from u to v (unit-radius sector area)
        return 0.5 * np.arctan2(ux * vy - uy * vx, ux * vx + uy * vy)

    def tri(ux, uy, vx, vy):
        return 0.5 * (ux * vy - uy * vx)

    safe_a = np.where(a > 0.0, a, 1.0)
    sq = np.sqrt(np.maximum(disc, 0.0))
    t1 = np.clip((-b - sq) / (2.0 * safe_a), 0.0, 1.0)
    t2 = np.clip((-b + sq) / (2.0 * safe_a), 0.0, 1.0)
    hit = (disc > 0.0) & (a > 0.0) & (t2 > t1)
    # no crossing inside the segment: wholly inside (both ends inside) or wholly outside
    p_in = c <= 0.0
    q_in = (qx * qx + qy * qy - 1.0) <= 0.0
    whole_in = ~hit & p_in & q_in
    out = np.where(whole_in, tri(px, py, qx, qy), out)
    out = np.where(~hit & ~whole_in, sector(px, py, qx, qy), out)
    # crossing: p .. t1 outside, t1 .. t2 inside, t2 .. q outside (pieces of zero length contribute nothing)
    ax, ay = px + t1 * dx, py + t1 * dy
    bx, by = px + t2 * dx, py + t2 * dy
    cross = sector(px, py, ax, ay) * (t1 > 0.0) + tri(ax, ay, bx, by) + sector(bx, by, qx, qy) * (t2 < 1.0)
    return np.where(hit, cross, out)


def ellipse_overlap(shape, xc, yc, a, b, theta):
    """Fraction of every pixel of an image of `shape` covered by the ellipse of semi-axes a (along theta, anticlockwise from
    +x) and b about (xc, yc) — pixel (row r, column c) is the unit square centred on (x, y) = (c, r), photutils' convention."""
    ny, nx = shape
    w = np.zeros((ny, nx))
    ex = np.sqrt((a * np.cos(theta)) ** 2 + (b * np.sin(theta)) ** 2)
    ey = np.sqrt((a * np.sin(theta)) ** 2 + (b * np.cos(theta)) ** 2)
    c0, c1 = max(0, int(np.floor(xc - ex - 0.5))), min(nx - 1, int(np.ceil(xc + ex + 0.5)))
    r0, r1 = max(0, int(np.floor(yc - ey - 0.5))), min(ny - 1, int(np.ceil(yc + ey + 0.5)))
    if c1 < c0 or r1 < r0:
        return w
    rr, cc = np.mgrid[r0:r1 + 1, c0:c1 + 1]
    ct, st = np.cos(theta), np.sin(theta)
    xs = [cc - 0.5 - xc, cc + 0.5 - xc, cc + 0.5 - xc, cc - 0.5 - xc]          # counter-clockwise corners
    ys = [rr - 0.5 - yc, rr - 0.5 - yc, rr + 0.5 - yc, rr + 0.5 - yc]
    us = [(x * ct + y * st) / a for x, y in zip(xs, ys)]
    vs = [(-x * st + y * ct) / b for x, y in zip(xs, ys)]
    area = np.zeros(rr.shape)
    for k in range(4):
        area += _edge_area(us[k], vs[k], us[(k + 1) % 4], vs[(k + 1) % 4])
    w[r0:r1 + 1, c0:c1 + 1] = np.clip(area * a * b, 0.0, 1.0)
    return w


def ellipse_sum(image, xc, yc, a, b, theta):
    """EllipticalAperture((xc, yc), a, b, theta).do_photometry(image, method="exact")[0][0]"""
    return float((np.asarray(image, dtype=np.float64) * ellipse_overlap(image.shape, xc, yc, a, b, theta)).sum())


def ellipse_overlap_supersampled(shape, xc, yc, a, b, theta, sub=40):
    """Brute force check of ellipse_overlap: fraction of sub x sub sample points of each pixel inside the ellipse."""
    ny, nx = shape
    off = (np.arange(sub) + 0.5) / sub - 0.5
    y = (np.arange(ny)[:, None] + off[None, :]).ravel()
    x = (np.arange(nx)[:, None] + off[None, :]).ravel()
    X, Y = np.meshgrid(x - xc, y - yc)
    ct, st = np.cos(theta), np.sin(theta)
    inside = ((X * ct + Y * st) / a) ** 2 + ((-X * st + Y * ct) / b) ** 2 <= 1.0
    return inside.reshape(ny, sub, nx, sub).mean(axis=(1, 3))


# ---- Sersic2D (astropy.modeling.functional_models.Sersic2D.evaluate) ----------------------------------------------------
def sersic2d(x, y, amplitude, r_eff, n, x_0, y_0, ellip, theta):
    bn = gammaincinv(2.0 * n, 0.5)
    a, b = r_eff, (1.0 - ellip) * r_eff
    ct, st = np.cos(theta), np.sin(theta)
    x_maj = (x - x_0) * ct + (y - y_0) * st
    x_min = -(x - x_0) * st + (y - y_0) * ct
    z = np.sqrt((x_maj / a) ** 2 + (x_min / b) ** 2)
    return amplitude * np.exp(-bn * (z ** (1.0 / n) - 1.0))


SERSIC_FIELDS = ("amplitude", "r_eff", "n", "x_0", "y_0", "ellip", "theta")


def sersic_start(image, centroid, fwhms, theta):
    """sersic.py:98-123: the start values (amplitude, r_eff, n, x_0, y_0, ellip, theta) of the fit."""
    image = np.asarray(image, dtype=np.float64)
    xc, yc = float(centroid[0]), float(centroid[1])
    b = 2.0 * min(fwhms)
    a = 2.0 * max(fwhms)
    ellip = 1.0 - (b / a)
    total = ellipse_sum(image, xc, yc, a, b, theta)
    delta = (a / 100.0) * 2.0
    cur = a - delta
    while True:
        cur = max(0.001, cur)
        if ellipse_sum(image, xc, yc, cur, b, theta) / total <= 0.5:
            a_min, a_max = cur, cur + delta
            break
        cur -= delta
    r_eff = brentq(lambda s: ellipse_sum(image, xc, yc, s, b, theta) / total - 0.5, a_min, a_max)
    a_in, a_out = r_eff - 0.5, r_eff + 0.5
    b_out = a_out - ellip
    b_in = b_out * a_in / a_out                                      # photutils' default inner minor axis
    flux = ellipse_sum(image, xc, yc, a_out, b_out, theta) - ellipse_sum(image, xc, yc, a_in, b_in, theta)
    area = np.pi * (a_out * b_out - a_in * b_in)
    return np.array([flux / area, r_eff, 2.0, xc, yc, ellip, theta])


def fit_sersic(image, centroid, fwhms, theta, star_mask=None, return_info=False):
    """sersic.fitSersic: the seven fitted Sersic2D parameters (SERSIC_FIELDS order)."""
    image = np.asarray(image, dtype=np.float64)
    if star_mask is not None:
        image = image * star_mask
    p0 = sersic_start(image, centroid, fwhms, theta)
    ny, nx = image.shape
    y, x = np.mgrid[0:ny, 0:nx]

    def objective(p):
        return np.ravel(sersic2d(x, y, *p) - image)

    with np.errstate(all="ignore"):
        p, _, info, _, ier = leastsq(objective, p0, Dfun=None, col_deriv=True, maxfev=1000, epsfcn=np.sqrt(np.finfo(float).eps),
                                     xtol=1e-8, full_output=True)
    if return_info:
        return p, dict(start=p0, nfev=info["nfev"], ier=ier, cost=float((objective(p) ** 2).sum()))
    return p
