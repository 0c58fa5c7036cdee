"""Property tests of the restated third-party leaves (oracle/leaves.py): scikit-image and photutils are not
installed here, so these geometric / algebraic identities are what checks the restatements."""
import math

import numpy as np

from oracle import leaves, pm_oracle as O


def test_exact_overlap_sums_to_circle_area_and_is_monotone():
    rng = np.random.default_rng(0)
    for _ in range(60):
        x, y, r = rng.uniform(20, 40), rng.uniform(20, 40), rng.uniform(0.05, 15)
        m = leaves._circle_mask(x, y, r)
        assert abs(m.data.sum() - math.pi * r * r) <= 1e-12 * max(1.0, math.pi * r * r)
        assert m.data.min() >= -1e-15 and m.data.max() <= 1 + 1e-15
        m2 = leaves._circle_mask(x, y, r * 1.1)
        img = np.ones((64, 64))
        assert m2.weighted_sum(img) > m.weighted_sum(img)


def test_overlap_matches_monte_carlo():
    rng = np.random.default_rng(1)
    pts = rng.uniform(-0.5, 0.5, (200000, 2))
    for (dx, dy, r) in ((3, 1, 3.2), (0, 4, 4.3), (2, 2, 2.9), (0, 0, 0.4)):
        w = leaves.overlap_single_exact(dx - 0.5, dy - 0.5, dx + 0.5, dy + 0.5, r)
        mc = np.mean((pts[:, 0] + dx) ** 2 + (pts[:, 1] + dy) ** 2 <= r * r)
        assert abs(w - mc) < 5e-3


def test_annulus_is_outer_minus_inner():
    img = np.random.default_rng(2).random((50, 50))
    a = leaves.CircularAnnulus((24.0, 26.0), 3.3, 11.7).do_photometry(img)[0][0]
    b = leaves.CircularAperture((24.0, 26.0), 11.7).do_photometry(img)[0][0] - \
        leaves.CircularAperture((24.0, 26.0), 3.3).do_photometry(img)[0][0]
    assert abs(a - b) < 1e-10 * abs(b)


def test_rotate_is_an_involution_and_matches_index_maps():
    rng = np.random.default_rng(3)
    img = rng.random((41, 41))
    c = (19, 22)
    r2 = leaves.rotate(leaves.rotate(img, 180.0, center=c, preserve_range=True), 180.0, center=c, preserve_range=True)
    inner = (slice(8, 33), slice(8, 33))
    np.testing.assert_allclose(r2[inner], img[inner], rtol=0, atol=1e-12)
    r90 = leaves.rotate(img, 90.0, center=c, preserve_range=True)
    cx, cy = c
    for r in range(10, 30):
        for col in range(10, 30):
            rr, cc = cy + (col - cx), cx - (r - cy)
            assert abs(r90[r, col] - img[rr, cc]) < 1e-12


def test_resize_index_maps():
    from scipy import ndimage as ndi
    src = np.arange(49.0).reshape(7, 7)
    up = leaves.resize(src, (23, 23), order=0, mode="edge")
    idx = np.floor((np.arange(23) + 0.5) * 7 / 23).astype(int)
    assert np.array_equal(up, src[idx][:, idx])
    f = np.random.default_rng(4).random((30, 30))
    z = leaves.resize(f, (10, 10), anti_aliasing=False, preserve_range=True)
    assert np.array_equal(z, ndi.zoom(f, 1 / 3, order=1, mode="mirror", grid_mode=True, prefilter=False))


def test_gini_and_moments_known_answers():
    assert abs(leaves.gini(np.full(100, 3.0))) < 1e-15
    x = np.zeros(1000); x[-1] = 1.0
    assert abs(leaves.gini(x) - 1.0) < 1e-12
    img = np.zeros((9, 9)); img[2, 5] = 2.0
    M = leaves.moments(img, order=1)
    assert (M[1, 0] / M[0, 0], M[0, 1] / M[0, 0]) == (2.0, 5.0)
    assert abs(leaves.moments_central(img, center=(2.0, 5.0), order=2)[2, 0]) < 1e-15


def test_aperture_closed_form_matches_the_float_form():
    for n in (15, 16, 31, 64):
        for r2 in (4, 9, 10, 37, 100, 113):
            hw = O.aperture_halfwidths(n, r2)
            ap = O.aperpixmap(n, math.sqrt(r2), 9, 0.1)
            c = n // 2
            for i in range(n):
                e = hw[abs(i - c)]
                want = np.zeros(n)
                if e >= 0:
                    want[max(c - e, 0):min(c + e, n - 1) + 1] = 1
                assert np.array_equal(ap[i], want)


def test_point_symmetric_image_has_zero_asymmetry():
    n, c = 65, 32
    yy, xx = np.mgrid[0:n, 0:n].astype(float)
    img = 100 * np.exp(-(((xx - c) / 6) ** 2 + ((yy - c) / 4) ** 2))
    seg = (img > 1.0).astype(float)
    ap = O.aperpixmap(n, O.calcRmax(img, seg), 9, 0.1)
    A = O.calcA(img, seg, ap, [c, c], 180.0)
    As = O.calcA(seg, seg, ap, [c, c], 180.0)
    assert abs(A[0]) < 1e-12 and abs(As[0]) < 1e-12


# ---- independent pins of the restated leaves against scipy (installed here): a different code path computes the same
# quantity, so a shared-mode error of the restatement would show ("leaf shims are unpinned and shared")
def test_rotate_matches_scipy_map_coordinates():
    """skimage.transform.rotate == bilinear inverse warp: the same warp done by scipy.ndimage.map_coordinates(order=1),
    with the output -> input map written independently here, for the angles of the path and generic ones, integer and
    fractional centres, cval 0 and 1 (asymmetry.py:189-196)."""
    from scipy import ndimage as ndi
    rng = np.random.default_rng(11)
    img = rng.random((37, 45)) * 100.0
    for angle, center, cval in ((180.0, (20, 17), 0.0), (90.0, (22, 18), 0.0), (37.0, (21.3, 16.8), 0.0), (180.0, (19.5, 18.5), 0.0),
                                (-63.5, (10.0, 30.0), 1.0), (180.0, (3, 4), 1.0)):
        got = leaves.rotate(img, angle, center=center, preserve_range=True, cval=cval, clip=False)
        th = math.radians(angle)
        yy, xx = np.mgrid[0:img.shape[0], 0:img.shape[1]].astype(np.float64)
        dx, dy = xx - center[0], yy - center[1]
        xin = math.cos(th) * dx - math.sin(th) * dy + center[0]
        yin = math.sin(th) * dx + math.cos(th) * dy + center[1]
        want = ndi.map_coordinates(img, [yin, xin], order=1, mode="constant", cval=cval)
        # scipy zeroes every sample whose coordinate lies outside [0, n - 1]; the tap-wise rule blends cval in within one pixel of
        # the edge: compare where all four taps are inside the frame
        inside = (xin >= 0) & (xin <= img.shape[1] - 1) & (yin >= 0) & (yin <= img.shape[0] - 1)
        np.testing.assert_allclose(got[inside], want[inside], rtol=0, atol=1e-10)
        outside = (xin < -1) | (xin > img.shape[1]) | (yin < -1) | (yin > img.shape[0])
        assert np.all(got[outside] == cval)


def test_exact_overlap_matches_an_independent_integral():
    """photutils' exact circle / pixel overlap (casgm.py:51-52, :333-336) against the area integral
    int max(0, min(y1, s(x)) - max(y0, -s(x))) dx, s = sqrt(r^2 - x^2), done by scipy.integrate.quad with the kinks as
    break points: random pixels around circles of every size the path meets, to 1e-11."""
    from scipy import integrate
    rng = np.random.default_rng(12)

    def area(x0, y0, x1, y1, r):
        a, b = max(x0, -r), min(x1, r)
        if a >= b:
            return 0.0
        f = lambda x: max(0.0, min(y1, math.sqrt(max(r * r - x * x, 0.0))) - max(y0, -math.sqrt(max(r * r - x * x, 0.0))))
        pts = [v for yy in (y0, y1) if abs(yy) < r for v in (-math.sqrt(r * r - yy * yy), math.sqrt(r * r - yy * yy)) if a < v < b]
        return integrate.quad(f, a, b, points=sorted(pts) or None, epsabs=1e-13, epsrel=1e-13, limit=200)[0]

    worst = 0.0
    for _ in range(400):
        r = float(rng.choice([0.3, 0.8, 1.5, 3.0, 7.7, 20.0, 61.0])) * rng.uniform(0.8, 1.25)
        ang = rng.uniform(0, 2 * math.pi)
        d = r + rng.uniform(-0.9, 0.9)                       # pixels that straddle the circle
        cx, cy = d * math.cos(ang) + rng.uniform(-0.3, 0.3), d * math.sin(ang) + rng.uniform(-0.3, 0.3)
        got = leaves.overlap_single_exact(cx - 0.5, cy - 0.5, cx + 0.5, cy + 0.5, r)
        want = area(cx - 0.5, cy - 0.5, cx + 0.5, cy + 0.5, r)
        worst = max(worst, abs(got - want))
    assert worst < 1e-11, worst
    # and as the aperture sum of a frame: integer-centred apertures (what minapix hands to calcR20_R80)
    img = rng.random((31, 31))
    for r in (2.4, 7.9, 13.2):
        got = leaves.CircularAperture((15.0, 14.0), r).do_photometry(img)[0][0]
        want = sum(img[j, i] * area(i - 15.0 - 0.5, j - 14.0 - 0.5, i - 15.0 + 0.5, j - 14.0 + 0.5, r) for j in range(31) for i in range(31))
        assert abs(got - want) < 1e-10 * want


def test_gini_and_moments_match_direct_definitions():
    rng = np.random.default_rng(13)
    x = rng.normal(5.0, 3.0, 300)
    n = len(x)
    # photutils.morphology.gini == mean absolute difference form with the n (n - 1) normalisation, on |x|
    xs = sorted(float(v) for v in x)          # photutils 0.7: values sorted as they are, |x_i| weighted by 2 i - n - 1, |mean| below
    want = sum((2 * (i + 1) - n - 1) * abs(v) for i, v in enumerate(xs)) / (abs(sum(xs) / n) * n * (n - 1))
    assert abs(leaves.gini(x) - want) < 1e-12 * abs(want)
    xp = np.abs(x)
    want_p = np.abs(xp[:, None] - xp[None, :]).sum() / (2.0 * n * (n - 1) * xp.mean())
    assert abs(leaves.gini(xp) - want_p) < 1e-12
    img = rng.random((9, 11))
    M = leaves.moments(img, order=2)
    rr, cc = np.mgrid[0:9, 0:11]
    for p in range(3):
        for q in range(3):
            assert abs(M[p, q] - (rr ** p * cc ** q * img).sum()) < 1e-9 * max(1.0, abs(M[p, q]))
    mu = leaves.moments_central(img, center=(4.2, 5.1), order=2)
    assert abs(mu[2, 0] - (((rr - 4.2) ** 2) * img).sum()) < 1e-9 and abs(mu[0, 2] - (((cc - 5.1) ** 2) * img).sum()) < 1e-9
