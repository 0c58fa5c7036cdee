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
