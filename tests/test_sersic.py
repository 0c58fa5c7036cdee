"""fitSersic (sersic.py:53-136) on the CPU build of the kernel sources: exact elliptical aperture sums, the start values and
the Sersic2D fit against the oracle (scipy's brentq and leastsq as the reference drives them; the exact ellipse / pixel
overlap restated and pinned against supersampling here)."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import sersic_oracle as S
from tests import emul_harness as H
from tests import sersic_cases as C

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _start_values():
    """`start_values` of the product's sersic.py, loaded without the package (importing it needs the CUDA library)."""
    src = open(os.path.join(REPO, "pawlikmorph-lsst_b200", "sersic.py")).read()
    a, b = src.index("def start_values("), src.index("def sersic_start_batch(")
    ns = {"np": np}
    exec(src[a:b], ns)
    return ns["start_values"]


def test_oracle_overlap_is_exact():
    rng = np.random.default_rng(0)
    for _ in range(8):
        xc, yc = rng.uniform(8, 24, 2)
        a = rng.uniform(0.3, 7)
        b = rng.uniform(0.2, a)
        th = rng.uniform(-3, 3)
        w = S.ellipse_overlap((32, 32), xc, yc, a, b, th)
        assert abs(w.sum() - np.pi * a * b) <= 1e-12 * np.pi * a * b                   # the ellipse lies inside the image
        assert np.abs(w - S.ellipse_overlap_supersampled((32, 32), xc, yc, a, b, th, sub=40)).max() < 4e-3
        assert w.min() >= 0 and w.max() <= 1
    tiny = S.ellipse_overlap((8, 8), 3.2, 4.1, 0.001, 0.0005, 0.3)                   # inside one pixel
    area = np.pi * 0.001 * 0.0005
    assert np.count_nonzero(tiny > 1e-9 * area) == 1 and abs(tiny.sum() - area) < 1e-9 * area      # (sectors of far pixels cancel to rounding)
    cut = S.ellipse_overlap((16, 16), 0.0, 7.0, 5.0, 2.0, 0.0)                       # half outside the image
    assert abs(cut.sum() - (np.pi * 10.0 / 2 + S.ellipse_overlap((16, 16), 8.0, 7.0, 5.0, 2.0, 0.0)[:, 8].sum() / 2)) < 1e-9


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_kernel_ellipse_sums(dtype):
    rng = np.random.default_rng(1)
    n = 64
    img = rng.normal(10, 1, (3, n, n)).astype(dtype)
    ell = [[rng.uniform(20, 44), rng.uniform(20, 44), rng.uniform(2, 15), rng.uniform(1, 8), rng.uniform(-3, 3)] for _ in range(5)]
    ell += [[31.3, 30.2, 0.001, 0.5, 0.4], [2.0, 60.0, 9.0, 4.0, 1.0], [-30.0, -30.0, 3.0, 2.0, 0.0], [32.0, 32.0, 80.0, 70.0, 0.2]]
    idx = [0, 1, 2, 1, 0, 1, 2, 0, 1]
    got = H.ellipse_sums(img, ell, idx=idx)
    want = np.array([S.ellipse_sum(img[i], *e) for i, e in zip(idx, ell)])
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)
    assert got[7] == 0.0 and abs(got[8] - img[1].astype(np.float64).sum()) <= 1e-9 * abs(got[8])   # off the image; covering it
    assert np.isnan(H.ellipse_sums(img[:1], [[10.0, 10.0, -1.0, 2.0, 0.0]])[0])


def test_start_values_follow_the_reference_walk():
    imgs, truths, cen, fw, th = C.cases()
    start_values = _start_values()

    def sums(xc, yc, a, b, theta, which):
        return H.ellipse_sums(imgs, np.stack([xc, yc, a, b, theta], axis=1), idx=which)

    p0, st = start_values(cen, fw, th, sums)
    assert np.all(st == 0)
    for k in range(len(cen)):
        np.testing.assert_allclose(p0[k], S.sersic_start(imgs[k], cen[k], fw[k], th[k]), rtol=1e-10)
    # what upstream cannot handle: negative widths (the -99 defaults of a run without skybgr), an empty image, NaN
    bad_fw = [list(fw[0]), [-99.0, -99.0], list(fw[2]), [np.nan, 1.0]]
    zero = imgs.copy()
    zero[2] = 0.0

    def sums0(xc, yc, a, b, theta, which):
        return H.ellipse_sums(zero, np.stack([xc, yc, a, b, theta], axis=1), idx=which)

    p0b, stb = start_values(cen, bad_fw, th, sums0)
    assert list(stb) == [0, 4, 4, 4] and np.isnan(p0b[1:]).all()
    np.testing.assert_allclose(p0b[0], p0[0], rtol=1e-12)


def test_kernel_fit_reaches_the_optimum_of_leastsq():
    imgs, truths, cen, fw, th = C.cases()
    p0 = np.stack([S.sersic_start(imgs[k], cen[k], fw[k], th[k]) for k in range(len(cen))])
    out = H.sersicfit(imgs, p0)
    assert np.all(out[:, 9] == 0)
    for k in range(len(cen)):
        want, info = S.fit_sersic(imgs[k], cen[k], fw[k], th[k], return_info=True)
        assert info["ier"] in (1, 2, 3, 4)
        assert C.close(out[k, :7], want), (k, out[k, :7], want)
        assert out[k, 7] <= info["cost"] * (1 + 1e-9)                                # at least as deep as leastsq's stop
        assert C.close(out[k, :7], truths[k], rtol=2e-2)                             # and it is the profile that was put in
    f32 = H.sersicfit(imgs.astype(np.float32), p0)
    for k in range(len(cen)):
        assert C.close(f32[k, :7], out[k, :7], rtol=1e-4)


def test_kernel_fit_failures():
    imgs, truths, cen, fw, th = C.cases(count=2)
    p0 = np.stack([S.sersic_start(imgs[k], cen[k], fw[k], th[k]) for k in range(2)])
    p0[1, 2] = np.nan
    out = H.sersicfit(imgs, p0)
    assert out[0, 9] == 0 and out[1, 9] == 2
    p0[1] = p0[0]
    p0[1, 2] = -1.0                                                                   # b_n of a negative index is NaN
    assert H.sersicfit(imgs, p0)[1, 9] == 2
