"""Pre-processing around the path: the oracle restatement against the reference's own gaussfitter.py
(golden vectors) and the emulator build of the kernels against the oracle — clipped statistics, Gaussian fit, source
detection, seeded maskstarsSEG."""
import os

import numpy as np
import pytest

import synth
from oracle import prep_oracle as PO
from tests import emul_harness as H
from tests.golden_util import GOLDEN_DIR

Z = np.load(os.path.join(GOLDEN_DIR, "gaussfit.npz"))
NG = int(Z["count"])


def _with_stars(B, n, seed):
    """synthetic cutouts with a few compact off-centre sources added (the things maskstarsSEG paints over)."""
    x = synth.make_exact(B, n, seed).astype(np.float64)
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:n, 0:n]
    for b in range(B):
        for _ in range(int(rng.integers(1, 5))):
            cy, cx = rng.uniform(4, n - 4, 2)
            if abs(cy - n / 2) < n / 6 and abs(cx - n / 2) < n / 6:
                continue
            x[b] += rng.uniform(60, 400) * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * rng.uniform(0.8, 2.0) ** 2))
    return x.astype(np.float32)


def test_oracle_gauss_fit_is_the_reference_gaussfitter():
    """oracle restatement of gaussfitter.moments / twodgaussian + curve_fit == vectors made by the unmodified reference file."""
    for k in range(0, NG, 3):
        img = Z[f"img_{k}"].astype(np.float64)
        assert np.array_equal(np.asarray(PO.moments(img), dtype=np.float64), Z[f"p0_{k}"])
        assert np.array_equal(PO.gauss2dfit(img, raw=True), Z[f"popt_{k}"])


def _same_gaussian(got, want, n):
    """Parameters describe the same Gaussian: curve_fit's result among the equivalent optima (sign of sigma, theta + k pi/2 with
    the sigmas swapped) depends on MINPACK's path; height, amplitude, centre and the widths with their axis directions do not."""
    # curve_fit stops at ftol = xtol = 1.5e-8 (relative cost change): its parameters sit ~1e-5 from the optimum the kernel runs to
    np.testing.assert_allclose(got[:4], want[:4], rtol=2e-4, atol=2e-4)

    def axes(p):
        sx, sy, th = abs(p[4]), abs(p[5]), p[6]
        major, minor = max(sx, sy), min(sx, sy)
        phi = (-th if sx >= sy else -th + np.pi / 2) % np.pi           # direction of the major axis in (column, row) space
        return major, minor, phi
    a, b = axes(got), axes(want)
    assert abs(a[0] - b[0]) <= 5e-4 * b[0] and abs(a[1] - b[1]) <= 5e-4 * b[1]
    if (b[0] - b[1]) / b[0] > 1e-3:
        d = abs(a[2] - b[2]) % np.pi
        assert min(d, np.pi - d) < 1e-3 / ((b[0] - b[1]) / b[0])


@pytest.mark.parametrize("k", range(NG))
def test_kernel_gauss_fit_finds_the_reference_optimum(k):
    img = Z[f"img_{k}"]
    want = Z[f"popt_{k}"]
    got = H.gaussfit(img[None])[0]
    assert got[9] == 0, "not converged"
    _same_gaussian(got[:7], want, img.shape[0])
    d = img.astype(np.float64)
    n = d.shape[0]
    X, Y = np.meshgrid(np.linspace(0, n - 1, n), np.linspace(0, n - 1, n))
    cost_ref = ((PO.twodgaussian((X, Y), *want) - d.ravel()) ** 2).sum()
    assert got[7] <= cost_ref * (1 + 1e-9)              # run to the optimum: never above curve_fit's cost
    assert got[7] >= cost_ref * (1 - 1e-6)
    # ... and on the golden set MINPACK's path and this one mostly end on the very same representative of the optimum


def test_kernel_gauss_fit_start_values_and_failures():
    img = Z["img_1"]
    got = H.gaussfit(img[None], p0=Z["p0_1"][None])[0]
    _same_gaussian(got[:7], Z["popt_1"], img.shape[0])
    bad = img.copy()
    bad[48, 10] = np.nan                                 # central row: gaussfitter.py:76-77 "something is nan"
    assert H.gaussfit(bad[None])[0][9] == 2


@pytest.mark.parametrize("maxiters", [5, None, 1])
def test_clipped_stats(maxiters):
    x = _with_stars(3, 96, 5)
    got = H.clipped_stats(x, 3.0, maxiters)
    for b in range(3):
        m, med, s = PO.clipped(x[b], 3.0, maxiters)
        np.testing.assert_allclose(got[b, 6:9], [m, med, s], rtol=1e-12)
        d = x[b].astype(np.float64)
        np.testing.assert_allclose(got[b, :4], [d.size, d.mean(), np.median(d), d.std()], rtol=1e-12)
        assert got[b, 9] == d.max()
    x64 = x.astype(np.float64) + 1e-9
    got = H.clipped_stats(x64[:1], 3.0, maxiters)
    np.testing.assert_allclose(got[0, 6:9], PO.clipped(x64[0], 3.0, maxiters), rtol=1e-12)


@pytest.mark.parametrize("n,seed", [(64, 1), (96, 2), (141, 3)])
def test_detect_sources_and_maskstars(n, seed):
    x = _with_stars(4, n, seed)
    res = H.maskstars(x, seed=1234 + seed)
    cleaned_any = 0
    for b in range(4):
        d = x[b].astype(np.float64)
        lab, nseg = PO.detect_sources(d, PO.detect_threshold(d, 1.5), 8)
        assert res["info"][b, 3] == 0 and res["info"][b, 0] == nseg
        assert np.array_equal(res["labels"][b], lab), "segmentation map differs from scipy.ndimage.label"
        want, info = PO.maskstarsSEG(d, seed=1234 + seed, cutout=b, return_info=True)
        assert (res["info"][b, 1], res["info"][b, 2]) == (info["n_cleaned"], info["n_replaced"])
        changed = want != d
        assert np.array_equal(res["clean"][b].astype(np.float64) != d, changed)
        np.testing.assert_allclose(res["clean"][b][changed], want[changed].astype(np.float32), rtol=1e-6)
        assert np.array_equal(res["clean"][b][~changed], x[b][~changed])
        cleaned_any += info["n_cleaned"]
    assert cleaned_any > 0, "the fixture has no external source: the test would be vacuous"
    # the central object survives: its pixels are not painted over
    c = n // 2
    assert np.array_equal(res["clean"][:, c - 2:c + 3, c - 2:c + 3], x[:, c - 2:c + 3, c - 2:c + 3])


def test_philox_stream_is_the_published_one():
    """Known-answer test of Philox4x32-10 (Random123 kat_vectors: counter = key = 0, and the all-ones vector)."""
    assert [int(v) for v in PO.philox4x32((0, 0, 0, 0), (0, 0))] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert [int(v) for v in PO.philox4x32((0xffffffff,) * 4, (0xffffffff, 0xffffffff))] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    v = PO.philox_normal(7, 3, np.arange(200000))
    assert abs(v.mean()) < 0.01 and abs(v.std() - 1.0) < 0.01


@pytest.mark.parametrize("k", [1, 5, 40])
def test_clipped_stats_two_stopping_rules_from_one_run(k):
    x = _with_stars(3, 96, 9)
    first, conv = H.clipped_stats_both(x, 3.0, k)
    assert np.array_equal(first, H.clipped_stats(x, 3.0, k), equal_nan=True)
    assert np.array_equal(conv, H.clipped_stats(x, 3.0, None), equal_nan=True)
