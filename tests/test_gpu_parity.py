"""GPU parity gate (run on a B200 with `pytest -m gpu`): the CUDA library, through the C ABI and the
package's Python API, against (1) the golden vectors produced by the unmodified reference,
(2) the CPU oracle on seeded synthetic cutouts, (3) size-independent properties at full size."""
import numpy as np
import pytest

from tests import parity_checks as P
from tests.golden_util import compare_row
from tests.runners import gpu_runner as run

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["sdss_s", "sdss_l", "synth"])
def test_golden_sets(name):
    """All 69 + 69 bundled SDSS cutouts and the 37 synthetic reference cases (incl. 512^2)."""
    g = P.golden(name)
    for i in range(g.n):
        P.check_golden_case(run, name, i)


def test_golden_batched_equals_single():
    """One launch over a ragged-cost batch gives the rows of the one-by-one runs, bitwise."""
    g = P.golden("sdss_s")
    imgs = np.stack([g.image(i).astype(np.float32) for i in range(g.n)])
    sky = np.array([g.sky(i)[0] for i in range(g.n)])
    err = np.array([g.sky(i)[1] for i in range(g.n)])
    res = run(imgs, sky, err, 3, 15)
    for i in range(0, g.n, 7):
        one = run(imgs[i:i + 1], sky[i:i + 1], err[i:i + 1], 3, 15)
        assert np.array_equal(one["rows"][0], res["rows"][i])
        assert np.array_equal(one["segmap"][0], res["segmap"][i])


@pytest.mark.parametrize("n,fs,seed,B", [(48, 3, 1, 6), (64, 5, 2, 6), (96, 3, 3, 6), (75, 7, 4, 4), (140, 3, 5, 6),
                                         (141, 3, 6, 4), (64, 1, 7, 3), (231, 3, 8, 3), (256, 3, 9, 6), (256, 5, 10, 2)])
def test_oracle_synth(n, fs, seed, B):
    P.check_against_oracle(run, P.synth_batch(B, n, seed), 100.0, 5.0, fs)


def test_float64_input():
    x = P.synth_batch(3, 96, 21).astype(np.float64) + 1e-9 * np.random.default_rng(0).standard_normal((3, 96, 96))
    P.check_against_oracle(run, x, 100.0, 5.0, 3, dtype=np.float64)


def test_exact_filter_path_matches_fast_path():
    import pawlikmorph_lsst_b200 as pm
    x = P.synth_batch(4, 128, 31)
    a = run(x, 100.0, 5.0, 3, 15)
    b = run(x, 100.0, 5.0, 3, 15 | pm._abi.PM_FORCE_EXACT_FILTER)
    assert np.array_equal(a["segmap"], b["segmap"]) and np.array_equal(a["rows"], b["rows"])


def test_threshold_ties_take_the_exact_path():
    from oracle import pm_oracle as O
    x = np.rint(P.synth_batch(1, 66, 41)[0]).astype(np.float32)
    z = O.mean_filter_shrink(x.astype(np.float64), 3)
    cands = np.sort(z[z > 104.0].ravel())
    thr = float(cands[len(cands) // 3])
    res = run(x[None], 0.0, 0.0, 3, 15, threshold_in=[thr])
    assert np.array_equal(res["segmap"][0], O.pixelmap(x.astype(np.float64), thr, 3).astype(np.uint8))


def test_flag_subsets_and_faint_centre():
    x = P.synth_batch(3, 64, 51)
    for flags in (1, 2, 3, 4, 5, 7, 12):
        P.check_against_oracle(run, x, 100.0, 5.0, 3, flags, check_sorted=False)
    res = run(x, 100.0, 1e6, 3, 15)
    assert np.all(res["rows"][:, 14] == 1) and np.all(res["rows"][:, :13] == -99) and not res["segmap"].any()


def test_full_size_properties():
    """At the bench size (256^2, thousands of cutouts) the oracle is too slow; check properties instead:
    determinism (two runs bitwise equal), order independence (a permuted batch gives permuted rows),
    point symmetry (a point-symmetric image about the pixel the search returns has A + Abgr == 0 before
    correction, As == 0), and sortedness of the canonical order."""
    import torch

    import pawlikmorph_lsst_b200 as pm
    import synth
    B, n = 2048, 256
    x = synth.make_batch(B, n, 777, device="cuda")
    r1 = pm.batch.analyse_batch(x, 100.0, 5.0, 3, 15, want_segmap=True).rows.clone()
    r2 = pm.batch.analyse_batch(x, 100.0, 5.0, 3, 15).rows.clone()
    assert torch.equal(r1, r2)
    perm = torch.randperm(B, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
    r3 = pm.batch.analyse_batch(x[perm].contiguous(), 100.0, 5.0, 3, 15).rows
    assert torch.equal(r3, r1[perm])
    rows = r1.cpu().numpy()
    ok = rows[:, 14] == 0
    assert ok.mean() > 0.99
    for name, lo, hi in (("A", -1.0, 2.0), ("As", 0.0, 1.0), ("As90", 0.0, 1.0), ("gini", 0.0, 1.0), ("C", 0.0, 10.0),
                         ("m20", -7.0, 0.0)):
        v = rows[ok, P.FIELDS.index(name)]
        assert np.all(np.isfinite(v)) and v.min() >= lo and v.max() <= hi, (name, v.min(), v.max())
    r20, r80, rmax = (rows[ok, P.FIELDS.index(k)] for k in ("r20", "r80", "rmax"))
    assert np.all((0 < r20) & (r20 < r80) & (r80 <= rmax))
    # sorted order: values along sorted_idx are non-increasing, ties by descending index
    sub = pm.batch.analyse_batch(x[:64], 100.0, 5.0, 3, 15, want_segmap=True, want_sorted=True)
    sidx, npos = sub.sorted_idx.cpu().numpy(), sub.n_positive.cpu().numpy()
    xs = x[:64].cpu().numpy()
    for i in range(64):
        idx = sidx[i][:npos[i]]
        v = xs[i].ravel()[idx]
        assert np.all(np.diff(v) <= 0)
        tie = np.diff(v) == 0
        assert np.all(np.diff(idx)[tie] < 0)
        assert sub.segmap[i].cpu().numpy().ravel()[idx].all()


def test_point_symmetric_image():
    """A image that is point-symmetric about an integer pixel: the asymmetry about that pixel is 0 and
    the shape asymmetry of the (symmetric) map is 0 — the known-answer test of SURVEY §4."""
    import pawlikmorph_lsst_b200 as pm
    n, c = 129, 64
    yy, xx = np.mgrid[0:n, 0:n].astype(np.float64)
    r2 = ((xx - c) / 9.0) ** 2 + ((yy - c) / 5.0) ** 2
    img = (1000.0 * np.exp(-r2 / 2.0) + 100.0).astype(np.float32)
    seg = pm.pixmap.pixelmap(img.astype(np.float64), 105.0, 3)
    assert np.array_equal(seg, seg[::-1, ::-1])
    ap = pm.apertures.aperpixmap(n, pm.pixmap.calcRmax(img - 100.0, seg), 9, 0.1)
    A = pm.asymmetry.calcA((img - 100.0).astype(np.float64), seg, ap, [c, c], 180.0)
    As = pm.asymmetry.calcA(seg, seg, ap, [c, c], 180.0)
    assert A[0] == 0.0 and A[1] == 0 and As[0] == 0.0


def test_per_cutout_api_matches_oracle_functions():
    """The reference-signature functions (pixmap / apertures / asymmetry / casgm), one by one."""
    import pawlikmorph_lsst_b200 as pm
    from oracle import pm_oracle as O
    for n, seed in ((96, 61), (141, 62)):
        img_sky = P.synth_batch(1, n, seed)[0].astype(np.float64)
        seg = pm.pixmap.pixelmap(img_sky, 105.0, 3)
        assert np.array_equal(seg, O.pixelmap(img_sky, 105.0, 3))
        assert pm.pixmap.checkPixelmapEdges(seg) == O.checkPixelmapEdges(seg)
        img = img_sky - 100.0
        rmax = pm.pixmap.calcRmax(img, seg)
        assert rmax == O.calcRmax(img, seg)
        ap = pm.apertures.aperpixmap(n, rmax, 9, 0.1)
        assert np.array_equal(ap, O.aperpixmap(n, rmax, 9, 0.1))
        ap_f = pm.apertures.aperpixmap(n, 17.3, 9, 0.1)       # float-form radius
        assert np.array_equal(ap_f, O.aperpixmap(n, 17.3, 9, 0.1))
        apix = pm.asymmetry.minapix(img, seg, ap, np.ones_like(img))
        assert apix == O.minapix(img, seg, ap, np.ones_like(img))
        for angle, nc in ((180.0, True), (180.0, False), (90.0, False)):
            got = pm.asymmetry.calcA(img, seg.copy(), ap, apix, angle, None, nc)
            want = O.calcA(img, seg.copy(), ap, apix, angle, None, nc)
            np.testing.assert_allclose(got, want, rtol=1e-9)
        np.testing.assert_allclose(pm.asymmetry.calcA(seg, seg, ap, apix, 90.0), O.calcA(seg, seg, ap, apix, 90.0), rtol=1e-12)
        np.testing.assert_allclose(pm.asymmetry.calculateAsymmetries(img, seg), O.calculateAsymmetries(img, seg.copy()), rtol=1e-9)
        r20, r80 = pm.casgm.calcR20_R80(img, seg, apix, rmax)
        w20, w80 = O.calcR20_R80(img, seg, apix, rmax)
        np.testing.assert_allclose([r20, r80], [w20, w80], rtol=1e-9)
        assert pm.casgm.concentration(r20, r80) == O.concentration(r20, r80)
        np.testing.assert_allclose(pm.casgm.gini(img, seg), O.gini(img, seg), rtol=1e-10)
        np.testing.assert_allclose(pm.casgm.m20(img, seg), O.m20(img, seg), rtol=1e-10)
        np.testing.assert_allclose(pm.casgm.smoothness(img, seg, apix, rmax, w20, 100.0),
                                   O.smoothness(img, seg, apix, rmax, w20, 100.0), rtol=1e-8)
        np.testing.assert_array_equal(pm.apertures.distarr(n, n, np.array([3, 5])), O.distarr(n, n, np.array([3, 5])))
        for npix_, ns_ in ((n, 9), (15, 9), (16, 3)):          # odd and even sizes (apertures.py:96-114 reads the first n blocks)
            c_ = np.array([npix_ // 2 + 1, npix_ // 2 + 1])
            np.testing.assert_array_equal(pm.apertures.subdistarr(npix_, ns_, c_), O.subdistarr(npix_, ns_, c_))
        np.testing.assert_array_equal(pm.apertures.apercentre(ap, np.array(apix)), O.apercentre(ap, np.array(apix)))
    with pytest.raises(AttributeError):
        pm.pixmap.pixelmap(img_sky, 1e9, 3)          # "Central pixel too faint" (pixmap.py:178-179)


def test_host_api_and_errors():
    import torch

    import pawlikmorph_lsst_b200 as pm
    x = P.synth_batch(5, 64, 71)
    rows = pm.batch.analyse_host(x, 100.0, 5.0, 3, 15, chunk=2)
    ref = run(x, 100.0, 5.0, 3, 15)["rows"]
    assert np.array_equal(rows, ref)
    with pytest.raises(ValueError):
        pm.batch.analyse_batch(torch.zeros((1, 64, 64), device="cuda"), 0.0, 0.0, 4)      # even filter size
    with pytest.raises(TypeError):
        pm.batch.analyse_batch(torch.zeros((1, 64, 64)), 0.0, 0.0, 3)                     # not on the GPU
    assert pm.batch.analyse_batch(torch.zeros((0, 64, 64), device="cuda"), 0.0, 0.0, 3).rows.shape == (0, 16)


def test_star_masks_in_the_per_cutout_api():
    """starMask arguments of pixelmap / minapix / calcA / calcMaskedFraction (pixmap.py:156-158, :58-88;
    asymmetry.py:82-86, :182-196, :219) against the oracle."""
    import pawlikmorph_lsst_b200 as pm
    from oracle import pm_oracle as O
    n = 96
    img_sky = P.synth_batch(1, n, 81)[0].astype(np.float64)
    star = np.ones((n, n))
    star[20:34, 50:70] = 0.0          # a masked star off-centre
    star[52:58, 30:36] = 0.0          # and one inside the galaxy
    seg = pm.pixmap.pixelmap(img_sky, 105.0, 3, star)
    assert np.array_equal(seg, O.pixelmap(img_sky, 105.0, 3, star))
    img = img_sky - 100.0
    rmax = pm.pixmap.calcRmax(img, seg)
    ap = pm.apertures.aperpixmap(n, rmax, 9, 0.1)
    apix = pm.asymmetry.minapix(img, seg, ap, star)
    assert apix == O.minapix(img, seg, ap, star)
    for angle, nc in ((180.0, False), (90.0, False)):
        got = pm.asymmetry.calcA(img, seg.copy(), ap, apix, angle, star, nc)
        want = O.calcA(img, seg.copy(), ap, apix, angle, star, nc)
        np.testing.assert_allclose(got, want, rtol=1e-9)
    # noisecorrect=True with a star mask: upstream multiplies the segmap by starMask * rotate(starMask) in place and
    # dilates it (asymmetry.py:219-220); skimage's bilinear rotation leaves ~1e-14 instead of 0 next to mask edges and
    # binary_dilation counts those as set.  The library uses the exact 0/1 rotated mask, so the
    # checker is the oracle fed with the exactly masked inputs.
    import torch
    from pawlikmorph_lsst_b200 import _single
    sm = _single.rotated_mask(torch.from_numpy(star).cuda(), 180.0, apix[0], apix[1]).cpu().numpy()
    s1 = seg.copy()
    got = pm.asymmetry.calcA(img, s1, ap, apix, 180.0, star, True)
    want = O.calcA(img * sm, seg * sm, ap, apix, 180.0, None, True)
    np.testing.assert_allclose(got, want, rtol=1e-9)
    assert np.array_equal(s1, seg * sm)                      # the in-place segmap update of asymmetry.py:219
    np.testing.assert_allclose(pm.asymmetry.calcA(seg, seg, ap, apix, 90.0, star), O.calcA(seg, seg, ap, apix, 90.0, star),
                               rtol=1e-12)
    np.testing.assert_allclose(pm.pixmap.calcMaskedFraction(seg, star, apix), O.calcMaskedFraction(seg, star, apix), rtol=1e-12)


def test_pruned_minapix():
    """The production configuration (no per-candidate output -> exact pruning in minapix) on every golden case and
    against the unpruned evaluation at bench size."""
    import torch

    import pawlikmorph_lsst_b200 as pm
    import synth
    from tests.runners import gpu_runner_pruned
    for name in ("sdss_s", "sdss_l", "synth"):
        g = P.golden(name)
        for i in range(g.n):
            P.check_golden_case(gpu_runner_pruned, name, i)
    x = synth.make_batch(4096, 256, 778, device="cuda")
    a = pm.batch.analyse_batch(x, 100.0, 5.0, 3, 15).rows
    b = pm.batch.analyse_batch(x, 100.0, 5.0, 3, 15 | pm._abi.PM_NO_PRUNE).rows
    assert torch.equal(a, b)


def test_maximum_size_and_supplied_segmaps():
    """n = PM_MAX_N (640) with filter sizes 1 and 15, and the -segmap path (main.py:428-439): a supplied map
    must give exactly the rows of the computed one."""
    import pawlikmorph_lsst_b200 as pm
    rng = np.random.default_rng(123)
    n = 640
    yy, xx = np.mgrid[0:n, 0:n]
    img = (100 + rng.normal(0, 5, (n, n)) + 900 * np.exp(-(((xx - 321) / 7.0) ** 2 + ((yy - 318) / 5.0) ** 2) / 2)).astype(np.float32)
    for fs in (1, 15):
        P.check_against_oracle(run, img[None], 100.0, 5.0, fs)
    x = P.synth_batch(4, 256, 97)
    a = run(x, 100.0, 5.0, 3, 15)
    b = run(x, 100.0, 5.0, 3, 15, segmap_in=a["segmap"])
    rows_a, rows_b = a["rows"].copy(), b["rows"].copy()
    rows_a[:, 13] = 0      # object_edge is only set on the pixelmap path (main.py:427)
    assert np.array_equal(rows_a, rows_b)


def test_small_radius_growth_curves():
    P.check_small_radius_growth_curves(run)


def test_star_masks_in_the_batch_api():
    """batch.analyse_batch_starmask: the -starmask flow (main.py:395-413, :463-495) for a batch, against the same chain of
    oracle calls (exact 0/1 rotated masks, like test_star_masks_in_the_per_cutout_api)."""
    import torch

    import pawlikmorph_lsst_b200 as pm
    from oracle import pm_oracle as O
    from pawlikmorph_lsst_b200 import _single
    n, B = 96, 4
    x = P.synth_batch(B, n, 83)
    rng = np.random.default_rng(83)
    star = np.ones((B, n, n), np.uint8)
    for b in range(B):
        for _ in range(3):
            r0, c0 = rng.integers(5, n - 20, 2)
            star[b, r0:r0 + rng.integers(4, 14), c0:c0 + rng.integers(4, 14)] = 0
        star[b, n // 2 - 3:n // 2 + 4, n // 2 - 3:n // 2 + 4] = 1          # the centre stays visible
    res = pm.batch.analyse_batch_starmask(torch.from_numpy(x).cuda(), 100.0, 5.0, torch.from_numpy(star).cuda(), 3, 15)
    rows, segs = res.rows.cpu().numpy(), res.segmap.cpu().numpy()
    for b in range(B):
        d = x[b].astype(np.float64)
        st = star[b].astype(np.float64)
        seg = O.pixelmap(d, 105.0, 3, st)
        assert np.array_equal(segs[b], seg.astype(np.uint8))
        im = d - 100.0
        rmax = O.calcRmax(im, seg)
        ap = O.aperpixmap(n, rmax, 9, 0.1)
        apix = O.minapix(im, seg, ap, st)
        s180 = _single.rotated_mask(torch.from_numpy(st).cuda(), 180.0, apix[0], apix[1]).cpu().numpy()
        s90 = _single.rotated_mask(torch.from_numpy(st).cuda(), 90.0, apix[0], apix[1]).cpu().numpy()
        seg2 = seg * s180
        A = O.calcA(im * s180, seg2.copy(), ap, apix, 180.0, None, True)
        As = O.calcA(seg2, seg2, ap, apix, 180.0, None)[0]
        As90 = O.calcA(seg2 * s90, seg2, ap, apix, 90.0, None)[0]
        r20, r80 = O.calcR20_R80(im, seg2, apix, rmax)
        want = dict(apix_col=apix[0], apix_row=apix[1], rmax=rmax, r20=r20, r80=r80, C=O.concentration(r20, r80), A=A[0], Abgr=A[1],
                    S=O.smoothness(im, seg2, apix, rmax, r20, 100.0) if int(r20) >= 1 else -99.0, As=As,      # uniform_filter(size=0) raises upstream
                    As90=As90, gini=O.gini(im, seg2), m20=O.m20(im, seg2),
                    object_edge=float(O.checkPixelmapEdges(seg)), status=0.0)
        errs = compare_row(P.rowd(rows, b), want, 1e-8, f"[{b}] ")
        assert not errs, "\n".join(errs)
    # all-ones masks give the plain path
    plain = pm.batch.analyse_batch(torch.from_numpy(x).cuda(), 100.0, 5.0, 3, 15).rows.cpu().numpy()
    ones = pm.batch.analyse_batch_starmask(torch.from_numpy(x).cuda(), 100.0, 5.0, torch.ones((B, n, n), dtype=torch.uint8).cuda(), 3, 15)
    np.testing.assert_allclose(ones.rows.cpu().numpy(), plain, rtol=1e-9, equal_nan=True)
