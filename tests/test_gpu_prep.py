"""GPU tests of the pre-processing kernels: clipped statistics, the Gaussian fit against the vectors of
the reference's own gaussfitter.py, source detection against scipy.ndimage.label, the seeded maskstarsSEG, the whole of
skybgr (incl. the -li mode), and the driver / CLI with the reference's default behaviour (device sky, cleaning)."""
import csv
import os

import numpy as np
import pytest

import synth
from oracle import pm_oracle as O
from oracle import prep_oracle as PO
from oracle import sky_oracle as SO
from tests.golden_util import GOLDEN_DIR, compare_row
from tests.test_prep import _same_gaussian, _with_stars

pytestmark = pytest.mark.gpu
Z = np.load(os.path.join(GOLDEN_DIR, "gaussfit.npz"))


def _cuda(x):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def test_clipped_stats_on_the_device():
    import pawlikmorph_lsst_b200 as pm
    x = _with_stars(6, 128, 11)
    for maxiters in (5, None):
        got = pm.skyBackground.clipped_stats_batch(_cuda(x), 3.0, maxiters).cpu().numpy()
        for b in range(6):
            np.testing.assert_allclose(got[b, 6:9], PO.clipped(x[b], 3.0, maxiters), rtol=1e-12)
            assert got[b, 9] == x[b].max() and got[b, 2] == np.median(x[b].astype(np.float64))


def test_gauss_fit_golden_vectors_on_the_device():
    import pawlikmorph_lsst_b200 as pm
    for k in range(int(Z["count"])):
        img = Z[f"img_{k}"]
        got = pm.skyBackground.gaussfit_batch(_cuda(img[None])).cpu().numpy()[0]
        assert got[9] == 0
        _same_gaussian(got[:7], Z[f"popt_{k}"], img.shape[0])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_detect_sources_and_maskstars_on_the_device(dtype):
    import pawlikmorph_lsst_b200 as pm
    x = _with_stars(8, 141, 21).astype(dtype)
    res = pm.imageutils.maskstars_batch(_cuda(x), seed=99, want_labels=True)
    info, labels, clean = res["info"].cpu().numpy(), res["labels"].cpu().numpy(), res["clean"].cpu().numpy()
    ncl = 0
    for b in range(8):
        d = x[b].astype(np.float64)
        lab, nseg = PO.detect_sources(d, PO.detect_threshold(d, 1.5), 8)
        assert info[b, 0] == nseg and info[b, 3] == 0 and np.array_equal(labels[b], lab)
        want, inf = PO.maskstarsSEG(d, seed=99, cutout=b, return_info=True)
        assert (info[b, 1], info[b, 2]) == (inf["n_cleaned"], inf["n_replaced"])
        ch = want != d
        assert np.array_equal(clean[b].astype(np.float64) != d, ch)
        np.testing.assert_allclose(clean[b][ch], want[ch].astype(dtype), rtol=1e-6 if dtype == np.float32 else 1e-12)
        ncl += inf["n_cleaned"]
    assert ncl > 0
    assert np.array_equal(pm.imageutils.maskstarsSEG(x[0], seed=99), clean[0].astype(np.float64)) or dtype == np.float32


def test_skybgr_batch_against_the_oracle():
    import pawlikmorph_lsst_b200 as pm
    x = _with_stars(6, 128, 31)
    r = pm.skyBackground.skybgr_batch(_cuda(x))
    assert (r["status"] == 0).all()
    for b in range(6):
        d = x[b].astype(np.float64)
        _, fw, th = PO.fit_gauss(d)                                          # curve_fit
        assert abs(max(r["fwhms"][b]) - max(fw)) < 1e-3 * max(fw) and abs(min(r["fwhms"][b]) - min(fw)) < 1e-3 * min(fw)
        row = SO.calc_sky(d, list(r["fwhms"][b]), float(r["theta"][b]))       # the statistics for the fit the device found
        assert abs(r["sky"][b] - row["sky"]) <= 1e-9 * abs(row["sky"]) and abs(r["sky_err"][b] - row["sky_err"]) <= 1e-9 * row["sky_err"]
        sky_ref = PO.calc_skybgr(d)[0]                                       # the whole of the oracle, its own fit
        assert abs(r["sky"][b] - sky_ref) < 5e-3 * abs(sky_ref)              # differs only through theta's representative
    # per-image function and its errors
    sky, err, fw, th = pm.skyBackground.skybgr(x[0].astype(np.float64))
    assert (sky, err) == (r["sky"][0], r["sky_err"][0])
    flat = np.full((64, 64), 100.0) + np.random.default_rng(1).normal(0, 1e-3, (64, 64))
    with pytest.raises(AttributeError):
        pm.skyBackground.skybgr(flat)                                        # "No object in image" (skyBackground.py:112-113)


def test_skybgr_large_image_mode():
    """-li (skyBackground.py:71-76): sources and sky region on the CLEANED larger cutout, the Gaussian on the small one."""
    import pawlikmorph_lsst_b200 as pm
    big = _with_stars(3, 192, 41)
    small = np.ascontiguousarray(big[:, 32:160, 32:160])
    r = pm.skyBackground.skybgr_batch(_cuda(small), _cuda(big), seed=5)
    assert (r["status"] == 0).all()
    for b in range(3):
        cleaned = PO.maskstarsSEG(big[b].astype(np.float64), seed=5, cutout=b)
        row = SO.calc_sky(cleaned, list(r["fwhms"][b]), float(r["theta"][b]))
        assert abs(r["sky"][b] - row["sky"]) <= 1e-6 * abs(row["sky"])        # float32 cleaned pixels on the device
    r0 = pm.skyBackground.skybgr_batch(_cuda(small))
    assert not np.array_equal(r0["sky"], r["sky"])


def test_driver_default_is_the_reference_flow(tmp_path):
    """calcMorphology with its defaults = main.py:349-495: device skybgr, pixelmap on the image with the sky in, sky
    subtraction, seeded maskstarsSEG, statistics on the cleaned image — against the same chain of oracle calls."""
    import pawlikmorph_lsst_b200 as pm
    from tests.fits_files import fits_bytes
    x = np.rint(_with_stars(5, 141, 51)).astype(np.float32)
    infos = []
    for k in range(5):
        (tmp_path / f"g{k}.fits").write_bytes(fits_bytes(x[k], 16))
        infos.append((f"g{k}.fits", 1.0, 2.0))
    res = pm.main.calcMorphology(infos, tmp_path / "out", root=tmp_path, seed=17)
    for k, r in enumerate(res):
        d = x[k].astype(np.float64)
        assert r.status == 0 and r.fwhms[0] > 0
        row = SO.calc_sky(d, list(r.fwhms), r.theta)
        assert abs(r.sky - row["sky"]) <= 1e-9 * abs(row["sky"])
        seg = O.pixelmap(d, r.sky + r.sky_err, 3)
        cleaned = PO.maskstarsSEG(d - r.sky, seed=17, cutout=k)
        want, _, _ = O.analyse(cleaned, 0.0, r.sky_err, 3, 7, segmap_in=seg)
        got = dict(apix_col=r.apix[0], apix_row=r.apix[1], rmax=r.rmax, r20=r.r20, r80=r.r80, C=r.C, A=r.A[0], Abgr=r.A[1], S=r.S,
                   As=r.As[0], As90=r.As90[0], gini=r.gini, m20=r.m20, object_edge=float(O.checkPixelmapEdges(seg)), status=r.status)
        want["object_edge"] = float(O.checkPixelmapEdges(seg))
        errs = compare_row(got, want, 1e-8, f"[{k}] ")
        assert not errs, "\n".join(errs)
        assert r.objectEdge == bool(O.checkPixelmapEdges(seg))
    # a frame without any source: default Result, the run carries on (main.py:350-368)
    flat = (1000.0 + np.random.default_rng(2).normal(0, 1e-3, (141, 141))).astype(np.float32)       # nothing above mean + 1.5 sigma in 8 connected pixels
    (tmp_path / "flat.fits").write_bytes(fits_bytes(flat, -32))
    res2 = pm.main.calcMorphology(infos[:1] + [("flat.fits", 1.0, 2.0)], tmp_path / "out2", root=tmp_path, seed=17)
    assert res2[0].A == res[0].A and res2[1].status == 1 and res2[1].rmax == -99 and res2[1].sky == -99.


def test_command_line(tmp_path, monkeypatch):
    """imganalysis.py with the README's flags: `-f images.csv -fo out -Aall -cas`."""
    import imganalysis
    from tests.fits_files import fits_bytes
    x = np.rint(_with_stars(3, 141, 61)).astype(np.float32)
    (tmp_path / "sample").mkdir()
    lines = ["filename, ra, dec"]
    for k in range(3):
        (tmp_path / "sample" / f"c{k}.fits").write_bytes(fits_bytes(x[k], 16))
        lines.append(f"sample/c{k}.fits, 149.0, 0.5")
    (tmp_path / "images.csv").write_text("\n".join(lines) + "\n")
    monkeypatch.chdir(tmp_path)
    assert imganalysis.main(["-f", "images.csv", "-fo", "out", "-Aall", "-cas", "-s", "200", "--seed", "3"]) == 0
    table = list(csv.reader(open(tmp_path / "out" / "parameters.csv")))
    assert len(table) == 4 and [r[0] for r in table[1:]] == ["c0.fits", "c1.fits", "c2.fits"]
    d = dict(zip(table[0], table[1]))
    assert float(d["A"]) != -99.0 and float(d["As"]) != -99.0 and float(d["C"]) != -99.0 and float(d["S"]) == -99.0 and float(d["sky"]) > 0
    assert imganalysis.main(["-f", "images.csv", "-fo", "out2", "-A", "-s", "200"]) == 0          # -A alone works here (not upstream)
    t2 = list(csv.reader(open(tmp_path / "out2" / "parameters.csv")))
    d2 = dict(zip(t2[0], t2[1]))
    assert float(d2["A"]) == float(d["A"]) and float(d2["As"]) == -99.0 and float(d2["C"]) == -99.0
    with pytest.raises(NotImplementedError):
        imganalysis.main(["-f", "images.csv", "-fo", "out", "-sersic"])
