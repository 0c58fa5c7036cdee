"""The file driver end to end on the GPU: the reference's 69 + 69 SDSS frames (rewritten as FITS files from the golden
fixtures, BITPIX 16 / BZERO 32768 like sample/) -> calcMorphology -> parameters.csv, against the golden rows."""
import ast
import os
import csv

import numpy as np
import pytest

from tests import parity_checks as P
from tests.fits_files import fits_bytes
from tests.golden_util import compare_row

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["sdss_s", "sdss_l"])
def test_calcMorphology_on_the_sdss_frames(name, tmp_path):
    import pawlikmorph_lsst_b200 as pm
    g = P.golden(name)
    (tmp_path / "sample").mkdir()
    lines = ["filename, ra, dec"]
    for i in range(g.n):
        fn = f"sample/{str(g.z['names'][i])}"
        (tmp_path / fn).write_bytes(fits_bytes(g.image(i), -32 if i % 7 == 3 else 16))
        lines.append(f"{fn}, {149.0 + i}, {0.5}")
    (tmp_path / "images.csv").write_text("\n".join(lines) + "\n")
    out = tmp_path / "output"
    res = pm.main.calcMorphology(pm.helpers.getFiles(str(tmp_path / "images.csv")), out, filterSize=3, allAsymmetry=True, CAS=True,
                                 smoothness=True, root=tmp_path, sky="border", clean=False)
    assert len(res) == g.n
    table = list(csv.reader(open(out / "parameters.csv")))
    assert table[0] == pm.result.CSV_HEADER and len(table) == g.n + 1
    for i, r in enumerate(res):
        want = g.row(i)
        assert (r.sky, r.sky_err) == g.sky(i)
        got = dict(apix_col=r.apix[0], apix_row=r.apix[1], rmax=r.rmax, r20=r.r20, r80=r.r80, C=r.C, A=r.A[0], Abgr=r.A[1], S=r.S,
                   As=r.As[0], As90=r.As90[0], gini=r.gini, m20=r.m20, object_edge=float(r.objectEdge), status=r.status)
        errs = compare_row(got, want, P.RTOL, f"{name}[{i}] ")
        assert not errs, "\n".join(errs)
        d = dict(zip(table[0], table[i + 1]))                      # the CSV carries the same numbers (repr round-trips)
        assert d["file"] == r.file and list(ast.literal_eval(d["apix"])) == list(r.apix)
        for col, v in (("r_max", r.rmax), ("C", r.C), ("A", r.A[0]), ("Abgr", r.A[1]), ("As", r.As[0]), ("As90", r.As90[0]),
                       ("Gini index", r.gini), ("M20", r.m20), ("r20", r.r20), ("r80", r.r80)):
            assert float(d[col]) == float(v) or (float(d[col]) != float(d[col]) and v != v)
        assert d["Object on image edge?"] == str(r.objectEdge)
    # -A only (main.py:469-476): the CAS and shape columns keep their sentinels
    res_a = pm.main.calcMorphology(list(pm.helpers.getFiles(str(tmp_path / "images.csv")))[:5], out, asymmetry=True, allAsymmetry=False,
                                   CAS=False, paramsaveFile="a.csv", root=tmp_path, sky="border", clean=False)
    for i, r in enumerate(res_a):
        assert r.A == res[i].A and r.As == [-99., -99.] and r.C == -99. and r.gini == -99. and r.S == -99.


def test_calcMorphology_cutouts_of_the_catalogue_positions(tmp_path):
    """imganalysis.py's default -s 128: the window around every row of images.csv (WCS cards of the reference's frames,
    tests/golden/sdss_wcs.json), cut and decoded on the device — bit-identical to the float host call on numpy slices, and
    within tolerance of the oracle on a few of them."""
    import json
    import os

    import pawlikmorph_lsst_b200 as pm
    from oracle import pm_oracle as O
    g = P.golden("sdss_l")
    J = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sdss_wcs.json")))
    index = {str(nm): i for i, nm in enumerate(g.z["names"])}
    infos, wins = [], []
    (tmp_path / "sample").mkdir()
    for k, e in enumerate(J):
        i = index[e["file"].split("/")[-1]]
        img = g.image(i)
        (tmp_path / e["file"]).write_bytes(fits_bytes(img, -32 if k % 5 == 2 else 16, extra=e["header"]))
        infos.append((e["file"], e["ra"], e["dec"]))
        r0, c0 = pm.image.cutout_origin(*pm.image.sky_to_pixel(e["header"], e["ra"], e["dec"]), 128, img.shape)
        wins.append(img[r0:r0 + 128, c0:c0 + 128].astype(np.float32))
    res = pm.main.calcMorphology(infos, tmp_path / "out", npix=128, smoothness=True, root=tmp_path, sky="border", clean=False)
    sky = np.array([r.sky for r in res]), np.array([r.sky_err for r in res])
    for k, r in enumerate(res):
        assert (r.sky, r.sky_err) == pm.main.border_sky(wins[k])
    rows = pm.batch.analyse_host(np.stack(wins), sky[0], sky[1], 3, 15)
    for k, r in enumerate(res):
        d = dict(zip(pm._abi.ROW_FIELDS, rows[k]))
        if d["status"] == 1:
            assert r.status == 1 and r.A == [-99., -99.]
            continue
        assert [r.apix[0], r.apix[1], r.rmax, r.r20, r.r80, r.C, r.A[0], r.A[1], r.S, r.As[0], r.As90[0], r.gini, r.m20] == \
            [d[f] for f in ("apix_col", "apix_row", "rmax", "r20", "r80", "C", "A", "Abgr", "S", "As", "As90", "gini", "m20")], infos[k][0]
    for k in (0, 17, 40):
        want, _, _ = O.analyse(wins[k].astype(np.float64), res[k].sky, res[k].sky_err, 3)
        errs = compare_row(dict(zip(pm._abi.ROW_FIELDS, rows[k])), want, P.RTOL, f"window[{k}] ")
        assert not errs, "\n".join(errs)


def test_calcMorphology_with_an_empty_list_writes_the_header(tmp_path):
    import pawlikmorph_lsst_b200 as pm
    assert pm.main.calcMorphology([], tmp_path / "o") == []
    assert list(csv.reader(open(tmp_path / "o" / "parameters.csv"))) == [pm.result.CSV_HEADER]


def test_calcMorphology_with_the_sky_measured_on_the_device(tmp_path):
    """gauss=(fwhms, theta) per image: sky / sky_err from the sky region of the decoded frame (pm_sky_stats_batch), fed to
    the analysis on the device — equal to the oracle's sky statistics, and the rows equal to the float host call given
    those numbers."""
    import pawlikmorph_lsst_b200 as pm
    from oracle import sky_oracle as SO
    g = P.golden("sdss_s")
    pick = list(range(0, 24))
    rng = np.random.default_rng(12)
    gauss = [([float(rng.uniform(3, 9)), float(rng.uniform(3, 9))], float(rng.uniform(-3, 3))) for _ in pick]
    infos = []
    for k, i in enumerate(pick):
        fn = f"g{k}.fits"
        (tmp_path / fn).write_bytes(fits_bytes(g.image(i), 16 if k % 3 else -32))
        infos.append((fn, 10.0, 1.0))
    res = pm.main.calcMorphology(infos, tmp_path / "o", gauss=gauss, smoothness=True, root=tmp_path)
    imgs = np.stack([g.image(i).astype(np.float32) for i in pick])
    for k, r in enumerate(res):
        w = SO.calc_sky(imgs[k].astype(np.float64), gauss[k][0], gauss[k][1])
        assert abs(r.sky - w["sky"]) <= 1e-9 * abs(w["sky"]) and abs(r.sky_err - w["sky_err"]) <= 1e-9 * w["sky_err"], k
    rows = pm.batch.analyse_host(imgs, np.array([r.sky for r in res]), np.array([r.sky_err for r in res]), 3, 15)
    for k, r in enumerate(res):
        d = dict(zip(pm._abi.ROW_FIELDS, rows[k]))
        if d["status"] == 1:
            assert r.status == 1
            continue
        assert [r.apix[0], r.apix[1], r.rmax, r.C, r.A[0], r.A[1], r.S, r.As[0], r.gini, r.m20] == \
            [d[f] for f in ("apix_col", "apix_row", "rmax", "C", "A", "Abgr", "S", "As", "gini", "m20")], k
    with pytest.raises(pm.skyBackground._SkyError):
        pm.main.calcMorphology(infos[:2], tmp_path / "o", gauss=[([60.0, 60.0], 0.0)] * 2, root=tmp_path)


def test_catalogue_flow(tmp_path):
    """calcMorphology(catalogue=...) — main.py:372-395: objectOccluded on the unmasked pixelmap, maskstarsPSF(adaptive=False), then the
    pixelmap and the statistics with that mask (checked against the same steps done by hand with the batch calls), star_flag and
    the masked pixel fraction in the Result."""
    import json

    import torch

    import pawlikmorph_lsst_b200 as pm
    import synth
    from tests.fits_files import fits_bytes
    A = pm._abi
    wcs = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "sdss_wcs.json")))[0]
    cards = dict(wcs["header"])
    n = cards["NAXIS1"]
    cards.update(PSF_S1=1.31, PSF_S2=2.94, EXPTIME=53.907456, PHT_AA=-24.05, PHT_KK=0.09, AIRMASS=1.19, SOFTBIAS=1000, PHT_B=1.2e-10)
    x = synth.make_batch(2, n, 314).numpy() + 1000.0                       # counts with the SDSS software bias in
    ra0, dec0 = wcs["ra"], wcs["dec"]
    px, py = pm.image.sky_to_pixel(cards, ra0, dec0)
    assert abs(px - n / 2) < 2 and abs(py - n / 2) < 2                     # the catalogue position is the frame centre
    # a bright star on the object's wing and one far away; the far one must not reach the mask
    rows_cat = ["objID,ra,dec,psfMag_r,type",
                f"1,{ra0 + 0.0015},{dec0 + 0.0005},15.5,STAR", f"2,{ra0 - 0.0060},{dec0 - 0.0065},16.0,STAR", f"3,{ra0 + 0.001},{dec0},17.0,GALAXY"]
    (tmp_path / "cat.csv").write_text("\n".join(rows_cat) + "\n")
    names = []
    for k in range(2):
        (tmp_path / f"f{k}.fits").write_bytes(fits_bytes(x[k], -32, extra=cards))
        names.append(f"f{k}.fits")
    infos = [(nm, ra0, dec0) for nm in names]
    sky = [(synth.SKY + 1000.0, synth.SKY_ERR)] * 2
    res = pm.main.calcMorphology(infos, str(tmp_path / "out"), catalogue=str(tmp_path / "cat.csv"), sky=sky, root=str(tmp_path), seed=5, smoothness=True)
    # by hand
    dev = torch.device("cuda:0")
    img = torch.from_numpy(x.astype(np.float32)).to(dev)
    sky_d = torch.full((2,), synth.SKY + 1000.0, dtype=torch.float64, device=dev)
    err_d = torch.full((2,), synth.SKY_ERR, dtype=torch.float64, device=dev)
    first = pm.batch.analyse_batch(img, sky_d, err_d, 3, 15 | A.PM_STOP_AFTER_PIXELMAP, want_segmap=True)
    tmp = first.segmap.cpu().numpy()
    masks, flags = [], []
    for k in range(2):
        flag, objs = pm.objectMasker.objectOccluded(tmp[k], (ra0, dec0), str(tmp_path / "cat.csv"), cards)
        flags.append(flag)
        masks.append(pm.imageutils.maskstarsPSF(np.zeros((n, n)), objs, cards, synth.SKY + 1000.0, 5.0, adaptive=False))
    assert flags == [True, True] and all(m.dtype == bool and (~m).sum() > 50 for m in masks)
    star = torch.as_tensor(np.stack(masks).astype(np.float64), device=dev)
    cleaned = pm.imageutils.maskstars_batch(img.to(torch.float64) - sky_d[:, None, None], seed=5)["clean"]
    want = pm.batch.analyse_batch_starmask(img, sky_d, err_d, star, 3, 15, cleaned=cleaned).rows.cpu().numpy()
    col = A.ROW_FIELDS.index
    for k, r in enumerate(res):
        assert r.status == 0 and r.star_flag is True
        assert r.A == [want[k, col("A")], want[k, col("Abgr")]] or np.allclose(r.A, [want[k, col("A")], want[k, col("Abgr")]], rtol=1e-12)
        assert np.isclose(np.ravel(r.As)[0], want[k, col("As")], rtol=1e-12) and np.isclose(r.r20, want[k, col("r20")], rtol=1e-12)
        frac = pm.pixmap.calcMaskedFraction(tmp[k], masks[k].astype(np.float64), [want[k, col("apix_col")], want[k, col("apix_row")]])
        assert 0.0 < r.maskedPixelFraction < 1.0 and np.isclose(r.maskedPixelFraction, frac, rtol=1e-12)
    # the mask matters: the plain flow gives another pixelmap-dependent answer
    plain = pm.main.calcMorphology(infos, str(tmp_path / "out2"), sky=sky, root=str(tmp_path), seed=5, smoothness=True)
    assert plain[0].star_flag is False and plain[0].maskedPixelFraction == -99.0 and np.ravel(plain[0].As)[0] != np.ravel(res[0].As)[0]
    # a header without the PSF cards: the catalogue is ignored with a message (main.py:389-391)
    bare = {k: v for k, v in cards.items() if not k.startswith(("PSF_", "PHT_"))}
    (tmp_path / "g.fits").write_bytes(fits_bytes(x[0], -32, extra=bare))
    r2 = pm.main.calcMorphology([("g.fits", ra0, dec0)], str(tmp_path / "out3"), catalogue=str(tmp_path / "cat.csv"), sky=sky[:1], root=str(tmp_path),
                                seed=5, smoothness=True)
    assert r2[0].star_flag is True and np.isclose(np.ravel(r2[0].As)[0], np.ravel(plain[0].As)[0], rtol=1e-12) and r2[0].maskedPixelFraction == 0.0


def test_segmap_starmask_and_saved_maps(tmp_path):
    """-spm / -sci write segmap_<name> / clean_<name> (float64 FITS under the frame's header, main.py:447-461); -segmap reads the
    pixelmaps back (main.py:429-440: no pixelmap, no edge flag) and gives the same statistics; -starmask reads starmask_<name>
    (main.py:397-417) and matches the star-mask composition."""
    import torch

    import pawlikmorph_lsst_b200 as pm
    import synth
    A = pm._abi
    n = 128
    x = synth.make_batch(3, n, 2718).numpy()
    names = []
    for k in range(3):
        (tmp_path / f"m{k}.fits").write_bytes(fits_bytes(x[k], -32))
        names.append(f"m{k}.fits")
    infos = [(nm, 0.0, 0.0) for nm in names]
    sky = [(synth.SKY, synth.SKY_ERR)] * 3
    out = tmp_path / "out"
    base = pm.main.calcMorphology(infos, str(out), sky=sky, root=str(tmp_path), seed=9, smoothness=True, saveSegMap=True, saveCleanImage=True)
    dev = torch.device("cuda:0")
    img = torch.from_numpy(x).to(dev)
    sky_d = torch.full((3,), synth.SKY, dtype=torch.float64, device=dev)
    err_d = torch.full((3,), synth.SKY_ERR, dtype=torch.float64, device=dev)
    seg = pm.batch.analyse_batch(img, sky_d, err_d, 3, 15 | A.PM_STOP_AFTER_PIXELMAP, want_segmap=True).segmap.cpu().numpy()
    cleaned = pm.imageutils.maskstars_batch(img.to(torch.float64) - sky_d[:, None, None], seed=9)["clean"].cpu().numpy()
    for k, r in enumerate(base):
        assert r.pixelMapFile == str(out / f"segmap_m{k}.fits") and r.cleanImage == str(out / f"clean_m{k}.fits")
        for path, want in ((r.pixelMapFile, seg[k].astype(np.float64)), (r.cleanImage, cleaned[k])):
            pay, hdr = pm.image.payload_of(open(path, "rb").read())
            assert hdr["BITPIX"] == -64 and hdr["NAXIS1"] == n and hdr["CTYPE1"] == "RA---TAN"
            assert np.array_equal(np.frombuffer(pay, dtype=">f8").reshape(n, n), want)
    again = pm.main.calcMorphology(infos, str(out), sky=sky, root=str(tmp_path), seed=9, smoothness=True, segmap=True)
    for a, b in zip(again, base):
        assert a.objectEdge is False
        for f in ("A", "As", "As90", "r20", "r80", "C", "S", "gini", "m20", "rmax", "apix"):
            assert np.allclose(np.ravel(getattr(a, f)).astype(float), np.ravel(getattr(b, f)).astype(float), rtol=1e-12), f
    # star masks from files
    rng = np.random.default_rng(4)
    star = np.ones((3, n, n))
    for k in range(3):
        r0, c0 = rng.integers(10, n - 30, 2)
        star[k, r0:r0 + 12, c0:c0 + 9] = 0
        star[k, n // 2 - 3:n // 2 + 4, n // 2 - 3:n // 2 + 4] = 1
        (out / f"starmask_m{k}.fits").write_bytes(pm.image.fits_image_bytes(star[k]))
    masked = pm.main.calcMorphology(infos, str(out), sky=sky, root=str(tmp_path), seed=9, smoothness=True, starMask=True)
    want = pm.batch.analyse_batch_starmask(img, sky_d, err_d, torch.from_numpy(star).to(dev), 3, 15,
                                           cleaned=torch.from_numpy(cleaned).to(dev)).rows.cpu().numpy()
    col = A.ROW_FIELDS.index
    for k, r in enumerate(masked):
        assert r.status == 0 and r.star_flag is False and r.maskedPixelFraction == -99.0
        assert np.allclose([r.A[0], np.ravel(r.As)[0], r.r20, r.gini], [want[k, col("A")], want[k, col("As")], want[k, col("r20")], want[k, col("gini")]], rtol=1e-12)
    with pytest.raises(ValueError):
        pm.main.calcMorphology(infos, str(out), sky=sky, root=str(tmp_path), segmap=True, starMask=True)
    with pytest.raises(OSError):
        pm.main.calcMorphology(infos, str(tmp_path / "empty"), sky=sky, root=str(tmp_path), segmap=True)
