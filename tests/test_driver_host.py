"""Host side of the file driver (no GPU): image list reader, Result CSV row, flag logic, header parser, sky stand-in."""
import csv
import io

import numpy as np

import pawlikmorph_lsst_b200 as pm
from tests import parity_checks as P
from tests.fits_files import fits_bytes


def test_getFiles_reads_the_reference_image_list_format(tmp_path):
    f = tmp_path / "images.csv"
    f.write_text("filename, ra, dec\nsample/a_rband.fits, 149.00575, 0.3580208\nsample/b.fits, 150.5, -1.25\n\n")   # images.csv:1-3
    assert list(pm.helpers.getFiles(str(f))) == [("sample/a_rband.fits", 149.00575, 0.3580208), ("sample/b.fits", 150.5, -1.25)]


def test_result_row_has_the_reference_columns():
    buf = io.StringIO()
    w = csv.writer(buf)
    w.writerow(pm.result.CSV_HEADER)
    row = dict(zip(pm._abi.ROW_FIELDS, [70.0, 71.0, 12.5, 3.0, 9.0, 2.385, 0.1, 0.02, 0.3, 0.2, 0.25, 0.5, -1.7, 1.0, 0.0, 40.0]))
    pm.result.Result.from_row(row, file="x.fits", sky=100.0, sky_err=5.0).write(w)
    pm.result.Result.from_row(dict(row, status=1.0), file="faint.fits", sky=1.0, sky_err=2.0).write(w)
    rows = list(csv.reader(io.StringIO(buf.getvalue())))
    assert len(rows[0]) == len(rows[1]) == len(rows[2]) == 28                       # result.py:81-92
    d = dict(zip(rows[0], rows[1]))
    assert d["file"] == "x.fits" and d["apix"] == "[70, 71]" and d["r_max"] == "12.5" and d["A"] == "0.1" and d["Abgr"] == "0.02"
    assert d["As"] == "0.2" and d["As90"] == "0.25" and d["Gini index"] == "0.5" and d["M20"] == "-1.7"
    assert d["fwhms"] == "[-99.0, -99.0]" and d["sersic_n"] == "-99.0" and d["Object on image edge?"] == "True"
    f = dict(zip(rows[0], rows[2]))                                                   # main.py:420-424: default Result
    assert f["apix"] == "(-99.0, -99.0)" and f["r_max"] == "-99" and f["A"] == "-99.0" and f["sky"] == "1.0" and f["sky_err"] == "2.0"


def test_flag_logic_of_the_driver():
    A, AS, CAS, S = pm._abi.PM_DO_A, pm._abi.PM_DO_AS, pm._abi.PM_DO_CAS, pm._abi.PM_DO_S
    f = pm.main.path_flags
    assert f(asymmetry=True, allAsymmetry=False, CAS=False) == A
    assert f(shapeAsymmetry=True, allAsymmetry=False, CAS=False) == AS
    assert f(allAsymmetry=True, CAS=False) == A | AS
    assert f(allAsymmetry=False, CAS=True) == A | CAS                                 # main.py:469: -cas alone computes A too
    assert f() == A | AS | CAS and f(smoothness=True) == A | AS | CAS | S


def test_header_parser_and_border_sky_on_a_golden_frame():
    g = P.golden("sdss_s")
    img = g.image(3)
    for bitpix in (16, -32):
        payload, hdr = pm.image.payload_of(fits_bytes(img, bitpix))
        assert (hdr["BITPIX"], hdr["NAXIS1"], hdr["NAXIS2"], hdr["CTYPE1"]) == (bitpix, 141, 141, "RA---TAN")
        assert np.array_equal(pm.main._host_pixels(payload, hdr), img.astype(np.float64))
    assert pm.main.border_sky(img) == g.sky(3)                                        # the estimator the goldens were made with


def test_driver_dataflow_on_the_emulator_build(tmp_path, monkeypatch):
    """calcMorphology with the batched GPU call replaced by the CPU build of the same kernel sources (tests/emul):
    grouping by frame format, order of the rows, sky plumbing and the CSV, against the golden rows of four frames."""
    import csv as _csv

    from tests import emul_harness as H
    from tests.golden_util import compare_row
    g = P.golden("sdss_s")
    pick = [0, 5, 9, 12]

    def fake(payload, bitpix, n, sky, sky_err, bzero=0.0, bscale=1.0, filter_size=3, flags=15, **kw):
        x = np.stack([H.fits_decode(p.tobytes(), bitpix, n * n, bzero, bscale, np.float32).reshape(n, n) for p in payload])
        return H.analyse(x, sky, sky_err, filter_size, flags, want=("segmap",))["rows"]

    import pawlikmorph_lsst_b200.batch as batch
    monkeypatch.setattr(batch, "analyse_host_payload", fake)
    infos = []
    for k, i in enumerate(pick):
        fn = f"f{k}.fits"
        (tmp_path / fn).write_bytes(fits_bytes(g.image(i), -32 if k == 1 else 16))
        infos.append((fn, 1.0, 2.0))
    res = pm.main.calcMorphology(infos, tmp_path / "out", smoothness=True, root=tmp_path, sky="border", clean=False)
    table = list(_csv.reader(open(tmp_path / "out" / "parameters.csv")))
    assert table[0] == pm.result.CSV_HEADER and [r[0] for r in table[1:]] == [f for f, _, _ in infos]
    for k, i in enumerate(pick):
        r = res[k]
        assert (r.sky, r.sky_err) == g.sky(i)
        got = dict(apix_col=r.apix[0], apix_row=r.apix[1], rmax=r.rmax, r20=r.r20, r80=r.r80, C=r.C, A=r.A[0], Abgr=r.A[1], S=r.S,
                   As=r.As[0], As90=r.As90[0], gini=r.gini, m20=r.m20, object_edge=float(r.objectEdge), status=r.status)
        errs = compare_row(got, g.row(i), P.RTOL, f"sdss_s[{i}] ")
        assert not errs, "\n".join(errs)
    import pytest
    with pytest.raises(NotImplementedError):
        pm.main.calcMorphology(infos, tmp_path / "out", calculateSersic=True, root=tmp_path, sky="border", clean=False)


def _wcs_by_name():
    import json
    import os
    J = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sdss_wcs.json")))
    return {e["file"].split("/")[-1]: e for e in J}


def test_driver_cutout_on_the_emulator_build(tmp_path, monkeypatch):
    """-s 128 (imganalysis.py:36): the 128 x 128 window around the catalogue position, cut on the 'device' (CPU build of
    the window decode), analysed by the CPU build of the kernel, against the oracle run on the numpy slice."""
    from oracle import pm_oracle as O
    from tests import emul_harness as H
    from tests.golden_util import compare_row
    g = P.golden("sdss_l")
    wcs = _wcs_by_name()
    seen = {}

    def fake(payload, bitpix, n, sky, sky_err, bzero=0.0, bscale=1.0, filter_size=3, flags=15, frame=None, origins=None, **kw):
        assert frame is not None and n == 128
        x = np.concatenate([H.fits_decode_window(p.tobytes(), bitpix, 1, frame[0], frame[1], origins[k], n, bzero, bscale, np.float32)
                            for k, p in enumerate(payload)])
        seen[frame] = x
        return H.analyse(x, sky, sky_err, filter_size, flags, want=("segmap",))["rows"]

    import pawlikmorph_lsst_b200.batch as batch
    monkeypatch.setattr(batch, "analyse_host_payload", fake)
    infos, pick = [], [29, 31]                                           # the 177- and 180-pixel frames
    for i in pick:
        name = str(g.z["names"][i])
        (tmp_path / name).write_bytes(fits_bytes(g.image(i), 16, extra=wcs[name]["header"]))
        infos.append((name, wcs[name]["ra"], wcs[name]["dec"]))
    res = pm.main.calcMorphology(infos, tmp_path / "out", npix=128, smoothness=True, root=tmp_path, sky="border", clean=False)
    for k, i in enumerate(pick):
        h = wcs[infos[k][0]]["header"]
        r0, c0 = pm.image.cutout_origin(*pm.image.sky_to_pixel(h, infos[k][1], infos[k][2]), 128, g.image(i).shape)
        win = g.image(i)[r0:r0 + 128, c0:c0 + 128].astype(np.float64)
        assert np.array_equal(seen[g.image(i).shape][0], win)
        assert (res[k].sky, res[k].sky_err) == pm.main.border_sky(win)
        want, _, _ = O.analyse(win, res[k].sky, res[k].sky_err, 3)
        r = res[k]
        got = dict(apix_col=r.apix[0], apix_row=r.apix[1], rmax=r.rmax, r20=r.r20, r80=r.r80, C=r.C, A=r.A[0], Abgr=r.A[1], S=r.S,
                   As=r.As[0], As90=r.As90[0], gini=r.gini, m20=r.m20, object_edge=float(r.objectEdge), status=r.status)
        errs = compare_row(got, want, P.RTOL, f"sdss_l[{i}] window ")
        assert not errs, "\n".join(errs)
    # dec == 0 is falsy upstream (Image.py:135): the whole frame is analysed
    seen.clear()
    monkeypatch.setattr(batch, "analyse_host_payload", lambda payload, bitpix, n, *a, frame=None, **kw: (
        seen.setdefault("n", n), np.full((len(payload), 16), -99.0))[1])
    pm.main.calcMorphology([(infos[0][0], infos[0][1], 0.0)], tmp_path / "out", npix=128, root=tmp_path, sky="border", clean=False)
    assert seen["n"] == 177


def test_driver_dec_first_header_and_bad_frames(tmp_path, monkeypatch):
    """A DEC-first copy of a bundled frame is cut at the same origin as the RA-first one (sky_to_pixel already returns
    (column, row); no second swap), and a frame that cannot be read / whose window leaves it gets the default Result
    while the others are analysed (main.py:331-339).  The file column holds the base name (main.py:341)."""
    g = P.golden("sdss_l")
    wcs = _wcs_by_name()
    name = str(g.z["names"][29])
    h = wcs[name]["header"]
    hr = dict(h, CTYPE1=h["CTYPE2"], CTYPE2=h["CTYPE1"], CRVAL1=h["CRVAL2"], CRVAL2=h["CRVAL1"],
              PC1_1=h["PC2_1"], PC1_2=h["PC2_2"], PC2_1=h["PC1_1"], PC2_2=h["PC1_2"])
    hr = {k: v for k, v in hr.items() if not k.startswith("CD")}
    sub = tmp_path / "sub"
    sub.mkdir()
    (sub / "ra_first.fits").write_bytes(fits_bytes(g.image(29), 16, extra=h))
    (sub / "dec_first.fits").write_bytes(fits_bytes(g.image(29), 16, extra=hr))
    (sub / "edge.fits").write_bytes(fits_bytes(g.image(29), 16, extra=dict(h, CRPIX1=float(h["CRPIX1"]) + 80.0)))
    seen = []

    def fake(payload, bitpix, n, sky, sky_err, frame=None, origins=None, **kw):
        seen.append(np.asarray(origins).copy())
        rows = np.full((len(payload), 16), -99.0)
        rows[:, 14] = 0.0
        rows[:, 2] = 7.0
        return rows

    import pawlikmorph_lsst_b200.batch as batch
    monkeypatch.setattr(batch, "analyse_host_payload", fake)
    ra, dec = wcs[name]["ra"], wcs[name]["dec"]
    infos = [("sub/ra_first.fits", ra, dec), ("sub/missing.fits", ra, dec), ("sub/dec_first.fits", ra, dec), ("sub/edge.fits", ra, dec)]
    res = pm.main.calcMorphology(infos, tmp_path / "out", npix=128, root=tmp_path, sky="border", clean=False)
    assert len(seen) == 1 and seen[0].shape == (2, 2) and np.array_equal(seen[0][0], seen[0][1])
    assert [r.file for r in res] == ["ra_first.fits", "missing.fits", "dec_first.fits", "edge.fits"]
    assert [r.status for r in res] == [0, 1, 0, 1] and res[1].rmax == -99 and res[3].rmax == -99 and res[0].rmax == 7.0
    table = list(csv.reader(open(tmp_path / "out" / "parameters.csv")))
    assert len(table) == 5 and table[2][0] == "missing.fits" and table[2][2] == "-99"


def test_sky_to_pixel_refuses_distorted_headers():
    import pytest
    h = _wcs_by_name()[next(iter(_wcs_by_name()))]["header"]
    for bad in (dict(h, CTYPE1="RA---TAN-SIP", CTYPE2="DEC--TAN-SIP"), dict(h, A_ORDER=2), dict(h, PV1_1=0.1)):
        with pytest.raises(NotImplementedError):
            pm.image.sky_to_pixel(bad, 1.0, 2.0)
    bare = {k: v for k, v in h.items() if not k.startswith(("PC", "CD"))}
    with pytest.raises(NotImplementedError):
        pm.image.sky_to_pixel(dict(bare, CROTA2=12.0), 1.0, 2.0)
