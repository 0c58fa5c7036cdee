"""fitSersic on the GPU against the oracle (scipy's brentq / leastsq as the reference drives them; exact overlap restated)."""
import numpy as np
import pytest

from oracle import sersic_oracle as S
from tests import sersic_cases as C

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_ellipse_sums_on_the_device(dtype):
    import torch

    import pawlikmorph_lsst_b200 as pm
    rng = np.random.default_rng(1)
    n = 128
    img = rng.normal(10, 1, (4, n, n)).astype(dtype)
    ell = [[rng.uniform(40, 88), rng.uniform(40, 88), rng.uniform(2, 30), rng.uniform(1, 16), rng.uniform(-3, 3)] for _ in range(12)]
    ell += [[61.3, 60.2, 0.001, 0.5, 0.4], [2.0, 120.0, 9.0, 4.0, 1.0], [-30.0, -30.0, 3.0, 2.0, 0.0], [64.0, 64.0, 200.0, 190.0, 0.2]]
    idx = [k % 4 for k in range(len(ell))]
    got = pm.sersic.ellipse_sum_batch(torch.from_numpy(img).cuda(), ell, idx=idx).cpu().numpy()
    want = np.array([S.ellipse_sum(img[i], *e) for i, e in zip(idx, ell)])
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)
    with pytest.raises(IndexError):
        pm.sersic.ellipse_sum_batch(torch.from_numpy(img).cuda(), ell[:1], idx=[7])


def test_fit_sersic_batch_against_the_oracle():
    import torch

    import pawlikmorph_lsst_b200 as pm
    imgs, truths, cen, fw, th = C.cases(n=128, count=6, seed=11)
    fw[4] = [-99.0, -99.0]                                           # no sky estimate: no start values
    r = pm.sersic.fitSersic_batch(torch.from_numpy(imgs).cuda(), cen, fw, th)
    assert list(r["status"]) == [0, 0, 0, 0, 4, 0] and np.isnan(r["params"][4]).all()
    for k in (0, 1, 2, 3, 5):
        np.testing.assert_allclose(r["start"][k], S.sersic_start(imgs[k], cen[k], fw[k], th[k]), rtol=1e-10)
        want, info = S.fit_sersic(imgs[k], cen[k], fw[k], th[k], return_info=True)
        assert C.close(r["params"][k], want), (k, r["params"][k], want)
        assert r["cost"][k] <= info["cost"] * (1 + 1e-9)
    again = pm.sersic.fitSersic_batch(torch.from_numpy(imgs).cuda(), cen, fw, th)
    assert np.array_equal(again["params"], r["params"], equal_nan=True)            # reproducible
    # one image, the reference's signature and return object; a star mask multiplies the image (sersic.py:88-91)
    p = pm.sersic.fitSersic(imgs[1], cen[1], fw[1], th[1])
    assert p.r_eff.value == r["params"][1, 1] and p.n.value == r["params"][1, 2]
    mask = np.ones_like(imgs[1])
    mask[:20] = 0.0
    pmask = pm.sersic.fitSersic(imgs[1], cen[1], fw[1], th[1], starMask=mask)
    assert C.close(pmask.parameters, S.fit_sersic(imgs[1], cen[1], fw[1], th[1], star_mask=mask), rtol=5e-4)
    with pytest.raises(ValueError):
        pm.sersic.fitSersic(imgs[1], cen[1], [-99.0, -99.0], th[1])


def test_driver_fills_the_sersic_columns(tmp_path):
    """calcMorphology(calculateSersic=True): main.py:480-488 after the default flow (skybgr widths, cleaned image, apix)."""
    import pawlikmorph_lsst_b200 as pm
    from tests.fits_files import fits_bytes
    rng = np.random.default_rng(5)
    n = 128
    y, x = np.mgrid[0:n, 0:n]
    names = []
    for k in range(3):
        t = [rng.uniform(150, 400), rng.uniform(5, 9), rng.uniform(1.0, 2.5), 64 + rng.uniform(-1, 1), 64 + rng.uniform(-1, 1), rng.uniform(0.1, 0.4),
             rng.uniform(0, 3)]
        img = (S.sersic2d(x, y, *t) + 100.0 + rng.normal(0, 2.0, (n, n))).astype(np.float32)
        name = f"g{k}.fits"
        (tmp_path / name).write_bytes(fits_bytes(img, -32))
        names.append(name)
    res = pm.main.calcMorphology([(nm, 0.0, 0.0) for nm in names], str(tmp_path / "out"), calculateSersic=True, root=str(tmp_path), seed=3)
    for r in res:
        assert r.status == 0 and r.sersic_r_eff > 0 and 0.3 < r.sersic_n < 6 and abs(r.sersic_x_0 - 64) < 3 and abs(r.sersic_y_0 - 64) < 3
    plain = pm.main.calcMorphology([(nm, 0.0, 0.0) for nm in names], str(tmp_path / "out2"), root=str(tmp_path), seed=3)
    assert all(p.sersic_n == -99.0 and p.A == r.A for p, r in zip(plain, res))
    with pytest.raises(NotImplementedError):
        pm.main.calcMorphology([(names[0], 0.0, 0.0)], str(tmp_path / "out3"), calculateSersic=True, sky="border", root=str(tmp_path))
