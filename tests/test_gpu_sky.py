"""Sky-region statistics on the GPU against the numpy restatement (oracle/sky_oracle.py, parity unpinned)."""
import numpy as np
import pytest

from tests import sky_cases as S

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,dtype", [(64, np.float32), (49, np.float64), (140, np.float32)])
def test_sky_stats_on_the_gpu(n, dtype):
    import torch

    import pawlikmorph_lsst_b200 as pm
    x, fw, th = S.cases(n=n, dtype=dtype)
    rows = pm.skyBackground.sky_stats_batch(torch.from_numpy(x).cuda(), fw, th).cpu().numpy()
    assert S.check(rows, x, fw, th)
    assert tuple(pm._abi.SKY_FIELDS) == S.FIELDS


def test_sky_stats_full_size_batch_and_single_image_api():
    import torch

    import pawlikmorph_lsst_b200 as pm
    import synth
    from oracle import sky_oracle as SO
    B, n = 600, 256                                             # more cutouts than CTAs in flight
    x = synth.make_batch(B, n, 41, device="cuda:0")
    rng = np.random.default_rng(8)
    fw = rng.uniform(2.0, 12.0, (B, 2))
    th = rng.uniform(-3.0, 3.0, B)
    rows = pm.skyBackground.sky_stats_batch(x, fw, th).cpu().numpy()
    again = pm.skyBackground.sky_stats_batch(x, fw, th).cpu().numpy()
    assert np.array_equal(rows, again)                          # reproducible
    xs = x.cpu().numpy()
    pick = [0, 1, 299, 598, 599]
    assert S.check(rows[pick], xs[pick], fw[pick], th[pick])
    assert np.all(rows[:, 11] == 0) and np.all(np.abs(rows[:, 9] - synth.SKY) < 10.0)    # near the synthetic sky level (galaxy wings reach outside small ellipses)
    sky, err = pm.skyBackground.skyFromRegion(xs[3].astype(np.float64), list(fw[3]), float(th[3]))
    w = SO.calc_sky(xs[3], fw[3], th[3])
    assert abs(sky - w["sky"]) <= 1e-9 * abs(w["sky"]) and abs(err - w["sky_err"]) <= 1e-9 * w["sky_err"]
    with pytest.raises(pm.skyBackground._SkyError):
        pm.skyBackground.skyFromRegion(xs[3][:40, :40].astype(np.float64), [30.0, 30.0], 0.0)


@pytest.mark.parametrize("cluster", ["0", "1", "2", "4", "8"])
def test_every_cluster_size_gives_the_same_statistics(cluster, monkeypatch):
    """One cutout per thread-block cluster, its rows spread over the CTAs' shared memory: counts, medians and extremes are
    exact for every split; means and deviations are sums in a different order."""
    import torch

    import pawlikmorph_lsst_b200 as pm
    import synth
    B, n = 300, 140
    x = synth.make_batch(B, n, 77, device="cuda:0")
    x[5, 3, 7] = float("nan")
    x[6] = 3.25                                                  # constant cutout: every key equal
    rng = np.random.default_rng(3)
    fw = rng.uniform(2.0, 12.0, (B, 2))
    th = rng.uniform(-3.0, 3.0, B)
    want_sky = pm.skyBackground.sky_stats_batch(x, fw, th).cpu().numpy()
    want_clip = pm.skyBackground.clipped_stats_batch(x, 3.0, None).cpu().numpy()
    xs = x.cpu().numpy()
    assert S.check(want_sky[:4], xs[:4], fw[:4], th[:4])
    monkeypatch.setenv("PM_STAT_CLUSTER", cluster)
    L = pm._lib.lib()
    import ctypes as C
    cl, res, smem = C.c_int(), C.c_int(), C.c_int()
    assert L.pm_stats_plan(n, pm._abi.PM_F32, 1, C.byref(cl), C.byref(res), C.byref(smem)) == 0
    assert (cl.value, res.value) == ((1, 0) if cluster == "0" else (int(cluster), 1))
    got_sky = pm.skyBackground.sky_stats_batch(x, fw, th).cpu().numpy()
    got_clip = pm.skyBackground.clipped_stats_batch(x, 3.0, None).cpu().numpy()
    exact_sky, exact_clip = [0, 2, 4, 5, 7, 11], [0, 2, 4, 5, 7, 9]
    assert np.array_equal(got_sky[:, exact_sky], want_sky[:, exact_sky], equal_nan=True)
    assert np.array_equal(got_clip[:, exact_clip], want_clip[:, exact_clip], equal_nan=True)
    np.testing.assert_allclose(got_sky, want_sky, rtol=1e-11, equal_nan=True)
    np.testing.assert_allclose(got_clip, want_clip, rtol=1e-11, equal_nan=True)


@pytest.mark.parametrize("n,dtype", [(256, np.float32), (512, np.float32), (256, np.float64), (640, np.float64)])
def test_default_plans_against_numpy(n, dtype):
    """The default plan at the BASELINE sizes and beyond."""
    import ctypes as C

    import torch

    import pawlikmorph_lsst_b200 as pm
    rng = np.random.default_rng(n)
    B = 40
    x = (100.0 + 4.0 * rng.standard_normal((B, n, n))).astype(dtype)
    x[:, n // 2 - 6:n // 2 + 6, n // 2 - 6:n // 2 + 6] += 500.0
    cl, res, smem = C.c_int(), C.c_int(), C.c_int()
    L = pm._lib.lib()
    assert L.pm_stats_plan(n, pm._abi.PM_F32 if dtype == np.float32 else pm._abi.PM_F64, 0, C.byref(cl), C.byref(res), C.byref(smem)) == 0
    assert (cl.value, res.value) == (1, 0)                     # default: one CTA per cutout, passes through L2
    got = pm.skyBackground.clipped_stats_batch(torch.from_numpy(x).cuda(), 3.0, 5).cpu().numpy()
    from oracle import prep_oracle as PO
    for b in (0, 17, B - 1):
        v = x[b].astype(np.float64)
        assert got[b, 0] == v.size and got[b, 2] == np.median(v) and got[b, 9] == v.max()
        np.testing.assert_allclose(got[b, [1, 3]], [v.mean(), v.std()], rtol=1e-12)
        np.testing.assert_allclose(got[b, 6:9], PO.clipped(x[b], 3.0, 5), rtol=1e-12)
