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
