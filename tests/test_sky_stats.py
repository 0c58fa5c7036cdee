"""Sky-region statistics (pm_sky_stats_batch) on the CPU build of the kernel sources against the numpy restatement."""
import numpy as np
import pytest

from oracle import sky_oracle as SO
from tests import emul_harness as H
from tests import sky_cases as S


@pytest.mark.parametrize("n,dtype", [(64, np.float32), (49, np.float64)])
def test_sky_stats_against_the_oracle(n, dtype):
    x, fw, th = S.cases(n=n, dtype=dtype)
    rows = H.sky_stats(x, fw, th)
    assert S.check(rows, x, fw, th)
    assert rows[1, 4] == 0 and rows[0, 4] > 0          # both branches are exercised
    assert rows[4, 11] == 1.0 or rows[4, 0] >= 100     # the large ellipse leaves few sky pixels


def test_oracle_pieces():
    m = SO.object_region(9, [1.0, 1.0], 0.0)           # a = b = 2 about (5, 5): the 9 pixels with dx^2 + dy^2 < 4
    assert m.sum() == 9 and m[5, 5] and m[4, 4] and not m[5, 7] and not m[3, 5]
    v = np.concatenate([np.zeros(99), [1000.0]])
    mean, med, std, it, cnt = SO.sigma_clipped_stats(v)
    assert (mean, med, std, cnt) == (0.0, 0.0, 0.0, 99) and it == 2


@pytest.mark.parametrize("mode", ["0", "1"])
def test_resident_and_global_memory_forms_agree(mode, monkeypatch):
    """Default / PM_STAT_CLUSTER=0: the cutout stays in global memory (every pass re-reads it); 1 holds it in shared
    memory: the same statistics bit for bit."""
    x, fw, th = S.cases(n=64, dtype=np.float32)
    base = H.sky_stats(x, fw, th)
    monkeypatch.setenv("PM_STAT_CLUSTER", mode)
    rows = H.sky_stats(x, fw, th)
    assert np.array_equal(rows, base, equal_nan=True)
    c = H.clipped_stats(x, 3.0, 5)
    monkeypatch.delenv("PM_STAT_CLUSTER")
    assert np.array_equal(c, H.clipped_stats(x, 3.0, 5), equal_nan=True)


def test_even_count_median_across_digit_boundaries():
    """The upper middle element of an even count comes from the last digit's histogram, from a later thread's run of
    bins, or (no larger key under the same prefix) from one extra pass."""
    n = 16
    base = np.full((n, n), 5.0, dtype=np.float32)
    cases = []
    a = base.copy(); a.flat[:128] = 1.0; a.flat[128:] = 9.0                      # middle pair (1, 9): different exponent
    cases.append(a)
    b = base.copy(); b.flat[:128] = 2.0; b.flat[128:] = np.nextafter(np.float32(2.0), np.float32(3.0))   # adjacent keys
    cases.append(b)
    c = base.copy(); c.flat[:128] = 3.0; c.flat[128:] = 3.0 + 2.0 ** -9           # same top digits, far apart in the last
    cases.append(c)
    d = base.copy(); d.flat[:127] = -4.0; d.flat[127] = -0.0; d.flat[128] = 0.0; d.flat[129:] = 7.0     # signed zeros in the middle
    cases.append(d)
    e = base.copy(); e.flat[:128] = -1.5; e.flat[128:] = 1.5                      # keys differ in the top bit
    cases.append(e)
    x = np.stack(cases)
    got = H.clipped_stats(x, 3.0, 1)
    for k in range(len(cases)):
        v = x[k].astype(np.float64)
        assert got[k, 2] == np.median(v), k
        assert got[k, 0] == v.size and got[k, 9] == v.max()
    x64 = x.astype(np.float64)
    x64[1].flat[128:] = np.nextafter(2.0, 3.0)
    got = H.clipped_stats(x64, 3.0, 1)
    for k in range(len(cases)):
        assert got[k, 2] == np.median(x64[k]), k
