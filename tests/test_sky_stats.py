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
