"""CPU suite: the product's kernel sources compiled against the CTA emulator (tests/emul), checked
against the golden vectors of the unmodified reference and against the oracle.  This exercises the
kernel LOGIC without a GPU; the real parity gate is tests/test_gpu_parity.py on a B200."""
import numpy as np
import pytest

from tests import parity_checks as P
from tests.runners import emul_runner as run

GOLD = ([("sdss_s", i) for i in range(0, 69, 4)] + [("sdss_l", i) for i in range(0, 69, 6)]
        + [("synth", i) for i in range(37) if i != 34])


@pytest.mark.parametrize("name,i", GOLD)
def test_golden(name, i):
    P.check_golden_case(run, name, i)


@pytest.mark.parametrize("n,fs,seed", [(48, 3, 1), (64, 5, 2), (96, 3, 3), (75, 7, 4), (141, 3, 5), (64, 1, 6)])
def test_oracle_synth(n, fs, seed):
    P.check_against_oracle(run, P.synth_batch(3, n, seed), 100.0, 5.0, fs)


def test_float64_input():
    x = P.synth_batch(2, 64, 21).astype(np.float64) + 1e-9 * np.random.default_rng(0).standard_normal((2, 64, 64))
    P.check_against_oracle(run, x, 100.0, 5.0, 3, dtype=np.float64)


def test_exact_filter_path_matches_fast_path():
    from tests.emul_harness import abi
    x = P.synth_batch(2, 64, 31)
    a = run(x, 100.0, 5.0, 3, 15)
    b = run(x, 100.0, 5.0, 3, 15 | abi.PM_FORCE_EXACT_FILTER)
    assert np.array_equal(a["segmap"], b["segmap"]) and np.array_equal(a["rows"], b["rows"])
    P.check_against_oracle(lambda *args, **kw: run(args[0], args[1], args[2], args[3], args[4] | abi.PM_FORCE_EXACT_FILTER),
                           x, 100.0, 5.0, 3)


def test_threshold_ties_take_the_exact_path():
    """Integer-valued image and a threshold that some shrunk pixel hits EXACTLY: `>` is strict
    (pixmap.py:203) and the decision must be scipy's, bit for bit."""
    from oracle import pm_oracle as O
    rng = np.random.default_rng(5)
    x = P.synth_batch(1, 66, 41)[0]
    x = np.rint(x).astype(np.float32)
    z = O.mean_filter_shrink(x.astype(np.float64), 3)
    cands = np.sort(z[z > 104.0].ravel())
    thr = float(cands[len(cands) // 3])          # an exact Z value
    res = run(x[None], 0.0, 0.0, 3, 15, threshold_in=[thr])
    seg = O.pixelmap(x.astype(np.float64), thr, 3)
    assert np.array_equal(res["segmap"][0], seg.astype(np.uint8))


def test_flag_subsets_and_faint_centre():
    x = P.synth_batch(2, 64, 51)
    for flags in (1, 2, 3, 4, 5, 7, 12):
        P.check_against_oracle(run, x, 100.0, 5.0, 3, flags, check_sorted=False)
    # faint centre -> status 1, all sentinels, empty map (pixmap.py:178-179, main.py:420-424)
    res = run(x, 100.0, 1e6, 3, 15)
    assert np.all(res["rows"][:, 14] == 1) and np.all(res["rows"][:, :13] == -99) and not res["segmap"].any()


@pytest.mark.parametrize("name,i", [("sdss_s", i) for i in range(1, 69, 9)] + [("sdss_l", i) for i in range(2, 69, 11)]
                         + [("synth", i) for i in (0, 11, 19, 26, 27, 29, 31)])
def test_golden_with_pruned_minapix(name, i):
    """Without the per-candidate output minapix drops candidates whose partial numerator already exceeds the best
    complete asymmetry (exact pruning): the chosen centre, and with it every statistic, must not change."""
    from tests.runners import emul_runner_pruned
    P.check_golden_case(emul_runner_pruned, name, i)


def test_pruned_and_unpruned_minapix_agree():
    from tests.emul_harness import abi
    from tests.runners import emul_runner_pruned
    x = P.synth_batch(6, 128, 91)
    a = emul_runner_pruned(x, 100.0, 5.0, 3, 15)
    b = emul_runner_pruned(x, 100.0, 5.0, 3, 15 | abi.PM_NO_PRUNE)
    assert np.array_equal(a["rows"], b["rows"], equal_nan=True)


@pytest.mark.parametrize("n,fs", [(8, 1), (9, 3), (12, 3), (15, 5), (16, 15), (17, 1), (33, 11), (40, 13)])
def test_tiny_cutouts_and_every_filter_size(n, fs):
    """Smallest sizes the ABI accepts and the whole range of filter sizes (imganalysis.py:45-47)."""
    rng = np.random.default_rng(7 + n)
    yy, xx = np.mgrid[0:n, 0:n]
    x = np.stack([(100 + rng.normal(0, 5, (n, n)) + 400 * np.exp(-((xx - n // 2) ** 2 + (yy - n // 2) ** 2) / (2 * (n / 8) ** 2)))
                  for _ in range(2)]).astype(np.float32)
    P.check_against_oracle(run, x, 100.0, 5.0, fs)


def test_large_cutout_with_large_filter_spills_staging():
    """n = 640 with filterSize 15: the pixelmap staging rows no longer fit shared memory twice, so the TMA
    double buffer must give way to the cooperative-load path (the emulator aborts on a TMA copy whose
    destination is not shared memory)."""
    rng = np.random.default_rng(123)
    n = 640
    yy, xx = np.mgrid[0:n, 0:n]
    img = (100 + rng.normal(0, 5, (n, n)) + 900 * np.exp(-(((xx - 321) / 7.0) ** 2 + ((yy - 318) / 5.0) ** 2) / 2)).astype(np.float32)
    P.check_against_oracle(run, img[None], 100.0, 5.0, 15)


@pytest.mark.parametrize("n", [15, 16, 31, 64, 141])
def test_aperture_sweep_matches_oracle(n):
    """aperpixmap (apertures.py:42-128) over integer-form (sqrt of an integer: what calcRmax yields) and
    float-form radii, including none / all of the frame: the windowed edge search and its fallback."""
    import math
    from oracle import pm_oracle as O
    from tests import emul_harness as H
    rng = np.random.default_rng(n)
    rads = [0.0, 0.3, 0.75, 1.0, math.sqrt(2.0), 2.5, float(n), 3.0 * n]
    rads += [math.sqrt(float(k)) for k in rng.integers(1, (n // 2 + 2) ** 2, size=24)]
    rads += [float(x) for x in rng.uniform(0.5, 0.75 * n, size=16)]
    got = H.aperpixmap(n, np.asarray(rads, dtype=np.float64))
    for k, r in enumerate(rads):
        assert np.array_equal(got[k], O.aperpixmap(n, r, 9, 0.1)), f"n={n} rad={r!r}"


def test_small_radius_growth_curves():
    P.check_small_radius_growth_curves(run)
