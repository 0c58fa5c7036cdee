"""CPU suite: the execution paths of pm_analyse_batch give the same rows — the three kernels split by resource class
(PM_PATH_SPLIT: front / minapix / tail with one CTA per cutout, or with one warp per cutout: PM_TAIL_WARP), the
monolith (the default), and the monolith pass that takes the cutouts whose lists do not fit a hand-over slot."""
import os

import numpy as np
import pytest

from tests import parity_checks as P
from tests.emul_harness import abi
from tests.runners import emul_runner, emul_runner_pruned


def _with_flags(run, extra):
    return lambda x, sky, err, fs, flags, **kw: run(x, sky, err, fs, flags | extra, **kw)


@pytest.mark.parametrize("extra", [abi.PM_PATH_SPLIT | abi.PM_TAIL_WARP, abi.PM_PATH_SPLIT, 0], ids=["split-warp", "split-cta", "monolith"])
@pytest.mark.parametrize("name,i", [("sdss_s", 7), ("sdss_l", 20), ("synth", 5), ("synth", 27)])
def test_golden_on_every_path(extra, name, i):
    P.check_golden_case(_with_flags(emul_runner, extra), name, i)
    P.check_golden_case(_with_flags(emul_runner_pruned, extra), name, i)


def test_paths_agree_on_a_batch():
    x = P.synth_batch(7, 96, 77)
    ref = emul_runner_pruned(x, 100.0, 5.0, 3, 15)
    for extra in (abi.PM_PATH_SPLIT, abi.PM_PATH_SPLIT | abi.PM_TAIL_WARP):
        got = emul_runner_pruned(x, 100.0, 5.0, 3, 15 | extra)
        assert np.array_equal(got["segmap"], ref["segmap"])
        np.testing.assert_allclose(got["rows"], ref["rows"], rtol=1e-11, atol=0, equal_nan=True)
        assert np.array_equal(got["rows"][:, [0, 1, 2, 13, 14, 15]], ref["rows"][:, [0, 1, 2, 13, 14, 15]])


def test_small_slots_take_the_monolith_pass(monkeypatch):
    """Hand-over slots too small for most cutouts: those go through the monolith pass, the others through the split
    kernels, and the table is the same."""
    x = P.synth_batch(6, 96, 78)
    ref = emul_runner_pruned(x, 100.0, 5.0, 3, 15)
    ng = ref["segmap"].reshape(6, -1).sum(axis=1)
    monkeypatch.setenv("PM_TEST_CAND_CAP", "8")
    monkeypatch.setenv("PM_TEST_LIST_CAP", str(int(np.sort(ng)[3]) + 8))      # about half of the cutouts fit
    got = emul_runner_pruned(x, 100.0, 5.0, 3, 15 | abi.PM_PATH_SPLIT)
    assert np.array_equal(got["segmap"], ref["segmap"])
    np.testing.assert_allclose(got["rows"], ref["rows"], rtol=1e-11, atol=0, equal_nan=True)
    P.check_against_oracle(_with_flags(emul_runner, abi.PM_PATH_SPLIT), x[:2], 100.0, 5.0, 3)


def test_stage_flags_and_overrides_on_the_split_path():
    """PM_STOP_AFTER_PIXELMAP / PM_STOP_AFTER_RMAX end in the front kernel, a supplied centre skips the minapix kernel:
    rows equal the monolith's."""
    x = P.synth_batch(3, 64, 79)
    for flags, kw in ((15 | abi.PM_STOP_AFTER_PIXELMAP, {}), (15 | abi.PM_STOP_AFTER_RMAX, {}),
                      (15, dict(centroid_in=np.tile([31.0, 33.0], (3, 1)))), (15, dict(radius_in=np.array([9.0, 11.5, 14.0])))):
        a = emul_runner_pruned(x, 100.0, 5.0, 3, flags | abi.PM_PATH_SPLIT, **kw)
        b = emul_runner_pruned(x, 100.0, 5.0, 3, flags, **kw)
        np.testing.assert_allclose(a["rows"], b["rows"], rtol=1e-11, atol=0, equal_nan=True)


def test_bucket_sort_and_radix_sort_give_the_same_order():
    """The bucket sort (default) and the LSD radix sort (its fallback for crowded buckets, PM_RADIX_SORT) produce the
    same canonical order — float noise, integer-valued frames with thousands of ties, float64 keys."""
    g = P.golden("sdss_s")
    cases = [(P.synth_batch(3, 96, 81), 100.0, 5.0), (np.rint(P.synth_batch(2, 96, 82)), 100.0, 5.0),
             (g.image(4).astype(np.float32)[None], *g.sky(4)),
             (P.synth_batch(2, 64, 83).astype(np.float64) + 1e-7, 100.0, 5.0)]
    for x, sky, err in cases:
        a = emul_runner(x, sky, err, 3, 15)
        b = emul_runner(x, sky, err, 3, 15 | abi.PM_RADIX_SORT)
        assert np.array_equal(a["sorted_idx"], b["sorted_idx"]) and np.array_equal(a["n_positive"], b["n_positive"])
        assert np.array_equal(a["rows"], b["rows"], equal_nan=True)


def test_pruned_minapix_on_shifted_galaxies():
    """phase_minapix_auto (groups pulled by warps, fast loop where the aperture provably covers a chunk) picks the same
    centre with and without pruning on every path — centred galaxies (fast loop), galaxies rolled to the frame edge (mirror
    images out of bounds, wrap-around of the rolled aperture: generic loop) — and agrees with the oracle."""
    x = P.synth_batch(6, 96, 85)
    xs = np.concatenate([x, np.roll(x[:3], (-38, 35), axis=(1, 2)), np.roll(x[3:], (30, -41), axis=(1, 2))])
    a = emul_runner_pruned(xs, 100.0, 5.0, 3, 15)
    cfull = emul_runner_pruned(xs, 100.0, 5.0, 3, 15 | abi.PM_NO_PRUNE)
    assert np.array_equal(a["rows"], cfull["rows"], equal_nan=True)
    for extra in (abi.PM_PATH_SPLIT, abi.PM_PATH_SPLIT | abi.PM_TAIL_WARP):
        m = emul_runner_pruned(xs, 100.0, 5.0, 3, 15 | extra)
        np.testing.assert_allclose(m["rows"], a["rows"], rtol=1e-11, atol=0, equal_nan=True)
    P.check_against_oracle(emul_runner, xs[6:10], 100.0, 5.0, 3)
    P.check_against_oracle(emul_runner_pruned, xs[8:12], 100.0, 5.0, 3, check_sorted=False)
