"""GPU parity at scale: >= 1000 seeded cutouts per size against oracle rows committed in tests/golden/seeded_rows.npz
(made by oracle/make_seeded_rows.py; the images are bit-reproducible from the seed: synth.make_exact).
140^2 with flags A|As (BASELINE.json configs[1]), 256^2 with every statistic (configs[2]), 512^2 with SUPPLIED segmaps
(configs[4]).  Pixelmaps are checked by CRC32 (bit-exact), integer fields exactly, floats to 1e-9."""
import os
import zlib

import numpy as np
import pytest

import synth
from tests.golden_util import GOLDEN_DIR, compare_row

pytestmark = pytest.mark.gpu
Z = None


def _z():
    global Z
    if Z is None:
        Z = np.load(os.path.join(GOLDEN_DIR, "seeded_rows.npz"))
    return Z


def _run(x, flags, **kw):
    import torch

    import pawlikmorph_lsst_b200 as pm
    r = pm.batch.analyse_batch(torch.from_numpy(x).cuda(), synth.SKY, synth.SKY_ERR, 3, flags, want_segmap=True, **kw)
    torch.cuda.synchronize()
    return r.rows.cpu().numpy(), r.segmap.cpu().numpy()


def _check(name, rows, segs, want_rows, want_crc, skip=()):
    fields = [str(f) for f in _z()["fields"]]
    bad = []
    for i in range(len(rows)):
        if zlib.crc32(np.packbits(segs[i]).tobytes()) != int(want_crc[i]):
            bad.append(f"{name}[{i}] pixelmap differs")
            continue
        got, want = dict(zip(fields, rows[i])), dict(zip(fields, want_rows[i]))
        for k in skip:
            got[k] = want[k]
        bad += compare_row(got, want, 1e-9, f"{name}[{i}] ")
        if got["n_candidates"] != want["n_candidates"]:
            bad.append(f"{name}[{i}] n_candidates {got['n_candidates']} != {want['n_candidates']}")
    assert not bad, f"{len(bad)} mismatches, first: " + "; ".join(bad[:5])


@pytest.mark.parametrize("name", ["n140_flags3", "n256_flags15"])
def test_thousand_seeded_cutouts(name):
    z = _z()
    n, flags, count, seed = (int(v) for v in z[name + "_meta"])
    x = synth.make_exact(count, n, seed)
    assert [zlib.crc32(x[i].tobytes()) for i in range(0, count, 97)] == [int(v) for v in z[name + "_imgcrc"][::97]], \
        "synth.make_exact is not bit-reproducible on this machine"
    rows, segs = _run(x, flags)
    _check(name, rows, segs, z[name + "_rows"], z[name + "_crc"])
    assert (z[name + "_rows"][:, 14] == 0).sum() >= 0.95 * count


def test_512_with_supplied_segmaps():
    import pawlikmorph_lsst_b200 as pm
    z = _z()
    name = "n512_segmap"
    n, flags, count, seed = (int(v) for v in z[name + "_meta"])
    x = synth.make_exact(count, n, seed)
    _, segs = _run(x, flags | pm._abi.PM_STOP_AFTER_PIXELMAP)                    # the pixelmap kernel makes the maps ...
    assert [zlib.crc32(np.packbits(s).tobytes()) for s in segs] == [int(v) for v in z[name + "_crc"]]
    import torch
    rows, segs2 = _run(x, flags, segmap_in=torch.from_numpy(segs).cuda())         # ... and the path takes them as input
    assert np.array_equal(segs, segs2)
    _check(name, rows, segs2, z[name + "_rows"], z[name + "_crc"])
    rows_s, _ = _run(x, flags | pm._abi.PM_PATH_SPLIT, segmap_in=torch.from_numpy(segs).cuda())
    np.testing.assert_allclose(rows_s, rows, rtol=1e-11, atol=0, equal_nan=True)
