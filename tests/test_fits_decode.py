"""Image payload decode (pm_fits_decode_batch) on the CPU build of the library sources, against numpy; and the oracle's
FITS reader against it on the reference's own frames when they are present (build container only)."""
import glob

import numpy as np
import pytest

from tests import emul_harness as H
from tests.fits_cases import cases


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_decode_every_bitpix(dtype):
    for bitpix, bzero, bscale, payload, want in cases():
        got = H.fits_decode(payload, bitpix, len(want), bzero, bscale, dtype)
        assert np.array_equal(got, want.astype(dtype)), (bitpix, bzero, bscale)


def test_bad_arguments():
    L = H.lib()
    buf = np.zeros(16, np.uint8)
    out = np.zeros(4, np.float32)
    assert L.pm_fits_decode_batch(buf.ctypes.data, 12, 0.0, 1.0, 4, out.ctypes.data, 0, None) != 0      # no such BITPIX
    assert L.pm_fits_decode_batch(buf.ctypes.data, 16, 0.0, 1.0, 4, out.ctypes.data, 7, None) != 0       # no such dtype
    assert L.pm_fits_decode_batch(None, 16, 0.0, 1.0, 0, None, 0, None) == 0                             # empty batch


def test_reference_frames_decode_like_the_oracle_reader():
    files = sorted(glob.glob("/root/reference/sample/*.fits"))[:12]
    if not files:
        pytest.skip("the reference's FITS frames are only present in the build container")
    from oracle import fitsio
    import pawlikmorph_lsst_b200 as pm          # payload_of is host-only code
    for f in files:
        raw = open(f, "rb").read()
        payload, hdr = pm.image.payload_of(raw)
        want = np.asarray(fitsio.getdata(f)).astype(np.float64)
        got = H.fits_decode(payload, hdr["BITPIX"], want.size, hdr.get("BZERO", 0.0), hdr.get("BSCALE", 1.0), np.float64)
        assert np.array_equal(got.reshape(want.shape), want), f
