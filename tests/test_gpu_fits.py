"""GPU side of the image payload decode: every BITPIX against numpy, and the payload host API against the float API."""
import numpy as np
import pytest

from tests import parity_checks as P
from tests.fits_cases import cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("count", [4099, 8 * 4096 + 3])
def test_decode_every_bitpix_on_the_gpu(count):
    import torch

    import pawlikmorph_lsst_b200 as pm
    for bitpix, bzero, bscale, payload, want in cases(count):
        raw = torch.frombuffer(bytearray(payload), dtype=torch.uint8).cuda()
        for dt, npdt in ((torch.float32, np.float32), (torch.float64, np.float64)):
            got = pm.image.decode_payload(raw, bitpix, (count,), bzero, bscale, dt).cpu().numpy()
            assert np.array_equal(got, want.astype(npdt)), (bitpix, bzero, bscale, dt)
        # an unaligned view takes the scalar path
        if count > 64:
            bpp = pm.image.BYTES_PER_PIXEL[bitpix]
            got = pm.image.decode_payload(raw[bpp:], bitpix, (count - 1,), bzero, bscale, torch.float64).cpu().numpy()
            assert np.array_equal(got, want[1:])


def test_payload_host_api_equals_the_float_api():
    import pawlikmorph_lsst_b200 as pm
    x = np.rint(P.synth_batch(6, 64, 33)).astype(np.float32)          # integer counts, as a 16-bit frame stores them
    assert x.min() >= 0 and x.max() < 65536
    stored = (x.astype(np.int32) - 32768).astype(">i2")                # BZERO = 32768 convention
    payload = np.frombuffer(stored.tobytes(), dtype=np.uint8).reshape(6, -1)
    a = pm.batch.analyse_host(x, 100.0, 5.0, 3, 15, chunk=4)
    b = pm.batch.analyse_host_payload(payload, 16, 64, 100.0, 5.0, bzero=32768.0, bscale=1.0, filter_size=3, flags=15, chunk=4)
    assert np.array_equal(a, b)
    f = pm.batch.analyse_host_payload(np.frombuffer(x.astype(">f4").tobytes(), dtype=np.uint8).reshape(6, -1), -32, 64, 100.0, 5.0,
                                      filter_size=3, flags=15, chunk=4)
    assert np.array_equal(a, f)


@pytest.mark.parametrize("npix", [20, 7, 36, 32, 24, 8])
def test_window_decode_every_bitpix_on_the_gpu(npix):
    """(multiples of 8 take the aligned-word kernel for the 16-bit frames; the last window ends on the last pixel of the payload)"""
    import torch

    import pawlikmorph_lsst_b200 as pm
    from tests.fits_cases import window_cases
    org = np.array([[0, 0], [1, 5], [0, 3]], dtype=np.int32)
    if npix % 8 == 0:
        org = np.array([[0, 0], [1, 5], [37 - npix, 41 - npix]], dtype=np.int32)
    for bitpix, bzero, bscale, payload, frames in window_cases():
        raw = torch.frombuffer(bytearray(payload), dtype=torch.uint8).cuda()
        for dt, npdt in ((torch.float32, np.float32), (torch.float64, np.float64)):
            got = pm.image.decode_window(raw, bitpix, (37, 41), org, npix, bzero, bscale, dt).cpu().numpy()
            want = np.stack([frames[b, org[b, 0]:org[b, 0] + npix, org[b, 1]:org[b, 1] + npix] for b in range(3)]).astype(npdt)
            assert np.array_equal(got, want), (bitpix, bzero, bscale, dt)
    with pytest.raises(pm.image.PartialOverlapError):
        pm.image.decode_window(raw, bitpix, (37, 41), org + 10, 36)
