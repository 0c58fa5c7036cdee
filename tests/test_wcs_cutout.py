"""Host side of the WCS cut-out (Image.py:175-186) and the window decode on the CPU build of the library sources."""
import json
import math
import os

import numpy as np
import pytest

import pawlikmorph_lsst_b200 as pm
from tests import emul_harness as H
from tests.fits_cases import window_cases

WCS = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sdss_wcs.json")))


def pixel_to_sky(h, x, y):
    """Forward TAN projection (FITS WCS paper II, eqs. 54-55 + the native -> celestial rotation for LONPOLE 180), written
    independently of image.sky_to_pixel: 0-based pixel -> (ra, dec) in degrees."""
    u, v = x + 1 - h["CRPIX1"], y + 1 - h["CRPIX2"]
    xi = math.radians(h["CDELT1"] * (h["PC1_1"] * u + h["PC1_2"] * v))
    eta = math.radians(h["CDELT2"] * (h["PC2_1"] * u + h["PC2_2"] * v))
    a0, d0 = math.radians(h["CRVAL1"]), math.radians(h["CRVAL2"])
    den = math.cos(d0) - eta * math.sin(d0)
    ra = a0 + math.atan2(xi, den)
    dec = math.atan2((math.sin(d0) + eta * math.cos(d0)) * math.cos(ra - a0), den)
    return math.degrees(ra), math.degrees(dec)


def test_catalogue_positions_fall_on_the_frame_centres():
    """all 68 rows of images.csv are RA---TAN frames centred on their catalogue position (within half a pixel);
    the 128-pixel window starts at 51 or 52 in the 231-pixel frames and at 24...44 in the smaller ones."""
    assert len(WCS) == 68
    for e in WCS:
        h = e["header"]
        x, y = pm.image.sky_to_pixel(h, e["ra"], e["dec"])
        assert abs(x - (h["NAXIS1"] - 1) / 2) <= 0.5 and abs(y - (h["NAXIS2"] - 1) / 2) <= 0.5, e["file"]
        r0, c0 = pm.image.cutout_origin(x, y, 128, (h["NAXIS2"], h["NAXIS1"]))
        if h["NAXIS1"] == 231:
            assert r0 in (51, 52) and c0 in (51, 52)
        else:
            assert 24 <= r0 <= 44 and 24 <= c0 <= 44


def test_sky_to_pixel_inverts_the_forward_projection():
    rng = np.random.default_rng(4)
    for e in WCS[::5]:
        h = e["header"]
        for _ in range(20):
            x, y = rng.uniform(-300, 600, 2)
            ra, dec = pixel_to_sky(h, x, y)
            gx, gy = pm.image.sky_to_pixel(h, ra, dec)
            assert abs(gx - x) < 1e-6 and abs(gy - y) < 1e-6
        hcd = {k: v for k, v in h.items() if not k.startswith("PC") and not k.startswith("CDELT")}     # CD-only header
        if "CD1_1" in hcd:
            assert np.allclose(pm.image.sky_to_pixel(hcd, e["ra"], e["dec"]), pm.image.sky_to_pixel(h, e["ra"], e["dec"]), atol=1e-6)
    # reversed axes: the same sky position, latitude on axis 1
    h = WCS[0]["header"]
    hr = dict(h, CTYPE1=h["CTYPE2"], CTYPE2=h["CTYPE1"], CRVAL1=h["CRVAL2"], CRVAL2=h["CRVAL1"],
              PC1_1=h["PC2_1"], PC1_2=h["PC2_2"], PC2_1=h["PC1_1"], PC2_2=h["PC1_2"])
    hr = {k: v for k, v in hr.items() if not k.startswith("CD")}
    assert np.allclose(pm.image.sky_to_pixel(hr, WCS[0]["ra"], WCS[0]["dec"]), pm.image.sky_to_pixel(h, WCS[0]["ra"], WCS[0]["dec"]), atol=1e-6)
    with pytest.raises(NotImplementedError):
        pm.image.sky_to_pixel(dict(h, CTYPE1="RA---SIN"), 1.0, 2.0)


def test_cutout_origin_follows_overlap_slices():
    f = pm.image.cutout_origin
    assert f(115.0, 115.0, 128, (231, 231)) == (51, 51)             # ceil(115 - 64)
    assert f(115.3, 114.7, 128, (231, 231)) == (51, 52)             # (row from y, column from x)
    assert f(64.0, 64.0, 128, (128, 128)) == (0, 0)
    assert f(10.2, 10.2, 5, (30, 30)) == (8, 8)                     # odd sizes: ceil(10.2 - 2.5)
    for x, y in ((63.0, 64.0), (64.0, 64.6), (200.0, 100.0)):
        with pytest.raises(pm.image.PartialOverlapError):
            f(x, y, 128, (128, 231))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_window_decode_every_bitpix(dtype):
    org = np.array([[0, 0], [5, 9], [17, 21]], dtype=np.int32)
    for npix in (20, 7):
        for bitpix, bzero, bscale, payload, frames in window_cases():
            got = H.fits_decode_window(payload, bitpix, 3, 37, 41, org, npix, bzero, bscale, dtype)
            want = np.stack([frames[b, org[b, 0]:org[b, 0] + npix, org[b, 1]:org[b, 1] + npix] for b in range(3)]).astype(dtype)
            assert np.array_equal(got, want), (bitpix, bzero, bscale, npix)


def test_window_decode_bad_arguments():
    L = H.lib()
    buf = np.zeros(64, np.uint8)
    out = np.zeros(16, np.float32)
    org = np.zeros(2, np.int32)
    a = (buf.ctypes.data, 16, 0.0, 1.0, 1, 4, 8, org.ctypes.data)
    assert L.pm_fits_decode_window_batch(*a, 4, out.ctypes.data, 0, None) == 0
    assert L.pm_fits_decode_window_batch(*a, 5, out.ctypes.data, 0, None) != 0                      # window larger than the frame
    assert L.pm_fits_decode_window_batch(buf.ctypes.data, 16, 0.0, 1.0, 1, 4, 8, None, 4, out.ctypes.data, 0, None) != 0
    assert L.pm_fits_decode_window_batch(None, 16, 0.0, 1.0, 0, 4, 8, None, 4, None, 0, None) == 0    # empty batch
