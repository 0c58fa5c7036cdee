"""FITS payload cases shared by the CPU (emulator build) and GPU tests of the image decode: big-endian pixel blocks
of every BITPIX with the scalings the reference's frames use, and the numpy statement of what `fits.getdata` +
`astype(float64)` (Image.py:132-134, :148) makes of them."""
import numpy as np

DT = {8: ">u1", 16: ">i2", 32: ">i4", -32: ">f4", -64: ">f8"}


def cases(count=4099, seed=5):
    rng = np.random.default_rng(seed)
    out = []
    for bitpix, bzero, bscale in ((16, 32768.0, 1.0),      # unsigned-16 convention of the bundled SDSS frames
                                  (16, 0.0, 1.0), (16, 1000.5, 0.25), (-32, 0.0, 1.0), (-32, 3.0, 2.0),
                                  (8, 0.0, 1.0), (32, 0.0, 1.0), (32, -7.0, 0.5), (-64, 0.0, 1.0),
                                  (8, -128.0, 1.0), (16, 0.5, 1.0),          # signed-byte convention; non-integer zero point
                                  (16, 3.3, 0.1), (32, 1.7, 1e-3), (-64, 0.1, 1.1), (-32, 0.3, 0.7)):   # a rounded multiply THEN a rounded add
        if bitpix == 8:
            stored = rng.integers(0, 256, count).astype(np.uint8)
        elif bitpix == 16:
            stored = rng.integers(-32768, 32768, count).astype(np.int16)
            stored[:4] = [-32768, 32767, 0, -1]
        elif bitpix == 32:
            stored = rng.integers(-2 ** 31, 2 ** 31, count).astype(np.int32)
        elif bitpix == -32:
            stored = (rng.normal(100, 50, count)).astype(np.float32)
            stored[:3] = [0.0, -0.0, 1e-38]
        else:
            stored = rng.normal(100, 50, count)
        payload = stored.astype(DT[bitpix]).tobytes()
        want = stored.astype(np.float64)
        if not (bscale == 1.0 and bzero == 0.0):
            want = want * bscale + bzero
        out.append((bitpix, bzero, bscale, payload, want))
    return out


def window_cases(B=3, H=37, W=41, seed=9):
    """(bitpix, bzero, bscale, payload bytes of B frames, float64 frames [B, H, W]) for the cut-out decode."""
    out = []
    for bitpix, bzero, bscale, payload, want in cases(B * H * W, seed):
        out.append((bitpix, bzero, bscale, payload, want.reshape(B, H, W)))
    return out
