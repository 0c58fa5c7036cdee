"""Minimal FITS primary-HDU reader (test infrastructure; NOT part of the product path).

Replaces ``astropy.io.fits.getdata`` for the reference's bundled fixtures
(`/root/reference/sample/*.fits`, read by the reference at `Image.py:128-134`).
Only what those files need: 2880-byte blocks, 80-char cards, BITPIX 16/32/-32/-64,
BSCALE/BZERO, two axes.
"""
import numpy as np

_DT = {8: ">u1", 16: ">i2", 32: ">i4", -32: ">f4", -64: ">f8"}


def read_header(raw: bytes):
    hdr, off = {}, 0
    while True:
        block = raw[off:off + 2880]
        if len(block) < 2880:
            raise IOError("truncated FITS header")
        off += 2880
        done = False
        for i in range(0, 2880, 80):
            card = block[i:i + 80].decode("ascii", "replace")
            key = card[:8].strip()
            if key == "END":
                done = True
                break
            if card[8:10] != "= ":
                continue
            val = card[10:].split("/")[0].strip()
            if val.startswith("'"):
                val = val.strip("'").strip()
            elif val in ("T", "F"):
                val = (val == "T")
            else:
                try:
                    val = int(val)
                except ValueError:
                    try:
                        val = float(val)
                    except ValueError:
                        pass
            hdr.setdefault(key, val)
        if done:
            return hdr, off


def getdata(path, with_header=False):
    """Return the primary image as a native-endian numpy array (physical values)."""
    raw = open(path, "rb").read()
    hdr, off = read_header(raw)
    bitpix, n1, n2 = hdr["BITPIX"], hdr["NAXIS1"], hdr["NAXIS2"]
    arr = np.frombuffer(raw, dtype=_DT[bitpix], count=n1 * n2, offset=off).reshape(n2, n1)
    bzero, bscale = hdr.get("BZERO", 0), hdr.get("BSCALE", 1)
    if bitpix > 0 and (bzero != 0 or bscale != 1):
        if bitpix == 16 and bzero == 32768 and bscale == 1:
            out = (arr.astype(np.int32) + 32768).astype(np.uint16)
        else:
            out = arr.astype(np.float64) * bscale + bzero
    else:
        out = arr.astype(arr.dtype.newbyteorder("="))
    return (out, hdr) if with_header else out
