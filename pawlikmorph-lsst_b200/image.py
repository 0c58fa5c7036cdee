"""FITS image payloads on the device — the input format of the morphology path.

The reference ingests every frame with astropy (`Image.py:132-134`: `fits.getdata(filename, header=True)` followed
by `img.byteswap().newbyteorder()`) and analyses `image.astype(np.float64)` (`Image.py:148`).  Here the payload —
the big-endian pixel block as it lies in the file, header stripped — is copied to the GPU as bytes and decoded there
(`pm_fits_decode_batch`), which halves the host->device traffic of the 16-bit SDSS frames on a path that is PCIe bound
end to end.  Header parsing (2880-byte blocks of 80-character cards) stays on the host: `payload_of` does the minimum
(BITPIX, NAXIS1/2, BZERO, BSCALE) without astropy.
"""
import numpy as np
import torch

from . import _abi
from ._lib import lib

BYTES_PER_PIXEL = {8: 1, 16: 2, 32: 4, -32: 4, -64: 8}


def payload_of(raw: bytes):
    """(payload bytes, header dict) of the primary HDU of a FITS file image held in memory."""
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
                val = val == "T"
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
            break
    nbytes = BYTES_PER_PIXEL[hdr["BITPIX"]] * hdr["NAXIS1"] * hdr["NAXIS2"]
    if len(raw) < off + nbytes:
        raise IOError("truncated FITS data")
    return raw[off:off + nbytes], hdr


def decode_payload(payload, bitpix, shape, bzero=0.0, bscale=1.0, dtype=torch.float32, stream=None):
    """Decode big-endian FITS pixels that already live on the GPU.

    payload: CUDA uint8 tensor holding prod(shape) pixels of type `bitpix`; returns a CUDA tensor of `shape` and
    `dtype` (float32 or float64) with physical values BZERO + BSCALE * stored."""
    if not torch.is_tensor(payload) or not payload.is_cuda or payload.dtype != torch.uint8:
        raise TypeError("decode_payload needs a CUDA uint8 tensor (there is no CPU path)")
    if bitpix not in BYTES_PER_PIXEL:
        raise ValueError(f"unsupported BITPIX {bitpix}")
    if dtype not in (torch.float32, torch.float64):
        raise TypeError("dtype must be float32 or float64")
    count = int(np.prod(shape))
    payload = payload.contiguous()
    if payload.numel() != count * BYTES_PER_PIXEL[bitpix]:
        raise ValueError(f"payload holds {payload.numel()} bytes, expected {count * BYTES_PER_PIXEL[bitpix]}")
    dev = payload.device
    L = lib()
    with torch.cuda.device(dev):
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        with torch.cuda.stream(st):
            out = torch.empty(tuple(shape), dtype=dtype, device=dev)
            rc = L.pm_fits_decode_batch(payload.data_ptr(), int(bitpix), float(bzero), float(bscale), count, out.data_ptr(),
                                        _abi.PM_F32 if dtype == torch.float32 else _abi.PM_F64, st.cuda_stream)
            _abi.check(L, rc, "pm_fits_decode_batch")
            payload.record_stream(st)
    return out
