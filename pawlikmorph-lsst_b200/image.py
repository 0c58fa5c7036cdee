"""FITS image payloads on the device — the input format of the morphology path.

The reference ingests every frame with astropy (`Image.py:132-134`: `fits.getdata(filename, header=True)` followed
by `img.byteswap().newbyteorder()`) and analyses `image.astype(np.float64)` (`Image.py:148`).  Here the payload —
the big-endian pixel block as it lies in the file, header stripped — is copied to the GPU as bytes and decoded there
(`pm_fits_decode_batch`), which halves the host->device traffic of the 16-bit SDSS frames on a path that is PCIe bound
end to end.  Header parsing (2880-byte blocks of 80-character cards) stays on the host: `payload_of` does the minimum
(BITPIX, NAXIS1/2, BZERO, BSCALE) without astropy.

The WCS cut-out of `Image.py:175-186` (`skycoord_to_pixel` + `Cutout2D(mode="strict")`) is split the same way: the
sky -> pixel inversion of the TAN projection and the window arithmetic are a few scalars per frame on the host
(`sky_to_pixel`, `cutout_origin`), the copy of the window out of the frame is fused with the decode on the device
(`decode_window` -> `pm_fits_decode_window_batch`).  astropy is not installed in the build container, so these two host
functions are restated from the FITS WCS paper II (gnomonic projection) and from astropy's documented
`overlap_slices` behaviour — parity unpinned; `tests/golden/sdss_wcs.json` anchors them on the reference's own frames
(every catalogue position of images.csv falls within half a pixel of its frame's centre).
"""
import math

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


def header_cards(raw: bytes):
    """The 80-character cards of the primary header of a FITS file image held in memory (without END)."""
    cards, off = [], 0
    while True:
        block = raw[off:off + 2880]
        if len(block) < 2880:
            raise IOError("truncated FITS header")
        off += 2880
        for i in range(0, 2880, 80):
            card = block[i:i + 80].decode("ascii", "replace")
            if card[:8].strip() == "END":
                return cards
            cards.append(card)


def fits_image_bytes(data, cards=()):
    """What `fits.PrimaryHDU(data=data, header=header).writeto(...)` stores for a 2-D float64 array (main.py:447-461: the clean image
    and the pixelmap): the mandatory cards for BITPIX -64, then the cards of the source header except the structural and scaling
    ones, big-endian pixels, 2880-byte blocks."""
    a = np.ascontiguousarray(np.asarray(data, dtype=np.float64))
    if a.ndim != 2:
        raise ValueError("fits_image_bytes: a 2-D image")

    def card(key, val, comment=""):
        v = f"{'T' if val else 'F':>20}" if isinstance(val, bool) else f"{val:>20}"
        return f"{key:<8}= {v}{' / ' + comment if comment else ''}".ljust(80)[:80]

    out = [card("SIMPLE", True, "conforms to FITS standard"), card("BITPIX", -64, "array data type"), card("NAXIS", 2, "number of array dimensions"),
           card("NAXIS1", a.shape[1]), card("NAXIS2", a.shape[0])]
    skip = {"SIMPLE", "BITPIX", "NAXIS", "NAXIS1", "NAXIS2", "BZERO", "BSCALE", "EXTEND", "END"}
    out += [c.ljust(80)[:80] for c in cards if c[:8].strip() not in skip]
    out.append("END".ljust(80))
    head = "".join(out).encode("ascii", "replace")
    head += b" " * (-len(head) % 2880)
    body = a.astype(">f8").tobytes()
    return head + body + b"\0" * (-len(body) % 2880)


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


class PartialOverlapError(ValueError):
    """The window leaves the frame (astropy.nddata.PartialOverlapError, which helpers.py:5 imports to skip such images)."""


def sky_to_pixel(hdr, ra, dec):
    """0-based pixel coordinates (x = column, y = row) of the sky position (ra, dec) in degrees for a `RA---TAN` /
    `DEC--TAN` header with a PC (+ CDELT) or CD matrix — what `wcs.utils.skycoord_to_pixel(position, wcs)` returns at
    Image.py:177-178 for the bundled frames (all of them TAN, LONPOLE 180, no distortion terms)."""
    t1, t2 = str(hdr.get("CTYPE1", "")).strip(), str(hdr.get("CTYPE2", "")).strip()
    if t1 == "RA---TAN" and t2 == "DEC--TAN":
        lon = 1
    elif t1 == "DEC--TAN" and t2 == "RA---TAN":
        lon = 2                                                                               # reversed axes (Image.py:179-182)
    else:
        raise NotImplementedError(f"only RA---TAN / DEC--TAN headers are handled, got {t1!r}, {t2!r}")
    # astropy (which the reference calls) applies SIP / TPV distortions and a CROTA2 rotation; they are not implemented
    # here, so such headers are refused instead of being projected a few pixels off
    if any(k in hdr for k in ("A_ORDER", "B_ORDER")) or any(str(k).startswith(("PV1_", "PV2_")) for k in hdr):
        raise NotImplementedError("distorted TAN headers (SIP / PV terms) are not handled")
    if "CROTA2" in hdr and float(hdr["CROTA2"]) != 0.0 and not any(k in hdr for k in ("PC1_1", "PC1_2", "PC2_1", "PC2_2", "CD1_1")):
        raise NotImplementedError("CROTA2 rotation without a PC / CD matrix is not handled")
    a0, d0 = math.radians(float(hdr[f"CRVAL{lon}"])), math.radians(float(hdr[f"CRVAL{3 - lon}"]))
    a, d = math.radians(float(ra)), math.radians(float(dec))
    cosc = math.sin(d0) * math.sin(d) + math.cos(d0) * math.cos(d) * math.cos(a - a0)
    if cosc <= 0:
        raise ValueError("position is on the far side of the tangent point")
    xi = math.degrees(math.cos(d) * math.sin(a - a0) / cosc)                                   # intermediate world coordinates
    eta = math.degrees((math.cos(d0) * math.sin(d) - math.sin(d0) * math.cos(d) * math.cos(a - a0)) / cosc)
    if lon == 2:
        xi, eta = eta, xi                                                                     # axis 1 carries the latitude
    if "PC1_1" in hdr or "PC2_2" in hdr or "CD1_1" not in hdr:                                # wcslib: PC wins when both are given
        c1, c2 = float(hdr.get("CDELT1", 1.0)), float(hdr.get("CDELT2", 1.0))
        m = [[c1 * float(hdr.get("PC1_1", 1.0)), c1 * float(hdr.get("PC1_2", 0.0))],
             [c2 * float(hdr.get("PC2_1", 0.0)), c2 * float(hdr.get("PC2_2", 1.0))]]
    else:
        m = [[float(hdr["CD1_1"]), float(hdr.get("CD1_2", 0.0))], [float(hdr.get("CD2_1", 0.0)), float(hdr["CD2_2"])]]
    det = m[0][0] * m[1][1] - m[0][1] * m[1][0]
    px = (m[1][1] * xi - m[0][1] * eta) / det + float(hdr["CRPIX1"])
    py = (-m[1][0] * xi + m[0][0] * eta) / det + float(hdr["CRPIX2"])
    return px - 1.0, py - 1.0                                                                # FITS pixels are 1-based


def cutout_origin(x, y, npix, shape):
    """(first row, first column) of the npix x npix window `Cutout2D(data, (x, y), (npix, npix), mode="strict")` takes
    (Image.py:184): astropy's overlap_slices puts the lower edge at ceil(position - npix / 2) on both axes and raises
    PartialOverlapError when the window is not wholly inside the frame."""
    r0, c0 = int(math.ceil(y - npix / 2.0)), int(math.ceil(x - npix / 2.0))
    if r0 < 0 or c0 < 0 or r0 + npix > shape[0] or c0 + npix > shape[1]:
        raise PartialOverlapError(f"{npix} x {npix} window at row {r0}, column {c0} leaves the {shape[0]} x {shape[1]} frame")
    return r0, c0


def decode_window(payload, bitpix, frame_shape, origins, npix, bzero=0.0, bscale=1.0, dtype=torch.float32, stream=None):
    """Decode one npix x npix window out of each of B frames that already live on the GPU.

    payload: CUDA uint8 tensor with B frames of frame_shape = (H, W) pixels of type `bitpix`; origins: int32 [B, 2]
    (first row, first column), CUDA or host; returns a CUDA tensor [B, npix, npix] of `dtype`."""
    if not torch.is_tensor(payload) or not payload.is_cuda or payload.dtype != torch.uint8:
        raise TypeError("decode_window needs a CUDA uint8 tensor (there is no CPU path)")
    if bitpix not in BYTES_PER_PIXEL:
        raise ValueError(f"unsupported BITPIX {bitpix}")
    if dtype not in (torch.float32, torch.float64):
        raise TypeError("dtype must be float32 or float64")
    H, W = int(frame_shape[0]), int(frame_shape[1])
    payload = payload.contiguous()
    per = H * W * BYTES_PER_PIXEL[bitpix]
    if payload.numel() % per:
        raise ValueError(f"payload holds {payload.numel()} bytes, not a multiple of {per}")
    B = payload.numel() // per
    dev = payload.device
    org = torch.as_tensor(origins, dtype=torch.int32).reshape(-1, 2)
    if org.shape[0] != B:
        raise ValueError(f"{B} frames but {org.shape[0]} window origins")
    oc = org.cpu() if org.is_cuda else org
    if B and (int(oc.min()) < 0 or int(oc[:, 0].max()) + npix > H or int(oc[:, 1].max()) + npix > W):
        raise PartialOverlapError("a window leaves its frame")
    L = lib()
    with torch.cuda.device(dev):
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        with torch.cuda.stream(st):
            org_d = org.to(dev, non_blocking=True).contiguous()
            out = torch.empty((B, npix, npix), dtype=dtype, device=dev)
            rc = L.pm_fits_decode_window_batch(payload.data_ptr(), int(bitpix), float(bzero), float(bscale), B, H, W,
                                               org_d.data_ptr(), int(npix), out.data_ptr(),
                                               _abi.PM_F32 if dtype == torch.float32 else _abi.PM_F64, st.cuda_stream)
            _abi.check(L, rc, "pm_fits_decode_window_batch")
            payload.record_stream(st)
            org_d.record_stream(st)
    return out
