"""Minimal FITS writer for the tests: a golden frame back into the format the reference's sample/ holds
(primary HDU, BITPIX 16 with BZERO 32768 — the unsigned-16 convention — or BITPIX -32)."""
import numpy as np


def _card(key, val, comment=""):
    if isinstance(val, bool):
        v = f"{'T' if val else 'F':>20}"
    elif isinstance(val, str):
        v = f"'{val:<8}'"
    else:
        v = f"{val:>20}"
    return f"{key:<8}= {v}{' / ' + comment if comment else ''}".ljust(80)[:80]


def fits_bytes(img, bitpix=16, extra=None):
    """extra: further header cards (e.g. the WCS cards of tests/golden/sdss_wcs.json) in place of the bare CTYPE1."""
    img = np.asarray(img)
    cards = [_card("SIMPLE", True, "conforms to FITS standard"), _card("BITPIX", bitpix), _card("NAXIS", 2),
             _card("NAXIS1", img.shape[1]), _card("NAXIS2", img.shape[0])]
    if bitpix == 16:
        cards += [_card("BSCALE", 1), _card("BZERO", 32768)]
        data = (img.astype(np.int64) - 32768).astype(">i2").tobytes()
    elif bitpix == -32:
        data = img.astype(">f4").tobytes()
    else:
        raise ValueError(bitpix)
    if extra:
        cards += [_card(k, v) for k, v in extra.items() if not k.startswith("NAXIS")]
    else:
        cards.append(_card("CTYPE1", "RA---TAN"))
    cards.append("END".ljust(80))
    head = "".join(cards).encode("ascii")
    head += b" " * (-len(head) % 2880)
    return head + data + b"\0" * (-len(data) % 2880)
