"""objectMasker.py of the reference (`objectOccluded`, :16-74; `_findStars`, :77-162; `_getObject`, :165-202): which catalogue
objects lie on the object's pixelmap.  Host functions — a catalogue file, a few sky -> pixel conversions and a 2 x 2 window
look-up per object; the sky -> pixel step is `image.sky_to_pixel` (TAN headers) where upstream goes through astropy's WCS."""
import sys

import numpy as np

__all__ = ["objectOccluded"]


def _getObject(ra, dec, psfMags, types, xmin, xmax, ymin, ymax):
    """objectMasker.py:165-202: [ra, dec, type, psfMag] of the objects strictly inside the box.  The four lists are walked together
    and the walk ends with the shortest: psfMags holds the stars only, so galaxies / cosmic rays / unknown objects appended after
    them are never returned — kept, it is what upstream does."""
    m = min(len(ra), len(dec), len(types), len(psfMags))
    a, d = np.asarray(ra[:m], dtype=np.float64), np.asarray(dec[:m], dtype=np.float64)
    inside = np.flatnonzero((a > xmin) & (a < xmax) & (d > ymin) & (d < ymax))
    return [[ra[i], dec[i], types[i], psfMags[i]] for i in inside]


def _findStars(catalogue, ra, dec, galaxy, cosmicray, unknown, size=0.008333333333333333):
    """objectMasker.py:77-162: the catalogue rows (objID, ra, dec, psfMag_r, type) within +-size degrees (0.5 arcmin) of (ra, dec)."""
    import pandas as pd                      # the reference's reader: its fast float parser decides the last digit of ra / dec / mag
    table = pd.read_csv(catalogue, dtype={"objID": "float", "ra": "float", "dec": "float", "psfMag_r": "float", "type": "str"})

    def columns(kind, names):
        rows = table.loc[table["type"] == kind]
        return [list(rows[c]) for c in names]

    try:
        ras, decs, mags = columns("STAR", ("ra", "dec", "psfMag_r"))
    except KeyError:
        print("Supplied catalogue has wrong format. Expected: objId, RA, DEC, psfMag_r, type")
        sys.exit()
    kinds = ["STAR"] * len(ras)
    for wanted, kind in ((galaxy, "GALAXY"), (cosmicray, "COSMIC_RAY"), (unknown, "UNKNOWN")):
        if wanted:
            more_ra, more_dec = columns(kind, ("ra", "dec"))
            ras, decs = ras + more_ra, decs + more_dec
            kinds += [kind] * len(ras)                       # (upstream's count: the list length AFTER the extension)
    return _getObject(ras, decs, mags, kinds, ra - size, ra + size, dec - size, dec + size)


def objectOccluded(segmap, radec, catalogue, header, galaxy=False, cosmicray=False, unknown=False):
    """objectMasker.py:16-74: (True, objects) when a catalogue object near (ra, dec) falls on the pixelmap — the 2 x 2 window
    segmap[int(y)-1:int(y)+1, int(x)-1:int(x)+1] around its pixel position holds a 1 —, else (False, [])."""
    from .image import sky_to_pixel
    segmap = np.asarray(segmap)
    on_map = []
    for obj in _findStars(catalogue, radec[0], radec[1], galaxy, cosmicray, unknown):
        x, y = (int(v) for v in sky_to_pixel(header, obj[0], obj[1]))
        if (segmap[y - 1:y + 1, x - 1:x + 1] == 1).any():
            on_map.append(obj)
    return (True, on_map) if on_map else (False, [])
