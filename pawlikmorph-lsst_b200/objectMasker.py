"""objectMasker.py of the reference (`objectOccluded`, :16-74; `_findStars`, :77-162; `_getObject`, :165-202): which catalogue
objects lie on the object's pixelmap.  Host functions — a catalogue file, a few sky -> pixel conversions and a 2 x 2 window
look-up per object; the sky -> pixel step is `image.sky_to_pixel` (TAN headers) where upstream goes through astropy's WCS."""
import sys

import numpy as np

__all__ = ["objectOccluded"]


def _getObject(ra, dec, psfMags, types, xmin, xmax, ymin, ymax):
    """objectMasker.py:165-202: [ra, dec, type, psfMag] of the objects strictly inside the box (zip stops at the shortest list:
    psfMags holds the stars only, so objects appended after them never pass — kept as upstream)."""
    pos = []
    for i, j, k, p in zip(ra, dec, types, psfMags):
        if i > xmin and i < xmax:
            if j > ymin and j < ymax:
                pos.append([i, j, k, p])
    return pos


def _findStars(catalogue, ra, dec, galaxy, cosmicray, unknown, size=0.008333333333333333):
    """objectMasker.py:77-162: the catalogue rows (objID, ra, dec, psfMag_r, type) within +-size degrees of (ra, dec)."""
    import pandas as pd                      # the reference's reader: its fast float parser decides the last digit of ra / dec / mag
    df = pd.read_csv(catalogue, dtype={"objID": "float", "ra": "float", "dec": "float", "psfMag_r": "float", "type": "str"})
    xmin, xmax, ymin, ymax = ra - size, ra + size, dec - size, dec + size

    def take(kind, cols):
        sel = df.loc[df["type"] == kind]
        return [list(sel[c]) for c in cols]

    try:
        ras, decs, psfMags = take("STAR", ("ra", "dec", "psfMag_r"))
    except KeyError:
        print("Supplied catalogue has wrong format. Expected: objId, RA, DEC, psfMag_r, type")
        sys.exit()
    types = ["STAR" for _ in range(len(ras))]
    for wanted, kind in ((galaxy, "GALAXY"), (cosmicray, "COSMIC_RAY"), (unknown, "UNKNOWN")):
        if wanted:
            r2, d2 = take(kind, ("ra", "dec"))
            ras += r2
            decs += d2
            types += [kind for _ in range(len(ras))]
    return _getObject(ras, decs, psfMags, types, xmin, xmax, ymin, ymax)


def objectOccluded(segmap, radec, catalogue, header, galaxy=False, cosmicray=False, unknown=False):
    """objectMasker.py:16-74: (True, objects) when a catalogue object near (ra, dec) falls on the pixelmap — the 2 x 2 window
    segmap[int(y)-1:int(y)+1, int(x)-1:int(x)+1] around its pixel position holds a 1 —, else (False, [])."""
    from .image import sky_to_pixel
    ra, dec = radec
    found = []
    segmap = np.asarray(segmap)
    for obj in _findStars(catalogue, ra, dec, galaxy, cosmicray, unknown):
        x, y = sky_to_pixel(header, obj[0], obj[1])
        if np.any(segmap[int(y) - 1:int(y) + 1, int(x) - 1:int(x) + 1] == 1):
            found.append(obj)
    if len(found) > 0:
        return True, found
    return False, []
