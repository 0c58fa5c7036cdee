"""tests/golden/sdss_wcs.json: the WCS cards of the frames images.csv lists (the reference's own sample/ headers) with the
catalogue position of each row.  Run in the build container (needs /root/reference): python -m oracle.make_wcs_fixture"""
import csv
import json
import os

from oracle import fitsio

REF = "/root/reference"
KEYS = ("NAXIS1", "NAXIS2", "CTYPE1", "CTYPE2", "CRPIX1", "CRPIX2", "CRVAL1", "CRVAL2", "PC1_1", "PC1_2", "PC2_1", "PC2_2",
        "CD1_1", "CD1_2", "CD2_1", "CD2_2", "CDELT1", "CDELT2", "LONPOLE", "LATPOLE")

if __name__ == "__main__":
    out = []
    with open(os.path.join(REF, "images.csv"), newline="") as fh:
        rows = csv.reader(fh, skipinitialspace=True)
        next(rows)
        for name, ra, dec in rows:
            hdr, _ = fitsio.read_header(open(os.path.join(REF, name.strip()), "rb").read())
            out.append(dict(file=name.strip(), ra=float(ra), dec=float(dec), header={k: hdr[k] for k in KEYS if k in hdr}))
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "sdss_wcs.json")
    json.dump(out, open(path, "w"), indent=0)
    print(len(out), "frames ->", path)
