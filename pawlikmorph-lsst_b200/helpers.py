"""File-list helpers of the reference's driver (`pawlikMorphLSST/helpers.py`), without astropy."""
import csv
from pathlib import Path

__all__ = ["getFiles", "getLocation"]


def getFiles(file: str):
    """Generator over (filename, ra, dec) of a CSV such as the reference's `images.csv` (helpers.py:11-36, which reads
    it with astropy.table).  The header there is `filename, ra, dec` — with spaces after the commas (images.csv:1) —
    so names and fields are stripped; ra / dec come back as floats."""
    with open(file, newline="") as fh:
        rows = csv.reader(fh, skipinitialspace=True)
        header = [h.strip() for h in next(rows)]
        col = {name: header.index(name) for name in ("filename", "ra", "dec")}
        for row in rows:
            if not row or not "".join(row).strip():
                continue
            yield row[col["filename"]].strip(), float(row[col["ra"]]), float(row[col["dec"]])


def getLocation(folder):
    """Output folder as a Path, created when missing: `folder`, or ./output (helpers.py:70-94)."""
    outfolder = Path(folder) if folder else Path().cwd() / "output"
    if not outfolder.exists():
        outfolder.mkdir()
    return outfolder
