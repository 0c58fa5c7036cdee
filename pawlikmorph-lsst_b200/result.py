"""The output record of the hot path: the fields of `pawlikMorphLSST/result.py:7-76` this path
fills, with the reference's -99 sentinels.  (CSV writing, result.py:78-92, is out of scope.)"""
from dataclasses import dataclass, field
from typing import List

from ._abi import ROW_FIELDS


@dataclass
class Result:
    file: str = ""
    outfolder: str = ""
    occludedFile: str = ""
    pixelMapFile: str = ""
    cleanImage: str = ""
    A: List[float] = field(default_factory=lambda: [-99., -99.])        # result.py:26
    As: List[float] = field(default_factory=lambda: [-99., -99.])
    As90: List[float] = field(default_factory=lambda: [-99., -99.])
    rmax: float = -99
    apix: List[float] = field(default_factory=lambda: [-99., -99.])
    sky: float = -99.
    sky_err: float = 99.
    C: float = -99.
    r20: float = -99.
    r80: float = -99.
    S: float = -99.
    gini: float = -99.
    m20: float = -99.
    objectEdge: bool = False
    status: int = 0
    n_candidates: int = 0

    @classmethod
    def from_row(cls, row, file="", sky=-99., sky_err=99.):
        """row: mapping over ROW_FIELDS (BatchResult.row_dict) or a 16-vector."""
        if not isinstance(row, dict):
            row = dict(zip(ROW_FIELDS, [float(v) for v in row]))
        r = cls(file=file, sky=sky, sky_err=sky_err)
        if int(row["status"]) == 1:          # main.py:420-424: default Result
            r.status = 1
            return r
        r.A = [row["A"], row["Abgr"]]
        r.As = [row["As"], 0 if row["As"] != -99. else -99.]     # calcA(..., noisecorrect=False)[1] == 0
        r.As90 = [row["As90"], 0 if row["As90"] != -99. else -99.]
        r.rmax = row["rmax"]
        r.apix = [int(row["apix_col"]), int(row["apix_row"])] if row["apix_col"] != -99. else [-99., -99.]
        r.C, r.r20, r.r80, r.S = row["C"], row["r20"], row["r80"], row["S"]
        r.gini, r.m20 = row["gini"], row["m20"]
        r.objectEdge = bool(row["object_edge"])
        r.status = int(row["status"])
        r.n_candidates = int(row["n_candidates"])
        return r
