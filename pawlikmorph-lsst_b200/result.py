"""The output record of the hot path: the fields of `pawlikMorphLSST/result.py:7-76` this path
fills, with the reference's -99 sentinels, the fields the path leaves at their defaults (Gaussian / Sersic fit
results, star flags: result.py:40-42, :56-74) and the CSV row writer of result.py:78-92."""
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
    fwhms: List[float] = field(default_factory=lambda: [-99., -99.])    # result.py:40; filled by skybgr (out of scope)
    theta: float = -99.
    sersic_amplitude: float = -99.                                      # result.py:56-68; filled by calculateSersic (main.py:480-488)
    sersic_r_eff: float = -99.
    sersic_n: float = -99.
    sersic_x_0: float = -99.
    sersic_y_0: float = -99.
    sersic_ellip: float = -99.
    sersic_theta: float = -99.
    time: float = 0.
    star_flag: bool = False
    maskedPixelFraction: float = -99.
    status: int = 0
    n_candidates: int = 0

    def write(self, objectfile):
        """One row of parameters.csv through a csv.writer: the 28 columns of result.py:78-92, in the order of the
        header main.py:119-125 writes (`CSV_HEADER` below)."""
        objectfile.writerow([f"{self.file}", f"{self.apix}", f"{self.rmax}", f"{self.r20}", f"{self.r80}", f"{self.sky}",
                             f"{self.sky_err}", f"{self.C}", f"{self.A[0]}", f"{self.A[1]}", f"{self.S}", f"{self.As[0]}",
                             f"{self.As90[0]}", f"{self.gini}", f"{self.m20}", f"{self.fwhms}", f"{self.theta}",
                             f"{self.sersic_amplitude}", f"{self.sersic_r_eff}", f"{self.sersic_n}", f"{self.sersic_x_0}",
                             f"{self.sersic_y_0}", f"{self.sersic_ellip}", f"{self.sersic_theta}", f"{self.time}",
                             f"{self.star_flag}", f"{self.maskedPixelFraction}", f"{self.objectEdge}"])

    @classmethod
    def from_row(cls, row, file="", sky=-99., sky_err=99.):
        """row: mapping over ROW_FIELDS (BatchResult.row_dict) or a 16-vector."""
        if not isinstance(row, dict):
            row = dict(zip(ROW_FIELDS, [float(v) for v in row]))
        r = cls(file=file, sky=sky, sky_err=sky_err)
        if int(row["status"]) == 1:          # main.py:420-424: default Result
            r.status = 1
            r.apix = (-99., -99.)            # result.py:34: the default is a tuple, and prints as one
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


# main.py:119-125
CSV_HEADER = ["file", "apix", "r_max", "r20", "r80", "sky", "sky_err", "C", "A", "Abgr", "S", "As", "As90", "Gini index", "M20",
              "fwhms", "theta", "sersic_amplitude", "sersic_r_eff", "sersic_n", "sersic_x_0", "sersic_y_0", "sersic_ellip",
              "sersic_theta", "time", "star_flag", "Masked pixel fraction", "Object on image edge?"]
