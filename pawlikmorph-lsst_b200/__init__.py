"""pawlikmorph-lsst_b200 — B200-native per-cutout morphology hot path of PawlikMorph-LSST.

Module names and function signatures mirror the reference's L1 modules
(`pawlikMorphLSST/{pixmap,apertures,asymmetry,casgm,result}.py`) so the package drops in for
that path; `batch` adds the batched / multi-GPU entry points,
`image` the FITS payload decode, `helpers` / `main` the file-list driver that writes `parameters.csv`.  All arithmetic runs in
hand-written sm_100a CUDA kernels behind the C ABI of `include/pawlik_b200.h`
(`libpawlik_b200.so`, built in-tree by `build.py`).  There is no CPU fallback: importing the
compute modules without the built library, or calling them without a CUDA device, raises.
"""
from . import _abi  # noqa: F401

__all__ = ["pixmap", "apertures", "asymmetry", "casgm", "result", "batch", "image", "helpers", "main", "skyBackground", "imageutils", "sersic", "objectMasker"]


def __getattr__(name):
    if name in __all__ or name in ("_lib", "build"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
