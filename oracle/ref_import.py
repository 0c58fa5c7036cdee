"""Import the reference's OWN, unmodified hot-path modules in this container.

TEST INFRASTRUCTURE.  `/root/reference` exists only in the build container, so
this module is used by `oracle/make_golden.py` and by `-m "not gpu"` tests that
skip when the reference tree is absent; nothing that runs on the GPU box needs it.

The four L1 files (`pixmap.py`, `apertures.py`, `asymmetry.py`, `casgm.py`) are
executed as they lie under `/root/reference/pawlikMorphLSST/`; their missing
third-party leaves (scikit-image, photutils, astropy — not installed, no network)
are supplied by the stand-ins in `oracle/shims/`, scipy / numpy / numba are real.
"""
import importlib
import os
import sys
import warnings

REF_ROOT = os.environ.get("PAWLIK_REFERENCE", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "pawlikMorphLSST", "pixmap.py"))


def load():
    """Return a namespace with the reference modules pixmap, apertures, asymmetry, casgm."""
    if not available():
        raise ImportError(f"reference tree not found at {REF_ROOT}")
    repo = os.path.dirname(_HERE)
    if repo not in sys.path:
        sys.path.insert(0, repo)
    for name in ("skimage", "photutils", "astropy"):
        try:
            importlib.import_module(name)
        except ImportError:
            shim_dir = os.path.join(_HERE, "shims")
            if shim_dir not in sys.path:
                sys.path.append(shim_dir)
            importlib.import_module(name)
    if REF_ROOT not in sys.path:
        sys.path.append(REF_ROOT)
    warnings.filterwarnings("ignore", category=DeprecationWarning, module=r"pawlikMorphLSST\..*")
    mods = {}
    for m in ("apertures", "pixmap", "asymmetry", "casgm"):
        mods[m] = importlib.import_module("pawlikMorphLSST." + m)

    class _NS:
        pass
    ns = _NS()
    ns.__dict__.update(mods)
    ns.shimmed = [n for n in ("skimage", "photutils", "astropy")
                  if "oracle-shim" in getattr(sys.modules[n], "__version__", "")]
    return ns
