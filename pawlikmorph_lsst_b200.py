"""Import shim: the package directory is named ``pawlikmorph-lsst_b200`` (not a valid
Python identifier), so this module loads it under the importable name
``pawlikmorph_lsst_b200``::

    import pawlikmorph_lsst_b200 as pm
    from pawlikmorph_lsst_b200 import pixmap, apertures, asymmetry, casgm
"""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pawlikmorph-lsst_b200")
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
