"""Loads libpawlik_b200.so (the CUDA library) and binds the C ABI.  Fails loudly when the
library has not been built — there is no other compute path."""
import ctypes
import os

from . import _abi
from .build import SO

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            raise ImportError(
                f"{SO} is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
                "pawlikmorph-lsst_b200 has no CPU fallback.")
        L = _abi.declare(ctypes.CDLL(SO))
        if L.pm_abi_version() != _abi.PM_ABI_VERSION:
            raise ImportError("libpawlik_b200.so ABI version mismatch: rebuild")
        _lib = L
    return _lib


def launch_count():
    return int(lib().pm_launch_count())
