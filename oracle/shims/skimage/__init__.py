"""Stand-in for scikit-image (not installed here): only the three entry points the
reference hot path calls.  Test infrastructure; see oracle/leaves.py."""
__version__ = "0.0-oracle-shim"
from . import transform, measure  # noqa: F401
