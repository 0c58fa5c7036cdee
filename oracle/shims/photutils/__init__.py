"""Stand-in for photutils (not installed here): exact circular apertures + gini.
Test infrastructure; see oracle/leaves.py."""
__version__ = "0.0-oracle-shim"
from oracle.leaves import CircularAperture, CircularAnnulus  # noqa: F401
from . import morphology  # noqa: F401
