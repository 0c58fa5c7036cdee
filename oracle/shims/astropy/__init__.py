"""Stand-in for astropy: `apertures.py:5` imports astropy.io.fits at module scope
(used only by the offline `makeaperpixmaps`).  Test infrastructure."""
__version__ = "0.0-oracle-shim"
