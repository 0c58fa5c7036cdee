"""CPU oracle for the PawlikMorph-LSST per-cutout morphology hot path.

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's CPU
baseline legs may import this package, and only as the checker.  The product
package (`pawlikmorph-lsst_b200/`) never imports it and has no CPU fallback.
"""
