"""TEST INFRASTRUCTURE: golden vectors of the Gaussian fit made by the reference's OWN gaussfitter.py.

    python oracle/make_prep_golden.py        (needs /root/reference; a minute)

`pawlikMorphLSST/gaussfitter.py` imports numpy only, so it is executed here unmodified (loaded by file path, outside its
package) exactly the way skyBackground._gauss2dfit drives it (skyBackground.py:306-331): start values from `moments`,
scipy.optimize.curve_fit on `twodgaussian` over the meshgrid.  Inputs (small float32 cutouts) and popt go to
tests/golden/gaussfit.npz."""
import importlib.util
import os
import sys

import numpy as np
from scipy import optimize

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
REF = "/root/reference/pawlikMorphLSST/gaussfitter.py"


def main():
    import synth
    from tests.golden_util import GoldenSet
    spec = importlib.util.spec_from_file_location("ref_gaussfitter", REF)
    G = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(G)
    imgs = [x for x in synth.make_exact(10, 96, 424242)] + [x for x in synth.make_batch(6, 128, 77).numpy()]
    g = GoldenSet("sdss_s")
    imgs += [g.image(i).astype(np.float32) for i in (0, 7, 21, 40)]
    out = {}
    for k, img in enumerate(imgs):
        d = img.astype(np.float64)
        n = d.shape[0]
        X, Y = np.meshgrid(np.linspace(0, n - 1, n), np.linspace(0, n - 1, n))
        p0 = G.moments(d)
        popt, _ = optimize.curve_fit(G.twodgaussian, (X, Y), d.ravel(), p0=p0)
        out[f"img_{k}"] = img
        out[f"p0_{k}"] = np.asarray(p0, dtype=np.float64)
        out[f"popt_{k}"] = popt
        print(k, n, np.round(popt, 4), flush=True)
    out["count"] = np.array(len(imgs))
    np.savez_compressed(os.path.join(REPO, "tests", "golden", "gaussfit.npz"), **out)


if __name__ == "__main__":
    main()
