"""Cases shared by the CPU (emulator build) and GPU tests of the sky-region statistics."""
import numpy as np

import synth
from oracle import sky_oracle as SO

FIELDS = ("count", "mean", "median", "std", "clip_iterations", "clipped_count", "clipped_mean", "clipped_median", "clipped_std",
          "sky", "sky_err", "status")
EXACT = ("count", "clip_iterations", "clipped_count", "status")


def cases(n=64, B=6, seed=3, dtype=np.float32):
    """images, fwhms, theta: Sersic + noise cutouts (positively skewed sky pixels -> the clipped branch), a pure-noise frame
    with a negative outlier tail (mean <= median -> the plain branch), an odd pixel count and a tiny sky region."""
    rng = np.random.default_rng(seed)
    x = synth.make_batch(B, n, seed).numpy().astype(np.float64)
    x[1] = 100.0 + rng.normal(0, 5, (n, n))
    x[1][rng.integers(0, n, 40), rng.integers(0, n, 40)] -= 80.0                 # negative outliers: mean < median
    x[2] = np.rint(x[2])                                                           # ties in the median
    x[3] -= 100.0                                                                  # values of both signs
    x[3][5, 7], x[3][6, 9], x[3][n - 2, 3] = -0.0, 0.0, -0.0
    fw = np.array([[3.0, 4.5], [2.0, 2.0], [5.0, 3.0], [1.5, 1.0], [n * 0.3, n * 0.45], [2.5, 6.0]])[:B]
    th = np.array([0.3, 0.0, 1.2, -0.7, 0.9, 2.5])[:B]
    return x.astype(dtype), fw, th


def check(rows, images, fw, th, rtol=1e-9):
    for i in range(len(images)):
        want = SO.calc_sky(images[i].astype(np.float64), fw[i], th[i])
        got = dict(zip(FIELDS, rows[i]))
        for k in FIELDS:
            if k in EXACT:
                assert got[k] == want[k], (i, k, got[k], want[k])
            else:
                assert abs(got[k] - want[k]) <= rtol * abs(want[k]) + 1e-12, (i, k, got[k], want[k])
    return True
