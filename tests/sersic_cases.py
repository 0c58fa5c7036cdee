"""Shared inputs of the Sersic tests (CPU build and GPU): synthetic Sersic2D profiles with noise, and the map of a fitted
parameter set to the canonical member of its family of equivalent optima."""
import numpy as np

from oracle import sersic_oracle as S


def cases(n=96, count=4, seed=3, dtype=np.float64):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:n, 0:n]
    imgs, truths = [], []
    for _ in range(count):
        t = np.array([rng.uniform(10, 60), rng.uniform(4, 10), rng.uniform(0.8, 3.5), n / 2 - 1 + rng.uniform(-2, 2), n / 2 + rng.uniform(-2, 2),
                      rng.uniform(0.05, 0.5), rng.uniform(0, 3)])
        imgs.append(S.sersic2d(x, y, *t) + rng.normal(0, 0.3, (n, n)))
        truths.append(t)
    imgs = np.stack(imgs).astype(dtype)
    cen = [[float(round(t[3])), float(round(t[4]))] for t in truths]
    fw = [[t[1] * 1.1, t[1] * (1 - t[5]) * 1.1] for t in truths]
    th = [float(t[6]) for t in truths]
    return imgs, np.array(truths), cen, fw, th


def canonical(p):
    """Sersic2D(r_eff, ellip, theta), Sersic2D(-r_eff, ellip, theta) and Sersic2D((1 - ellip) r_eff, 1 - 1 / (1 - ellip), theta + pi / 2)
    are the same surface, and theta only counts modulo pi: the member with r_eff > 0, ellip >= 0 and theta in [0, pi).  Which
    member a fit reports depends on its path."""
    p = np.array(p, dtype=np.float64)
    p[1] = abs(p[1])                     # only r_eff^2 enters the profile
    if p[5] < 0:
        p[1] = p[1] * (1.0 - p[5])
        p[5] = 1.0 - 1.0 / (1.0 - p[5])
        p[6] = p[6] + np.pi / 2
    p[6] = np.mod(p[6], np.pi)
    return p


def close(got, want, rtol=2e-4):
    g, w = canonical(got), canonical(want)
    dth = abs(g[6] - w[6])
    dth = min(dth, np.pi - dth)
    return bool(np.all(np.abs(g[:6] - w[:6]) <= rtol * np.maximum(np.abs(w[:6]), 1e-3)) and dth <= 5e-3 / max(w[5], 0.02))
