"""Synthetic Sersic-plus-noise cutouts for tests, smoke() and bench.py.

Bench/test infrastructure: one torch generator that runs on CPU or on the GPU.
Per galaxy: Sersic index n_s~U(0.8,4), r_e~U(0.03,0.10)*n px, axis ratio q~U(0.4,1),
PA~U(0,pi), I_e~logU(10,100) counts, centre (n/2,n/2)+U(-1,1)^2; with p=0.3 an
exponential companion (r_e=0.02n, offset U(0.05,0.15)*n, flux ratio U(0.05,0.3));
plus sky 100.0 and N(0, 5^2) noise.  sky=100, sky_err=5 are what the caller
passes to the hot path (threshold 105).
"""
import torch

SKY, SKY_ERR = 100.0, 5.0


def make_batch(B, n, seed, device="cpu", dtype=torch.float32, chunk=4096):
    """Return float32 [B, n, n] cutouts (sky included)."""
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    out = torch.empty((B, n, n), dtype=dtype, device=device)
    yy, xx = torch.meshgrid(torch.arange(n, device=device, dtype=torch.float32),
                            torch.arange(n, device=device, dtype=torch.float32), indexing="ij")

    def U(k, lo, hi):
        return lo + (hi - lo) * torch.rand(k, generator=g, device=device)

    for s in range(0, B, chunk):
        k = min(chunk, B - s)
        ns = U(k, 0.8, 4.0)
        re = U(k, 0.03, 0.10) * n
        q = U(k, 0.4, 1.0)
        pa = U(k, 0.0, 3.141592653589793)
        ie = 10.0 ** U(k, 1.0, 2.0)
        cx = n / 2 + U(k, -1.0, 1.0)
        cy = n / 2 + U(k, -1.0, 1.0)
        has_c = (torch.rand(k, generator=g, device=device) < 0.3).float()
        off = U(k, 0.05, 0.15) * n
        offa = U(k, 0.0, 6.283185307179586)
        fr = U(k, 0.05, 0.3)
        bn = 2.0 * ns - 1.0 / 3.0 + 4.0 / (405.0 * ns)
        v = lambda t: t.view(k, 1, 1)
        dx, dy = xx - v(cx), yy - v(cy)
        c, sn = torch.cos(v(pa)), torch.sin(v(pa))
        xr = dx * c + dy * sn
        yr = (-dx * sn + dy * c) / v(q)
        r = torch.sqrt(xr * xr + yr * yr)
        img = v(ie) * torch.exp(-v(bn) * ((r / v(re)).clamp_min(1e-6) ** (1.0 / v(ns)) - 1.0))
        # companion: exponential disc carrying `fr` of the main flux (2 pi n_s-free approximation)
        rc = 0.02 * n
        dxc = dx - v(off * torch.cos(offa))
        dyc = dy - v(off * torch.sin(offa))
        main_flux = img.sum(dim=(1, 2), keepdim=True)
        comp = torch.exp(-torch.sqrt(dxc * dxc + dyc * dyc) / rc)
        comp = comp * (v(fr * has_c) * main_flux / comp.sum(dim=(1, 2), keepdim=True))
        noise = torch.randn((k, n, n), generator=g, device=device) * SKY_ERR
        out[s:s + k] = (img + comp + SKY + noise).to(dtype)
    return out


def make_exact(B, n, seed):
    """float32 [B, n, n] numpy cutouts that are BIT-IDENTICAL on every machine: integer random numbers (PCG64 raw
    integers) and IEEE-exact operations only (+, -, *, /, sqrt; no exp / pow / sin whose last bit depends on the
    maths library).  Used where oracle rows are committed and the images are re-made from the seed on the GPU box
    (tests/golden/seeded_*.npz).  Per galaxy: a Moffat profile I0 / (1 + (r/a)^2)^beta, beta in {1, 1.5, 2, 2.5},
    elliptical (axis ratio 0.4..1, rational rotation), centre (n/2, n/2) + U(-1, 1)^2, with p = 0.3 a round companion;
    sky 100 + noise of sigma 5 (sum of four uniforms)."""
    import numpy as np
    rng = np.random.Generator(np.random.PCG64(int(seed)))

    def U(k, lo, hi):               # exact: 24 random bits scaled
        return lo + (hi - lo) * (rng.integers(0, 1 << 24, size=k).astype(np.float64) / float(1 << 24))

    yy, xx = np.meshgrid(np.arange(n, dtype=np.float64), np.arange(n, dtype=np.float64), indexing="ij")
    out = np.empty((B, n, n), dtype=np.float32)
    for b in range(B):
        beta2 = int(rng.integers(2, 6))                       # 2 beta
        a = float(U(1, 0.02, 0.07)[0]) * n
        q = float(U(1, 0.4, 1.0)[0])
        t = float(U(1, -1.0, 1.0)[0])
        c, s = (1.0 - t * t) / (1.0 + t * t), 2.0 * t / (1.0 + t * t)
        i0 = float(U(1, 60.0, 1500.0)[0])
        cx, cy = n / 2 + float(U(1, -1.0, 1.0)[0]), n / 2 + float(U(1, -1.0, 1.0)[0])
        dx, dy = xx - cx, yy - cy
        xr = dx * c + dy * s
        yr = (dy * c - dx * s) / q
        u = 1.0 + (xr * xr + yr * yr) / (a * a)

        def powm(v, twice_beta):      # v ** (-twice_beta / 2) with exact operations
            w = np.ones_like(v)
            for _ in range(twice_beta // 2):
                w = w * v
            if twice_beta & 1:
                w = w * np.sqrt(v)
            return 1.0 / w

        img = i0 * powm(u, beta2)
        if int(rng.integers(0, 10)) < 3:
            off, t2 = float(U(1, 0.05, 0.15)[0]) * n, float(U(1, -1.0, 1.0)[0])
            ox, oy = off * (1.0 - t2 * t2) / (1.0 + t2 * t2), off * 2.0 * t2 / (1.0 + t2 * t2)
            uc = 1.0 + ((dx - ox) ** 2 + (dy - oy) ** 2) / ((0.02 * n) ** 2)
            img = img + float(U(1, 0.05, 0.5)[0]) * i0 * powm(uc, 4)
        noise = rng.integers(0, 1 << 16, size=(4, n, n)).astype(np.float64).sum(axis=0) / float(1 << 16) - 2.0   # var 1/3
        out[b] = (img + SKY + noise * (SKY_ERR * 1.7320508075688772)).astype(np.float32)
    return out
