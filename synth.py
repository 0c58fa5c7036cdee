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
