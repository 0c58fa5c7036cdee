"""sersic.fitSersic of the reference (sersic.py:53-136, called at main.py:480-488) on the GPU.

The reference measures the flux inside the ellipse of semi-axes 2 max(fwhm), 2 min(fwhm) with photutils' exact aperture
photometry, walks the semi-major axis down in steps of a/50 until half of that flux is left, refines the half-light
semi-major axis with brentq, takes the mean flux of a one-pixel-wide elliptical annulus there as the amplitude, and fits
astropy's Sersic2D to the whole image with LevMarLSQFitter (scipy's leastsq).  Here the two pieces that touch pixels are
kernels — `pm_ellipse_sum_batch` (exact ellipse / pixel overlap) and `pm_sersicfit_batch` (the trust-region
Levenberg-Marquardt iteration of the Gaussian fit with the Sersic model and its analytic Jacobian) — and the scalar logic
between them runs on the host for the whole batch in lock-step: one launch per walk step and per root-search step,
whatever the number of cutouts.

Deviations: the half-light root is bracketed like upstream and then found by bisection to 1e-13 of the bracket
(the flux fraction is monotone in the semi-major axis, so brentq's root is the same number to its own xtol of 2e-12); the
fit runs to 1e-10 where leastsq stops at 1e-8, so parameters agree to leastsq's distance from the optimum (~1e-6 relative
on clean profiles), and a fit that upstream abandons after 1000 function evaluations is still iterated to the optimum or to
200 steps here.  A walk that never reaches half of the flux (non-positive total) loops forever upstream; here it ends with
status 4 and the -99 defaults."""
import numpy as np

from . import _abi, _lib

FIELDS = ("amplitude", "r_eff", "n", "x_0", "y_0", "ellip", "theta")


def _cuda_images(images):
    import torch
    if not torch.is_tensor(images) or not images.is_cuda:
        raise TypeError("images must be a CUDA tensor [B, n, n] (there is no CPU path)")
    if images.dim() != 3 or images.shape[1] != images.shape[2]:
        raise ValueError("images: [B, n, n] square cutouts")
    if images.dtype not in (torch.float32, torch.float64):
        images = images.to(torch.float64)
    return images.contiguous()


def ellipse_sum_batch(images, ell, idx=None, stream=None):
    """EllipticalAperture((xc, yc), a, b, theta).do_photometry(image, method="exact")[0][0] for M entries.
    images: CUDA [B, n, n]; ell: [M, 5] = xc, yc, a, b, theta (numpy or tensor); idx: [M] cutout index of every entry
    (None: entry i is cutout i).  Returns a CUDA float64 tensor [M]."""
    import torch
    images = _cuda_images(images)
    dev = images.device
    ell_d = torch.as_tensor(np.ascontiguousarray(np.asarray(ell, dtype=np.float64).reshape(-1, 5)), device=dev) if not torch.is_tensor(ell) \
        else ell.to(device=dev, dtype=torch.float64).reshape(-1, 5).contiguous()
    M = int(ell_d.shape[0])
    if idx is None:
        if M != int(images.shape[0]):
            raise ValueError("without idx there must be one entry per cutout")
        idx_d = None
    else:
        idx_d = torch.as_tensor(np.ascontiguousarray(np.asarray(idx, dtype=np.int64)), device=dev) if not torch.is_tensor(idx) \
            else idx.to(device=dev, dtype=torch.int64).contiguous()
        if int(idx_d.shape[0]) != M:
            raise ValueError("idx: one cutout index per entry")
        if M and (int(idx_d.min()) < 0 or int(idx_d.max()) >= int(images.shape[0])):
            raise IndexError("idx out of range")
    out = torch.empty(M, dtype=torch.float64, device=dev)
    L = _lib.lib()
    with torch.cuda.device(dev):
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        rc = L.pm_ellipse_sum_batch(images.data_ptr(), _abi.PM_F32 if images.dtype == torch.float32 else _abi.PM_F64, int(images.shape[1]),
                                    M, None if idx_d is None else idx_d.data_ptr(), ell_d.data_ptr(), out.data_ptr(), st.cuda_stream)
    _abi.check(L, rc, "pm_ellipse_sum_batch")
    return out


def sersicfit_batch(images, p0, stream=None):
    """The Sersic2D least-squares fit from the start values p0 [B, 7] (FIELDS order): CUDA float64 [B, 10] = the seven
    parameters, cost, iterations, status (0 converged, 1 iteration limit, 2 NaN, 3 singular)."""
    import torch
    images = _cuda_images(images)
    B = int(images.shape[0])
    p0_d = torch.as_tensor(np.ascontiguousarray(np.asarray(p0, dtype=np.float64).reshape(B, 7)), device=images.device)
    out = torch.empty((B, _abi.PM_GAUSS_COLS), dtype=torch.float64, device=images.device)
    L = _lib.lib()
    with torch.cuda.device(images.device):
        st = stream if stream is not None else torch.cuda.current_stream(images.device)
        rc = L.pm_sersicfit_batch(images.data_ptr(), _abi.PM_F32 if images.dtype == torch.float32 else _abi.PM_F64, B, int(images.shape[1]),
                                  p0_d.data_ptr(), out.data_ptr(), st.cuda_stream)
    _abi.check(L, rc, "pm_sersicfit_batch")
    return out


def start_values(centroids, fwhms, thetas, sums):
    """The start values of the fit (sersic.py:98-129) for B cutouts in lock-step: (p0 [B, 7], status [B]); status 0 ok, 4 no
    start values (non-positive axes, a total flux that is not positive, a walk that never reaches half of it, or an annulus
    photutils rejects).  sums(xc, yc, a, b, theta, which) -> the exact elliptical aperture sums of the cutouts `which`."""
    cen = np.asarray(centroids, dtype=np.float64).reshape(-1, 2)
    B = cen.shape[0]
    fw = np.asarray(fwhms, dtype=np.float64).reshape(B, 2)
    th = np.asarray(thetas, dtype=np.float64).reshape(B)
    b = 2.0 * fw.min(axis=1)
    a = 2.0 * fw.max(axis=1)
    status = np.zeros(B, dtype=np.int64)
    with np.errstate(all="ignore"):
        ok = np.isfinite(a) & np.isfinite(b) & (a > 0) & (b > 0) & np.isfinite(th) & np.isfinite(cen).all(axis=1)
        a_s, b_s = np.where(ok, a, 1.0), np.where(ok, b, 1.0)
        cen_s, th_s = np.where(ok[:, None], cen, 0.0), np.where(ok, th, 0.0)
        ellip = 1.0 - b_s / a_s

        def ap(semi_major, semi_minor, which):
            if which.size == 0:
                return np.zeros(0)
            return np.asarray(sums(cen_s[which, 0], cen_s[which, 1], semi_major, semi_minor, th_s[which], which), dtype=np.float64)

        total = ap(a_s, b_s, np.arange(B))
        ok &= np.isfinite(total) & (total > 0)
        # the walk down the semi-major axis (sersic.py:103-116)
        delta = (a_s / 100.0) * 2.0
        cur = a_s - delta
        a_min, a_max = np.full(B, np.nan), np.full(B, np.nan)
        active = ok.copy()
        for _ in range(400):
            if not active.any():
                break
            w = np.nonzero(active)[0]
            cur[w] = np.maximum(0.001, cur[w])
            frac = ap(cur[w], b_s[w], w) / total[w]
            hit = frac <= 0.5
            a_min[w[hit]] = cur[w[hit]]
            a_max[w[hit]] = cur[w[hit]] + delta[w[hit]]
            stuck = ~hit & (cur[w] <= 0.001)                          # upstream never leaves this loop
            ok[w[stuck]] = False
            active[w[hit | stuck]] = False
            cur[w[~hit]] -= delta[w[~hit]]
        ok &= ~active
        # the half-light semi-major axis: the flux fraction is monotone in it, bisect the bracket
        w = np.nonzero(ok)[0]
        lo, hi = a_min[w].copy(), a_max[w].copy()
        f_hi = ap(hi, b_s[w], w) / total[w] - 0.5
        bad_bracket = ~(f_hi >= 0)                                    # brentq would raise ValueError upstream
        for _ in range(60):
            todo = np.nonzero(hi - lo > 1e-13 * np.maximum(hi, 1.0))[0]   # per cutout: the result does not depend on the batch
            if todo.size == 0:
                break
            mid = 0.5 * (lo[todo] + hi[todo])
            up = ap(mid, b_s[w[todo]], w[todo]) / total[w[todo]] - 0.5 > 0
            hi[todo] = np.where(up, mid, hi[todo])
            lo[todo] = np.where(up, lo[todo], mid)
        r_eff = np.full(B, np.nan)
        r_eff[w] = 0.5 * (lo + hi)
        ok[w[bad_bracket]] = False
        # amplitude: mean flux of the annulus r_eff -/+ 0.5 (sersic.py:118-123; photutils' inner minor axis b_out a_in / a_out)
        w = np.nonzero(ok)[0]
        a_in, a_out = r_eff[w] - 0.5, r_eff[w] + 0.5
        b_out = a_out - ellip[w]
        b_in = b_out * a_in / a_out
        good = (a_in > 0) & (b_out > 0) & (b_in > 0)                  # photutils raises ValueError otherwise
        ok[w[~good]] = False
        w, a_in, a_out, b_out, b_in = w[good], a_in[good], a_out[good], b_out[good], b_in[good]
        flux = ap(a_out, b_out, w) - ap(a_in, b_in, w)
        amp = np.full(B, np.nan)
        amp[w] = flux / (np.pi * (a_out * b_out - a_in * b_in))
    p0 = np.stack([amp, r_eff, np.full(B, 2.0), cen[:, 0], cen[:, 1], ellip, th], axis=1)
    status[~ok] = 4
    p0[~ok] = np.nan
    return p0, status


def sersic_start_batch(images, centroids, fwhms, thetas, stream=None):
    """start_values with the sums taken on the GPU: one launch per walk step and per bisection step for the whole batch."""
    images = _cuda_images(images)

    def sums(xc, yc, a, b, theta, which):
        return ellipse_sum_batch(images, np.stack([xc, yc, a, b, theta], axis=1), idx=which, stream=stream).cpu().numpy()

    return start_values(centroids, fwhms, thetas, sums)


def fitSersic_batch(images, centroids, fwhms, thetas, star_masks=None, stream=None):
    """sersic.fitSersic for a batch of cutouts on the GPU.  images: CUDA [B, n, n]; centroids [B, 2] = (x, y) = (column, row)
    of `Result.apix`; fwhms [B, 2]; thetas [B]; star_masks: CUDA [B, n, n] multiplied into the images (sersic.py:88-91) or
    None.  Returns dict(params [B, 7] in FIELDS order — NaN where status != 0 —, start [B, 7], cost, iterations, status [B]:
    0 ok, 1 iteration limit (parameters of the last step are kept, like upstream's warning), 2 / 3 the fit broke down,
    4 no start values)."""
    images = _cuda_images(images)
    if star_masks is not None:
        images = images * star_masks.to(images.dtype)
    B = int(images.shape[0])
    p0, status = sersic_start_batch(images, centroids, fwhms, thetas, stream=stream)
    params = np.full((B, 7), np.nan)
    cost, iters = np.full(B, np.nan), np.zeros(B)
    w = np.nonzero(status == 0)[0]
    if w.size:
        import torch
        sel = images if w.size == B else images[torch.as_tensor(w, device=images.device)]
        fit = sersicfit_batch(sel, p0[w], stream=stream).cpu().numpy()
        st = fit[:, 9].astype(np.int64)
        keep = st <= 1
        params[w[keep]] = fit[keep, :7]
        cost[w], iters[w] = fit[:, 7], fit[:, 8]
        status[w] = st
    return dict(params=params, start=p0, cost=cost, iterations=iters, status=status)


class _Value:
    def __init__(self, value):
        self.value = float(value)

    def __float__(self):
        return self.value

    def __repr__(self):
        return f"Parameter(value={self.value})"


class SersicParameters:
    """What main.py:482-488 reads off the astropy model fitSersic returns: `.amplitude.value`, `.r_eff.value`, ..."""

    def __init__(self, values):
        for name, v in zip(FIELDS, values):
            setattr(self, name, _Value(v))
        self.parameters = np.asarray(values, dtype=np.float64)

    def __repr__(self):
        return "SersicParameters(" + ", ".join(f"{n}={getattr(self, n).value}" for n in FIELDS) + ")"


def fitSersic(image, centroid, fwhms, theta, starMask=None, device="cuda:0"):
    """sersic.fitSersic (sersic.py:53-136) for one image: an object with `.amplitude.value`, `.r_eff.value`, `.n.value`, `.x_0.value`,
    `.y_0.value`, `.ellip.value`, `.theta.value`.  Raises ValueError where upstream raises (photutils / brentq) or never returns."""
    import torch
    img = torch.as_tensor(np.ascontiguousarray(np.asarray(image, dtype=np.float64)), device=device)[None]
    mask = None if starMask is None else torch.as_tensor(np.ascontiguousarray(np.asarray(starMask, dtype=np.float64)), device=device)[None]
    r = fitSersic_batch(img, [list(centroid)], [list(fwhms)], [float(theta)], star_masks=mask)
    if r["status"][0] not in (0, 1):
        raise ValueError(f"fitSersic: no fit (status {int(r['status'][0])})")
    return SersicParameters(r["params"][0])
