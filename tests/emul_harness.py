"""TEST INFRASTRUCTURE: builds tests/emul/_build/libpawlik_emul.so (the kernel sources of the
product compiled by g++ against the fibre CTA emulator) and calls it through the same C ABI
with numpy (host) buffers.  Lets `pytest -m "not gpu"` exercise the kernel logic on a box
without a GPU.  Not a product path: the package never imports this."""
import ctypes as C
import importlib.util
import os
import subprocess

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "pawlikmorph-lsst_b200")
BUILD = os.path.join(REPO, "tests", "emul", "_build")
SO = os.path.join(BUILD, "libpawlik_emul.so")

_spec = importlib.util.spec_from_file_location("_pm_abi", os.path.join(PKG, "_abi.py"))
abi = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(abi)


def _sources():
    cs = os.path.join(PKG, "csrc")
    return [os.path.join(cs, f) for f in os.listdir(cs) if f.endswith((".cu", ".cuh"))] + [os.path.join(REPO, "tests", "emul", f)
                                                              for f in ("cuda_emul.h", "cuda_emul.cpp")] + \
        [os.path.join(REPO, "include", "pawlik_b200.h")]


def build(force=False):
    os.makedirs(BUILD, exist_ok=True)
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(s) for s in _sources()):
        return SO
    cmd = ["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-DPM_EMULATE",
           "-I" + os.path.join(REPO, "tests", "emul"), "-I" + os.path.join(REPO, "include"),
           "-x", "c++"] + [os.path.join(PKG, "csrc", u) for u in ("pm_capi.cu", "pm_k_mono.cu", "pm_k_front.cu", "pm_k_minapix.cu",
                                                                   "pm_k_tailcta.cu", "pm_tail.cu", "pm_k_prep.cu", "pm_k_stats.cu")] + [
           os.path.join(REPO, "tests", "emul", "cuda_emul.cpp"),
           "-o", SO]
    subprocess.run(cmd, check=True)
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = abi.declare(C.CDLL(build()))
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data


def analyse(images, sky, sky_err, filter_size=3, flags=abi.PM_DO_ALL, segmap_in=None, apermask_in=None,
            centroid_in=None, radius_in=None, r20_in=None, threshold_in=None, smooth_sky_in=None,
            want=("segmap", "sorted", "aper", "cand")):
    """images [B,n,n] float32/float64 numpy.  Returns dict(rows=[B,16], segmap=..., sorted_idx=..., ...)."""
    L = lib()
    images = np.ascontiguousarray(images)
    assert images.dtype in (np.float32, np.float64) and images.ndim == 3
    B, n, _ = images.shape
    dtype = abi.PM_F32 if images.dtype == np.float32 else abi.PM_F64
    f64 = lambda a: None if a is None else np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), (B,) + np.shape(a)[1:] if np.ndim(a) else (B,)))
    sky = f64(sky)
    sky_err = f64(sky_err)
    u8 = lambda a: None if a is None else np.ascontiguousarray(np.asarray(a).astype(np.uint8).reshape(B, n, n))
    keep = dict(seg=u8(segmap_in), ap=u8(apermask_in),
                cen=None if centroid_in is None else np.ascontiguousarray(np.asarray(centroid_in, dtype=np.float64).reshape(B, 2)),
                rad=f64(radius_in), r20=f64(r20_in), thr=f64(threshold_in), ssky=f64(smooth_sky_in))
    ov = abi.Overrides(_p(keep["seg"]), None, _p(keep["ap"]), _p(keep["cen"]), _p(keep["rad"]), _p(keep["r20"]),
                       _p(keep["thr"]), _p(keep["ssky"]))
    res = {}
    cap = 512
    if "segmap" in want:
        res["segmap"] = np.full((B, n, n), 255, np.uint8)
    if "sorted" in want:
        res["sorted_idx"] = np.full((B, n * n), -7, np.int32)
        res["n_positive"] = np.full((B,), -7, np.int32)
    if "aper" in want:
        res["apermask"] = np.full((B, n, n), 255, np.uint8)
    if "cand" in want:
        res["cand_a"] = np.full((B, cap), -7.0, np.float64)
    out = abi.Outputs(_p(res.get("segmap")), _p(res.get("sorted_idx")), _p(res.get("n_positive")),
                      _p(res.get("apermask")), _p(res.get("cand_a")), cap, None)
    rows = np.full((B, abi.ROW_DOUBLES), np.nan, np.float64)
    wsb = L.pm_workspace_bytes(B, n, dtype)
    ws = np.zeros(wsb // 8 + 2, np.float64)
    rc = L.pm_analyse_batch(images.ctypes.data, dtype, B, n, _p(sky), _p(sky_err), filter_size, flags,
                            C.byref(ov), C.byref(out), rows.ctypes.data, ws.ctypes.data, wsb, None)
    abi.check(L, rc, "pm_analyse_batch")
    res["rows"] = rows
    return res


def row_dict(rows, i):
    return dict(zip(abi.ROW_FIELDS, rows[i]))


def aperpixmap(n, rads, nsubpix=9, frac=0.1):
    L = lib()
    rads = np.ascontiguousarray(np.atleast_1d(np.asarray(rads, dtype=np.float64)))
    out = np.full((len(rads), n, n), 255, np.uint8)
    rc = L.pm_aperpixmap_batch(n, rads.ctypes.data, len(rads), nsubpix, frac, out.ctypes.data, None)
    abi.check(L, rc, "pm_aperpixmap_batch")
    return out


def fits_decode(payload, bitpix, count, bzero=0.0, bscale=1.0, dtype=np.float32):
    L = lib()
    raw = np.ascontiguousarray(np.frombuffer(payload, dtype=np.uint8))
    out = np.empty(count, dtype=dtype)
    rc = L.pm_fits_decode_batch(raw.ctypes.data, int(bitpix), float(bzero), float(bscale), int(count), out.ctypes.data,
                                abi.PM_F32 if dtype == np.float32 else abi.PM_F64, None)
    abi.check(L, rc, "pm_fits_decode_batch")
    return out


def fits_decode_window(payload, bitpix, B, H, W, origins, npix, bzero=0.0, bscale=1.0, dtype=np.float32):
    L = lib()
    raw = np.ascontiguousarray(np.frombuffer(payload, dtype=np.uint8))
    org = np.ascontiguousarray(np.asarray(origins, dtype=np.int32).reshape(B, 2))
    out = np.empty((B, npix, npix), dtype=dtype)
    rc = L.pm_fits_decode_window_batch(raw.ctypes.data, int(bitpix), float(bzero), float(bscale), int(B), int(H), int(W),
                                       org.ctypes.data, int(npix), out.ctypes.data,
                                       abi.PM_F32 if dtype == np.float32 else abi.PM_F64, None)
    abi.check(L, rc, "pm_fits_decode_window_batch")
    return out


def sky_stats(images, fwhms, theta):
    import math
    L = lib()
    x = np.ascontiguousarray(images)
    B, n = x.shape[0], x.shape[1]
    f = np.asarray(fwhms, dtype=np.float64).reshape(B, 2)
    ell = np.empty((B, 4))
    ell[:, 0], ell[:, 1] = 2.0 * f.min(axis=1), 2.0 * f.max(axis=1)
    ell[:, 2], ell[:, 3] = [math.cos(v) for v in theta], [math.sin(v) for v in theta]
    out = np.empty((B, abi.PM_SKY_COLS))
    rc = L.pm_sky_stats_batch(x.ctypes.data, abi.PM_F32 if x.dtype == np.float32 else abi.PM_F64, B, n, ell.ctypes.data,
                              out.ctypes.data, None)
    abi.check(L, rc, "pm_sky_stats_batch")
    return out


def clipped_stats(images, sigma=3.0, maxiters=5):
    """images [B,n,n] float32/float64 -> [B, 10] (CLIP_FIELDS); maxiters=None: until convergence."""
    L = lib()
    images = np.ascontiguousarray(images)
    B, n, _ = images.shape
    out = np.full((B, abi.PM_CLIP_COLS), np.nan)
    rc = L.pm_clipped_stats_batch(images.ctypes.data, abi.PM_F32 if images.dtype == np.float32 else abi.PM_F64, B, n, float(sigma),
                                  0 if maxiters is None else int(maxiters), out.ctypes.data, None)
    abi.check(L, rc, "pm_clipped_stats_batch")
    return out


def clipped_stats_both(images, sigma=3.0, maxiters=5):
    """-> ([B, 10] at maxiters, [B, 10] converged) from one run of the two-output entry point."""
    L = lib()
    images = np.ascontiguousarray(images)
    B, n, _ = images.shape
    first, conv = np.full((B, abi.PM_CLIP_COLS), np.nan), np.full((B, abi.PM_CLIP_COLS), np.nan)
    rc = L.pm_clipped_stats2_batch(images.ctypes.data, abi.PM_F32 if images.dtype == np.float32 else abi.PM_F64, B, n, float(sigma), int(maxiters),
                                   first.ctypes.data, conv.ctypes.data, None)
    abi.check(L, rc, "pm_clipped_stats2_batch")
    return first, conv


def gaussfit(images, p0=None):
    """-> [B, 10] (GAUSS_FIELDS); start values from gaussfitter.moments unless p0 [B, 7] is given."""
    L = lib()
    images = np.ascontiguousarray(images)
    B, n, _ = images.shape
    out = np.full((B, abi.PM_GAUSS_COLS), np.nan)
    med = mx = None
    if p0 is None:
        cs = clipped_stats(images, 3.0, 1)
        med, mx = np.ascontiguousarray(cs[:, 2]), np.ascontiguousarray(cs[:, 9])
    else:
        p0 = np.ascontiguousarray(np.asarray(p0, dtype=np.float64).reshape(B, 7))
    rc = L.pm_gaussfit_batch(images.ctypes.data, abi.PM_F32 if images.dtype == np.float32 else abi.PM_F64, B, n, _p(med), _p(mx), _p(p0),
                             out.ctypes.data, None)
    abi.check(L, rc, "pm_gaussfit_batch")
    return out


def ellipse_sums(images, ell, idx=None):
    """exact elliptical aperture sums: ell [M, 5] = xc, yc, a, b, theta; idx [M] cutout per entry (None: identity)."""
    L = lib()
    images = np.ascontiguousarray(images)
    _, n, _ = images.shape
    ell = np.ascontiguousarray(np.asarray(ell, dtype=np.float64).reshape(-1, 5))
    M = ell.shape[0]
    idx = None if idx is None else np.ascontiguousarray(np.asarray(idx, dtype=np.int64))
    out = np.full(M, np.nan)
    rc = L.pm_ellipse_sum_batch(images.ctypes.data, abi.PM_F32 if images.dtype == np.float32 else abi.PM_F64, n, M, _p(idx), ell.ctypes.data,
                                out.ctypes.data, None)
    abi.check(L, rc, "pm_ellipse_sum_batch")
    return out


def sersicfit(images, p0):
    """-> [B, 10] (SERSIC_FIELDS) from the start values p0 [B, 7]."""
    L = lib()
    images = np.ascontiguousarray(images)
    B, n, _ = images.shape
    p0 = np.ascontiguousarray(np.asarray(p0, dtype=np.float64).reshape(B, 7))
    out = np.full((B, abi.PM_GAUSS_COLS), np.nan)
    rc = L.pm_sersicfit_batch(images.ctypes.data, abi.PM_F32 if images.dtype == np.float32 else abi.PM_F64, B, n, p0.ctypes.data,
                              out.ctypes.data, None)
    abi.check(L, rc, "pm_sersicfit_batch")
    return out


def maskstars(images, seed=0, nsigma=1.5, npixels=8, clean=True):
    """imageutils.maskstarsSEG through the emulator build: dict(clean, labels, info)."""
    L = lib()
    images = np.ascontiguousarray(images)
    B, n, _ = images.shape
    c5, cn = clipped_stats(images, 3.0, 5), clipped_stats(images, 3.0, None)
    thr = np.ascontiguousarray(cn[:, 6] + nsigma * cn[:, 8])
    mean, std = np.ascontiguousarray(c5[:, 6]), np.ascontiguousarray(c5[:, 8])
    labels = np.full((B, n, n), -1, np.int32)
    out_img = np.full_like(images, np.nan) if clean else None
    info = np.full((B, abi.PM_STAR_COLS), np.nan)
    wsb = L.pm_maskstars_workspace_bytes(B, n)
    ws = np.zeros(wsb // 8 + 2, np.float64)
    rc = L.pm_maskstars_batch(images.ctypes.data, abi.PM_F32 if images.dtype == np.float32 else abi.PM_F64, B, n, thr.ctypes.data, npixels,
                              mean.ctypes.data, std.ctypes.data, seed, labels.ctypes.data, _p(out_img), info.ctypes.data, ws.ctypes.data, wsb, None)
    abi.check(L, rc, "pm_maskstars_batch")
    return dict(clean=out_img, labels=labels, info=info, threshold=thr, mean=mean, std=std)
