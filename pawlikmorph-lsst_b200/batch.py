"""Batched entry points: a whole batch of cutouts through the fused per-galaxy kernel.

This is what replaces the reference's per-image driver loop (`main.py:136-170`,
`_analyseImage` `main.py:218-501`, steps `:421-495`): one call per batch instead of one Python
call chain per galaxy.  PyTorch is used for device memory, streams and `torch.distributed`
only; the arithmetic is in libpawlik_b200.so.
"""
import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _abi
from ._abi import (PM_DO_A, PM_DO_ALL, PM_DO_AS, PM_DO_CAS, PM_DO_S, ROW_DOUBLES, ROW_FIELDS)  # noqa: F401
from ._lib import lib

_ws_cache = {}


def _workspace(device, nbytes, stream):
    # one workspace per (device, stream): launches on different streams may run concurrently and each
    # needs its own work counter and per-CTA scratch
    key = (device.type, device.index, int(stream.cuda_stream))
    t = _ws_cache.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _ws_cache[key] = t
    return t


def _per_galaxy(v, B, device, name):
    if v is None:
        return None
    if not torch.is_tensor(v):
        v = torch.as_tensor(np.asarray(v, dtype=np.float64))
    v = v.to(device=device, dtype=torch.float64)
    if v.ndim == 0:
        v = v.expand(B)
    v = v.contiguous()
    if v.shape[0] != B:
        raise ValueError(f"{name}: expected {B} values, got {tuple(v.shape)}")
    return v


def _mask(v, B, n, device, name):
    if v is None:
        return None
    if not torch.is_tensor(v):
        v = torch.as_tensor(np.asarray(v))
    v = (v.to(device) != 0).to(torch.uint8).reshape(B, n, n).contiguous()
    return v


def _ptr(t):
    return None if t is None else t.data_ptr()


@dataclass
class BatchResult:
    rows: torch.Tensor                      # [B, 16] float64, columns = ROW_FIELDS
    segmap: Optional[torch.Tensor] = None   # [B, n, n] uint8
    sorted_idx: Optional[torch.Tensor] = None
    n_positive: Optional[torch.Tensor] = None
    apermask: Optional[torch.Tensor] = None
    cand_a: Optional[torch.Tensor] = None

    def column(self, name):
        return self.rows[:, ROW_FIELDS.index(name)]

    def row_dict(self, i):
        r = self.rows[i].tolist()
        return dict(zip(ROW_FIELDS, r))


def analyse_batch(images, sky, sky_err, filter_size=3, flags=PM_DO_ALL, *, segmap_in=None, apermask_in=None,
                  centroid_in=None, radius_in=None, r20_in=None, threshold_in=None, smooth_sky_in=None,
                  want_segmap=False, want_sorted=False, want_apermask=False, cand_cap=0, rows_out=None,
                  segmap_out=None, stream=None, phase_cycles_out=None):
    """Run the hot path of main.py:421-495 on `images` [B, n, n] (CUDA tensor, float32 or float64,
    sky still in).  Asynchronous on `stream` (default: the current torch stream).  Returns a
    BatchResult whose tensors live on the same device."""
    if not torch.is_tensor(images) or not images.is_cuda:
        raise TypeError("analyse_batch needs a CUDA tensor (there is no CPU path)")
    if images.ndim != 3 or images.shape[1] != images.shape[2]:
        raise ValueError("images must be [B, n, n]")
    if images.dtype not in (torch.float32, torch.float64):
        raise TypeError("images must be float32 or float64")
    images = images.contiguous()
    dev = images.device
    B, n = int(images.shape[0]), int(images.shape[1])
    dtype = _abi.PM_F32 if images.dtype == torch.float32 else _abi.PM_F64
    L = lib()
    with torch.cuda.device(dev):
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        with torch.cuda.stream(st):
            sky_t = _per_galaxy(sky, B, dev, "sky")
            err_t = _per_galaxy(sky_err, B, dev, "sky_err")
            keep = dict(seg=_mask(segmap_in, B, n, dev, "segmap_in"), ap=_mask(apermask_in, B, n, dev, "apermask_in"),
                        rad=_per_galaxy(radius_in, B, dev, "radius_in"), r20=_per_galaxy(r20_in, B, dev, "r20_in"),
                        thr=_per_galaxy(threshold_in, B, dev, "threshold_in"),
                        ssky=_per_galaxy(smooth_sky_in, B, dev, "smooth_sky_in"), cen=None)
            if centroid_in is not None:
                cen = centroid_in if torch.is_tensor(centroid_in) else torch.as_tensor(np.asarray(centroid_in, dtype=np.float64))
                keep["cen"] = cen.to(device=dev, dtype=torch.float64).reshape(B, 2).contiguous()
            ov = _abi.Overrides(_ptr(keep["seg"]), None, _ptr(keep["ap"]), _ptr(keep["cen"]), _ptr(keep["rad"]),
                                _ptr(keep["r20"]), _ptr(keep["thr"]), _ptr(keep["ssky"]))
            res = BatchResult(rows=rows_out if rows_out is not None else torch.empty((B, ROW_DOUBLES), dtype=torch.float64, device=dev))
            if res.rows.shape != (B, ROW_DOUBLES) or res.rows.dtype != torch.float64 or not res.rows.is_contiguous():
                raise ValueError("rows_out must be a contiguous [B, 16] float64 tensor")
            if segmap_out is not None:
                if segmap_out.shape != (B, n, n) or segmap_out.dtype != torch.uint8 or not segmap_out.is_contiguous():
                    raise ValueError("segmap_out must be a contiguous [B, n, n] uint8 tensor")
                res.segmap = segmap_out
            elif want_segmap:
                res.segmap = torch.empty((B, n, n), dtype=torch.uint8, device=dev)
            if want_sorted:
                res.sorted_idx = torch.empty((B, n * n), dtype=torch.int32, device=dev)
                res.n_positive = torch.empty((B,), dtype=torch.int32, device=dev)
            if want_apermask:
                res.apermask = torch.empty((B, n, n), dtype=torch.uint8, device=dev)
            if cand_cap:
                res.cand_a = torch.empty((B, int(cand_cap)), dtype=torch.float64, device=dev)
            out = _abi.Outputs(_ptr(res.segmap), _ptr(res.sorted_idx), _ptr(res.n_positive), _ptr(res.apermask),
                               _ptr(res.cand_a), int(cand_cap), _ptr(phase_cycles_out))
            if B == 0:
                return res
            wsb = L.pm_workspace_bytes(B, n, dtype)
            ws = _workspace(dev, wsb, st)
            rc = L.pm_analyse_batch(images.data_ptr(), dtype, B, n, _ptr(sky_t), _ptr(err_t), int(filter_size),
                                    int(flags), C.byref(ov), C.byref(out), res.rows.data_ptr(), ws.data_ptr(), wsb,
                                    C.c_void_p(st.cuda_stream))
            _abi.check(L, rc, "pm_analyse_batch")
            # keep the temporaries alive until the stream has consumed them
            for t in list(keep.values()) + [sky_t, err_t, images, ws]:
                if t is not None:
                    t.record_stream(st)
    return res


def _rotated_masks(star, angle, cx, cy):
    """Batched starMask * rotate(starMask, angle, center=(cx, cy), cval=1) (asymmetry.py:186-191) for integer centres and
    the angles of the path: an index permutation with 1 where the source falls outside the frame.  star: [B, n, n] bool / u8;
    cx, cy: int64 [B].  Device tensor ops (0/1 arithmetic, exact)."""
    B, n = star.shape[0], star.shape[1]
    dev = star.device
    r = torch.arange(n, device=dev)[None, :, None]
    c = torch.arange(n, device=dev)[None, None, :]
    cxv, cyv = cx[:, None, None], cy[:, None, None]
    if angle == 180:
        sr, sc = 2 * cyv - r, 2 * cxv - c
    else:
        sr, sc = cyv + (c - cxv), cxv - (r - cyv)
    sr, sc = sr.expand(B, n, n), sc.expand(B, n, n)
    ok = (sr >= 0) & (sr < n) & (sc >= 0) & (sc < n)
    flat = star.reshape(B, -1).to(torch.uint8)
    idx = (sr.clamp(0, n - 1) * n + sc.clamp(0, n - 1)).reshape(B, -1)
    rot = torch.where(ok.reshape(B, -1), torch.gather(flat, 1, idx), torch.ones_like(flat)).reshape(B, n, n)
    return (star.to(torch.uint8) * rot)


def analyse_batch_starmask(images, sky, sky_err, starmask, filter_size=3, flags=PM_DO_ALL, stream=None, cleaned=None):
    """The path with a star mask per cutout (1 = keep) — the `-starmask` flow of main.py:395-413 and what follows it:
    pixelmap(img, thr, fs, starMask) (pixmap.py:156-158), calcRmax without the mask, minapix on image * segmap * starMask
    (asymmetry.py:82-86), calcA with S = starMask * rotate(starMask) about apix, the segmap multiplied by S in place
    (:186-196, :219) and used like that by As, As90 (with the 90 degree S), r20 / r80, Gini and M20 (main.py:470-495).
    A composition of launches of the fused kernel with its overrides (the mask changes between the steps once apix is
    known), for the rare catalogue / star-mask runs; `pm_analyse_batch` itself rejects `starmask`.  With noisecorrect the
    reference's result depends on ~1e-14 leakage of skimage's bilinear rotation of the mask; the exact 0/1 mask is used.
    cleaned: CUDA float64 [B, n, n], the sky-subtracted image after main.py:442-445 (`img -= sky`, maskstarsSEG) — everything
    after the pixelmap then runs on it, as in the catalogue flow of main.py:372-395; None: img - sky itself.
    Returns a BatchResult (rows + the pixelmap)."""
    if not torch.is_tensor(images) or not images.is_cuda:
        raise TypeError("analyse_batch_starmask needs a CUDA tensor (there is no CPU path)")
    dev = images.device
    B, n = int(images.shape[0]), int(images.shape[1])
    star = _mask(starmask, B, n, dev, "starmask")
    sky_t = _per_galaxy(sky, B, dev, "sky")
    err_t = _per_galaxy(sky_err, B, dev, "sky_err")
    col = ROW_FIELDS.index
    kw = dict(stream=stream)
    # 1. pixelmap of image * starMask (sky still in)
    p1 = analyse_batch(images * star.to(images.dtype), sky_t, err_t, filter_size, flags | _abi.PM_STOP_AFTER_PIXELMAP, want_segmap=True, **kw)
    rows = p1.rows.clone()
    seg = p1.segmap
    ok = rows[:, col("status")] == 0
    zero = torch.zeros_like(sky_t)
    base, base_sky = (images, sky_t) if cleaned is None else (cleaned, zero)
    # 2. rmax from the unmasked image and the map (pixmap.calcRmax, main.py:463)
    p2 = analyse_batch(base, base_sky, err_t, filter_size, _abi.PM_STOP_AFTER_RMAX, segmap_in=seg, **kw)
    rmax = p2.rows[:, col("rmax")].contiguous()
    rows[:, col("rmax")] = torch.where(ok, rmax, rows[:, col("rmax")])
    status2 = p2.rows[:, col("status")]
    # 3. minapix: image * segmap * starMask == the map with the masked pixels taken out
    p3 = analyse_batch(base, base_sky, err_t, filter_size, 0, segmap_in=seg * star, radius_in=torch.where(rmax > 0, rmax, torch.ones_like(rmax)), **kw)
    for name in ("apix_col", "apix_row", "n_candidates"):
        rows[:, col(name)] = torch.where(ok, p3.rows[:, col(name)], rows[:, col(name)])
    status3 = p3.rows[:, col("status")]
    cx = p3.rows[:, col("apix_col")].clamp(min=0).to(torch.int64)
    cy = p3.rows[:, col("apix_row")].clamp(min=0).to(torch.int64)
    cen = torch.stack([cx, cy], dim=1).to(torch.float64)
    good = ok & (status2 == 0) & (status3 == 0)
    # 4. A, Abgr, As, r20, r80, C, Gini, M20 on I = (img - sky) * S180 with the map * S180
    s180 = _rotated_masks(star, 180, cx, cy)
    sub = images.to(torch.float64) - sky_t[:, None, None] if cleaned is None else cleaned
    seg180 = seg * s180
    f4 = flags & (_abi.PM_DO_A | _abi.PM_DO_AS | _abi.PM_DO_CAS)
    if f4:
        p4 = analyse_batch(sub * s180.to(torch.float64), zero, err_t, filter_size, f4, segmap_in=seg180, centroid_in=cen, radius_in=rmax, **kw)
        for name in ("A", "Abgr", "As", "r20", "r80", "C", "gini", "m20"):
            rows[:, col(name)] = torch.where(good, p4.rows[:, col(name)], rows[:, col(name)])
        status4 = p4.rows[:, col("status")]
    else:
        status4 = torch.zeros_like(status3)
    # 5. As90 with S90 = starMask * rotate(starMask, 90) on the (already multiplied) map
    if flags & _abi.PM_DO_AS:
        s90 = _rotated_masks(star, 90, cx, cy)
        p5 = analyse_batch(sub, zero, err_t, filter_size, _abi.PM_DO_AS, segmap_in=seg180 * s90, centroid_in=cen, radius_in=rmax, **kw)
        rows[:, col("As90")] = torch.where(good, p5.rows[:, col("As90")], rows[:, col("As90")])
    # 6. smoothness on the unmasked image with the multiplied map (casgm.smoothness takes img, segmap: main.py:494)
    if (flags & _abi.PM_DO_S) and (flags & _abi.PM_DO_CAS):
        r20 = rows[:, col("r20")].contiguous()
        p6 = analyse_batch(sub, zero, err_t, filter_size, _abi.PM_DO_S, segmap_in=seg180, centroid_in=cen, radius_in=rmax,
                           r20_in=torch.where(r20 > 0, r20, torch.ones_like(r20)), smooth_sky_in=sky_t, **kw)
        rows[:, col("S")] = torch.where(good & (status4 == 0) & (r20 > 0), p6.rows[:, col("S")], rows[:, col("S")])
    st = rows[:, col("status")]
    st = torch.where(ok & (status2 != 0), status2, st)
    st = torch.where(ok & (status2 == 0) & (status3 != 0), status3, st)
    st = torch.where(good & (status4 != 0), status4, st)
    rows[:, col("status")] = st
    return BatchResult(rows=rows, segmap=seg)


_host_cache = {}


def clear_cache():
    """Release the workspaces, staging buffers, streams and pinned rows kept between calls."""
    _ws_cache.clear()
    _host_cache.clear()
    torch.cuda.empty_cache()


def _staging(dev, row_bytes_key, chunk, row_shape, dtype, B):
    """Reusable buffers of the host calls: two device staging tensors and two streams (chunks alternate), device rows and a
    pinned host copy of them.  One entry per (device, row shape, dtype); the buffers GROW to the largest chunk / batch seen
    (a caller looping over groups of varying size does not leak), `clear_cache()` releases them."""
    key = (dev.type, dev.index, row_bytes_key, dtype)
    h = _host_cache.get(key)
    if h is None:
        h = dict(bufs=None, streams=[torch.cuda.Stream(dev), torch.cuda.Stream(dev)], rows_d=None, rows_h=None, pin=None)
        _host_cache[key] = h
    if h["bufs"] is None or h["bufs"][0].shape[0] < chunk:
        h["bufs"] = [torch.empty((chunk,) + tuple(row_shape), dtype=dtype, device=dev) for _ in range(2)]
    if h["rows_d"] is None or h["rows_d"].shape[0] < B:
        h["rows_d"] = torch.empty((B, ROW_DOUBLES), dtype=torch.float64, device=dev)
        h["rows_h"] = torch.empty((B, ROW_DOUBLES), dtype=torch.float64).pin_memory()
    return h


def _pinned_chunks(h, x, s, e):
    """x[s:e] as a pinned tensor: itself when the input is pinned, else staged through a pinned buffer (two, alternating)
    so that the copy to the device stays asynchronous."""
    if x.is_pinned():
        return x[s:e], None
    n = e - s
    if h["pin"] is None or h["pin"][0].shape[0] < n or h["pin"][0].shape[1:] != x.shape[1:] or h["pin"][0].dtype != x.dtype:
        h["pin"] = [torch.empty((n,) + tuple(x.shape[1:]), dtype=x.dtype).pin_memory() for _ in range(2)]
        h["pin_ev"] = [None, None]
    return None, n


def analyse_host(images, sky, sky_err, filter_size=3, flags=PM_DO_ALL, device="cuda:0", chunk=8192, **kw):
    """The call a user of the reference's driver makes: HOST arrays in, HOST rows out.

    `images` is a numpy array / CPU tensor [B, n, n]; chunks alternate between two streams, each with its own device staging
    buffer, so that the host->device copy of chunk i+1 overlaps the kernel of chunk i.  Pinned input is copied straight
    from where it is; pageable input is staged through two pinned buffers.  Returns a [B, 16] float64 numpy array
    (columns = ROW_FIELDS) that the caller owns."""
    dev = torch.device(device)
    x = images if torch.is_tensor(images) else torch.from_numpy(np.ascontiguousarray(images))
    if x.dtype not in (torch.float32, torch.float64):
        x = x.to(torch.float64) if x.dtype in (torch.int64, torch.uint64) else x.to(torch.float32)
    B, n = int(x.shape[0]), int(x.shape[1])
    if B == 0:
        return np.empty((0, ROW_DOUBLES))
    chunk = max(1, min(int(chunk), B))
    h = _staging(dev, ("img", n), chunk, (n, n), x.dtype, B)
    cur = torch.cuda.current_stream(dev)
    sky_d = _per_galaxy(sky, B, dev, "sky")
    err_d = _per_galaxy(sky_err, B, dev, "sky_err")
    ready = torch.cuda.Event()
    ready.record(cur)
    rows_d = h["rows_d"]
    for i, s in enumerate(range(0, B, chunk)):
        e = min(B, s + chunk)
        st = h["streams"][i & 1]
        if i < 2:
            st.wait_event(ready)           # the sky vectors (and whatever the caller queued before)
        src, m = _pinned_chunks(h, x, s, e)
        if src is None:                    # pageable input: through the pinned buffer of this stream
            if h["pin_ev"][i & 1] is not None:
                h["pin_ev"][i & 1].synchronize()
            src = h["pin"][i & 1][:m]
            src.copy_(x[s:e])
        with torch.cuda.stream(st):
            d = h["bufs"][i & 1][:e - s]
            d.copy_(src, non_blocking=True)
            if not x.is_pinned():
                ev = torch.cuda.Event()
                ev.record(st)
                h["pin_ev"][i & 1] = ev
            analyse_batch(d, sky_d[s:e], err_d[s:e], filter_size, flags, rows_out=rows_d[s:e], stream=st, **kw)
    for st in h["streams"]:
        cur.wait_stream(st)
    out = torch.empty((B, ROW_DOUBLES), dtype=torch.float64)
    rows_h = h["rows_h"][:B]
    rows_h.copy_(rows_d[:B], non_blocking=True)
    cur.synchronize()
    out.copy_(rows_h)
    return out.numpy()


def payload_dtype(bitpix, bzero=0.0, bscale=1.0):
    """Decode type that loses nothing: float32 holds BITPIX 8 / 16 pixels with integer scaling and BITPIX -32 exactly;
    BITPIX 32 / -64 and fractional BSCALE / BZERO need float64 (what the reference analyses, Image.py:148)."""
    exact = bitpix == -32 or (bitpix in (8, 16) and float(bzero).is_integer() and float(bscale).is_integer()
                              and abs(bscale) * 65536 + abs(bzero) < 2 ** 24)
    return torch.float32 if exact else torch.float64


def analyse_host_payload(payload, bitpix, n, sky, sky_err, bzero=0.0, bscale=1.0, filter_size=3, flags=PM_DO_ALL,
                         device="cuda:0", chunk=8192, frame=None, origins=None, gauss=None, sky_rows_out=None, dtype=None, **kw):
    """Like analyse_host, but from the big-endian FITS pixel blocks of B cutouts of n x n pixels (`payload`: uint8
    numpy array / CPU tensor [B, n * n * bytes_per_pixel], pinned memory makes the copies asynchronous).  The bytes are
    copied as they are and decoded on the device (image.decode_payload): 2 bytes per pixel on the wire for BITPIX 16.
    `dtype`: the decoded type, by default the smallest that is exact for (bitpix, bzero, bscale) (`payload_dtype`).
    With `frame=(H, W)` and `origins` (int32 [B, 2]: first row, first column) every payload row is a whole H x W frame
    and the n x n window at its origin is analysed (image.decode_window: the cut-out of Image.py:175-186 on the device).
    With `gauss=(fwhms [B, 2], theta [B])` — the Gaussian fit of skyBackground.py:167-262 — `sky` / `sky_err` are not
    read: they are measured on the device from the sky region of every decoded cutout (skyBackground.sky_stats_batch)
    and feed the analysis without leaving the GPU; `sky_rows_out` (numpy [B, 12]) receives the statistics."""
    from .image import BYTES_PER_PIXEL, decode_payload, decode_window
    dev = torch.device(device)
    x = payload if torch.is_tensor(payload) else torch.from_numpy(np.ascontiguousarray(payload))
    H, W = (int(frame[0]), int(frame[1])) if frame is not None else (n, n)
    if x.dtype != torch.uint8 or x.ndim != 2 or x.shape[1] != H * W * BYTES_PER_PIXEL[bitpix]:
        raise ValueError("payload must be uint8 [B, rows * columns * bytes_per_pixel]")
    if frame is not None:
        origins = torch.as_tensor(origins, dtype=torch.int32).reshape(-1, 2)
        if origins.shape[0] != x.shape[0]:
            raise ValueError("one window origin per frame")
    B = int(x.shape[0])
    if B == 0:
        return np.empty((0, ROW_DOUBLES))
    dt = dtype if dtype is not None else payload_dtype(bitpix, bzero, bscale)
    chunk = max(1, min(int(chunk), B))
    h = _staging(dev, ("payload", int(x.shape[1])), chunk, (int(x.shape[1]),), torch.uint8, B)
    cur = torch.cuda.current_stream(dev)
    if gauss is None:
        sky_d = _per_galaxy(sky, B, dev, "sky")
        err_d = _per_galaxy(sky_err, B, dev, "sky_err")
        skyrows_d = None
    else:
        from .skyBackground import sky_stats_batch
        g_fw = np.asarray(gauss[0], dtype=np.float64).reshape(B, 2)
        g_th = np.broadcast_to(np.asarray(gauss[1], dtype=np.float64), (B,))
        skyrows_d = torch.empty((B, _abi.PM_SKY_COLS), dtype=torch.float64, device=dev)
    ready = torch.cuda.Event()
    ready.record(cur)
    rows_d = h["rows_d"]
    for i, s in enumerate(range(0, B, chunk)):
        e = min(B, s + chunk)
        st = h["streams"][i & 1]
        if i < 2:
            st.wait_event(ready)
        with torch.cuda.stream(st):
            d = h["bufs"][i & 1][:e - s]
            d.copy_(x[s:e], non_blocking=True)
            if frame is None:
                img = decode_payload(d, bitpix, (e - s, n, n), bzero, bscale, dt, stream=st)
            else:
                img = decode_window(d, bitpix, (H, W), origins[s:e], n, bzero, bscale, dt, stream=st)
            if gauss is None:
                sk, er = sky_d[s:e], err_d[s:e]
            else:
                sr = sky_stats_batch(img, g_fw[s:e], g_th[s:e], stream=st)
                skyrows_d[s:e].copy_(sr)
                sk, er = sr[:, 9].contiguous(), sr[:, 10].contiguous()
            analyse_batch(img, sk, er, filter_size, flags, rows_out=rows_d[s:e], stream=st, **kw)
    for st in h["streams"]:
        cur.wait_stream(st)
    out = torch.empty((B, ROW_DOUBLES), dtype=torch.float64)
    rows_h = h["rows_h"][:B]
    rows_h.copy_(rows_d[:B], non_blocking=True)
    cur.synchronize()
    out.copy_(rows_h)
    if skyrows_d is not None and sky_rows_out is not None:
        sky_rows_out[...] = skyrows_d.cpu().numpy()
    return out.numpy()


# ---- multi-GPU: galaxies are independent, so the batch shards by galaxy (main.py:136-148 did the
# same with a process pool); the only collective is the gather of the result rows. -----------------
def shard_bounds(B, world_size, rank):
    """Contiguous slice [lo, hi) of a batch of B galaxies owned by `rank`."""
    base, rem = divmod(int(B), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_rows(rows_local, B_total, group=None):
    """All-gather the per-rank result rows (contiguous shards in rank order) into [B_total, 16].
    Works with the NCCL backend on CUDA tensors and with gloo on CPU tensors."""
    import torch.distributed as dist
    W = dist.get_world_size(group)
    counts = [shard_bounds(B_total, W, r) for r in range(W)]
    sizes = [hi - lo for lo, hi in counts]
    if len(set(sizes)) == 1:
        out = torch.empty((B_total, rows_local.shape[1]), dtype=rows_local.dtype, device=rows_local.device)
        dist.all_gather_into_tensor(out, rows_local.contiguous(), group=group)
        return out
    mx = max(sizes)
    pad = torch.zeros((mx, rows_local.shape[1]), dtype=rows_local.dtype, device=rows_local.device)
    pad[:rows_local.shape[0]] = rows_local
    parts = [torch.empty_like(pad) for _ in range(W)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)


def analyse_sharded(images_local, sky_local, sky_err_local, B_total, filter_size=3, flags=PM_DO_ALL, group=None, **kw):
    """Each rank analyses its own contiguous shard (already resident on its GPU) and all ranks get
    the full [B_total, 16] table.  Results do not depend on the world size."""
    res = analyse_batch(images_local, sky_local, sky_err_local, filter_size, flags, **kw)
    return gather_rows(res.rows, B_total, group=group)
