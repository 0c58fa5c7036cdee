"""File driver around the batched path: the part of `pawlikMorphLSST/main.py:36-176` (`calcMorphology`) that reads the
frames of an image list, analyses them and writes `parameters.csv` — with the per-image Python loop / process pool /
parsl executors (main.py:136-159) replaced by one batched GPU call per frame shape.

What is kept: argument names and meaning of `calcMorphology` for this path (filterSize, asymmetry, shapeAsymmetry,
allAsymmetry, CAS, paramsaveFile), the flag logic of `_analyseImage` (main.py:469-491: A for -A/-Aall/-cas, As and As90
for -As/-Aall, r20/r80/C/Gini/M20 for -cas, S never — main.py:494 is commented out — unless `smoothness=True`), the
default Result for a frame whose centre is too faint (main.py:420-424), the CSV header and row format.
The cut-out of Image.py:135-138, :175-186 is kept too: with `npix` not larger than the frame and truthy ra / dec the
npix x npix window around the catalogue position is analysed (sky -> pixel on the host, the copy fused with the decode
on the device); otherwise the whole frame, like the reference.
`skybgr` (skyBackground.py) and the `maskstarsSEG` cleaning (imageutils.py:42-91, with a seeded generator in place of the
reference's unseeded one) run on the device too, and so does `fitSersic` (`calculateSersic`, sersic.py).  What is not here
(DESIGN §7): segmap / clean-image FITS input and output, diagnostic plots.  The star-catalogue flow (`catalogue`,
main.py:372-395: objectOccluded, maskstarsPSF, the statistics with the mask) runs as host mask construction + the star-mask
composition of batch.analyse_batch_starmask.  Options that would need them raise NotImplementedError.
"""
import csv
import os
import time

import numpy as np

from . import _abi
from .image import BYTES_PER_PIXEL, PartialOverlapError, cutout_origin, header_cards, payload_of, sky_to_pixel
from .result import CSV_HEADER, Result

_NP = {8: ">u1", 16: ">i2", 32: ">i4", -32: ">f4", -64: ">f8"}


def border_sky(img, w=10):
    """(sky, sky_err) = 3-sigma-clipped mean / std (five rounds at most) of the outer `w`-pixel border.  A documented
    stand-in for `skyBackground.skybgr` (skyBackground.py:32-165, out of scope); host-side, not part of the hot path."""
    b = np.concatenate([img[:w].ravel(), img[-w:].ravel(), img[w:-w, :w].ravel(), img[w:-w, -w:].ravel()]).astype(np.float64)
    for _ in range(5):
        m, s = b.mean(), b.std()
        keep = np.abs(b - m) <= 3 * s
        if keep.all():
            break
        b = b[keep]
    return float(b.mean()), float(b.std())


def _host_pixels(payload, hdr):
    """float64 host view of a payload (for the sky estimate only)."""
    a = np.frombuffer(payload, dtype=_NP[hdr["BITPIX"]]).astype(np.float64)
    bz, bs = float(hdr.get("BZERO", 0.0)), float(hdr.get("BSCALE", 1.0))
    if not (bz == 0.0 and bs == 1.0):
        a = a * bs + bz
    return a.reshape(hdr["NAXIS2"], hdr["NAXIS1"])


def path_flags(asymmetry=False, shapeAsymmetry=False, allAsymmetry=True, CAS=True, smoothness=False):
    """main.py:469-491 as a flag word of the batched call."""
    f = 0
    if asymmetry or allAsymmetry or CAS:
        f |= _abi.PM_DO_A
    if shapeAsymmetry or allAsymmetry:
        f |= _abi.PM_DO_AS
    if CAS:
        f |= _abi.PM_DO_CAS
    if smoothness:
        f |= _abi.PM_DO_S
    return f


def _decode_dtype(bitpix, bzero, bscale):
    from .batch import payload_dtype
    return payload_dtype(bitpix, bzero, bscale)


def _device_group(images, large, sky, sky_err, flags, filterSize, clean, seed, sersic=False, catalogue=None, star_masks=None,
                  segmaps=None, want_maps=False):
    """One group of equally sized cutouts that already live on the GPU through main.py:349-495: skybgr (unless sky is given),
    pixelmap on the image with the sky in (:421), `img -= sky`, maskstarsSEG (:442-445), the rest on the cleaned image.
    Returns a dict of numpy arrays: rows [B, 16], sky, sky_err [B], fwhms [B, 2], theta [B], and
      sersic [B, 7]              with sersic=True: the fitted Sersic2D parameters of main.py:480-488 (-99 where there is no fit);
      star_flag, masked [B]      with catalogue = dict(path, radec [B, 2], headers [B], numberSigmas) — the `-cc` flow of
                                 main.py:372-395: the catalogue objects near each frame that fall on its pixelmap (objectOccluded), the
                                 PSF-sized star mask of those (maskstarsPSF, adaptive=False), the pixelmap and all statistics with it;
      segmap, clean              with want_maps=True: the pixelmaps [B, n, n] uint8 and the analysed images (img - sky, cleaned)
                                 [B, n, n] float64 as CUDA tensors (main.py:447-461 writes them out).
    star_masks: CUDA [B, n, n] (1 = keep), the `-starmask` flow of main.py:397-417; segmaps: CUDA [B, n, n] uint8, the `-segmap` flow
    of :429-440 (no pixelmap, no edge flag)."""
    import torch

    from .batch import analyse_batch, analyse_batch_starmask
    from .imageutils import maskstars_batch
    from .skyBackground import skybgr_batch
    B, n = int(images.shape[0]), int(images.shape[1])
    col = _abi.ROW_FIELDS.index
    fwhms, theta = np.full((B, 2), -99.0), np.full(B, -99.0)
    sky_fail = np.zeros(B, dtype=bool)
    if sky is None:
        r = skybgr_batch(images, large, seed=seed)
        sky, sky_err, fwhms, theta = r["sky"], r["sky_err"], r["fwhms"], r["theta"]
        sky_fail = r["status"] != 0                                       # _SkyError / ValueError: default Result (main.py:350-368)
        sky, sky_err = np.where(sky_fail, 0.0, sky), np.where(sky_fail, 0.0, sky_err)
    dev = images.device
    sky_d = torch.as_tensor(np.asarray(sky, dtype=np.float64), device=dev)
    err_d = torch.as_tensor(np.asarray(sky_err, dtype=np.float64), device=dev)
    zero = torch.zeros_like(sky_d)
    need_final = clean or sersic or want_maps or catalogue is not None or star_masks is not None or segmaps is not None
    final = None                                                          # the image everything after the pixelmap sees
    if need_final:
        final = images.to(torch.float64) - sky_d[:, None, None]           # main.py:442
        if clean:
            final = maskstars_batch(final, seed=seed)["clean"]            # main.py:445
    star, segmap_d = None, None
    star_flag, masked = np.zeros(B, dtype=bool), np.full(B, -99.0)
    if segmaps is not None:                                               # main.py:429-440
        segmap_d = segmaps
        rows = analyse_batch(final, zero, err_d, filterSize, flags, segmap_in=segmaps, smooth_sky_in=sky_d).rows.cpu().numpy()
        rows[:, col("object_edge")] = 0.0
    elif catalogue is not None or star_masks is not None:
        first = analyse_batch(images, sky_d, err_d, filterSize, flags | _abi.PM_STOP_AFTER_PIXELMAP, want_segmap=True)
        rows1 = first.rows.cpu().numpy()
        if catalogue is not None:
            from .imageutils import maskstarsPSF
            from .objectMasker import objectOccluded
            tmpmask = first.segmap.cpu().numpy()
            masks = np.ones((B, n, n))
            dummy = np.zeros((n, n))                                      # maskstarsPSF(adaptive=False) only takes the shape
            for k in range(B):
                if rows1[k, col("status")] != 0 or sky_fail[k]:
                    continue                                              # main.py:375-378: no pixelmap, default Result
                ra, dec = catalogue["radec"][k]
                star_flag[k], objs = objectOccluded(tmpmask[k], (ra, dec), catalogue["path"], catalogue["headers"][k])
                try:
                    masks[k] = maskstarsPSF(dummy, objs, catalogue["headers"][k], float(sky[k]), catalogue["numberSigmas"], adaptive=False)
                except KeyError as e:                                     # main.py:389-391
                    print(e, ", so not using star catalogue to mask stars!")
            star = torch.as_tensor(masks, device=dev)
        else:
            star = star_masks.to(device=dev, dtype=torch.float64)
        res = analyse_batch_starmask(images, sky_d, err_d, star, filterSize, flags, cleaned=final)
        rows = res.rows.cpu().numpy()
        segmap_d = res.segmap
        early = rows1[:, col("status")] != 0                              # main.py:375-378, :401-405: default Result, nothing else is computed
        rows[early] = rows1[early]
        if catalogue is not None and flags & (_abi.PM_DO_A | _abi.PM_DO_CAS):          # main.py:472-474
            from .pixmap import calcMaskedFraction
            for k in range(B):
                if rows[k, col("status")] == 0 and not sky_fail[k]:
                    masked[k] = calcMaskedFraction(tmpmask[k], masks[k], [rows[k, col("apix_col")], rows[k, col("apix_row")]])
    elif not clean and not want_maps:
        rows = analyse_batch(images, sky_d, err_d, filterSize, flags).rows.cpu().numpy()
    else:
        first = analyse_batch(images, sky_d, err_d, filterSize, flags | _abi.PM_STOP_AFTER_PIXELMAP, want_segmap=True)
        rows1 = first.rows.cpu().numpy()
        segmap_d = first.segmap
        rows = analyse_batch(final, zero, err_d, filterSize, flags, segmap_in=first.segmap, smooth_sky_in=sky_d).rows.cpu().numpy()
        rows[:, col("object_edge")] = rows1[:, col("object_edge")]                               # main.py:427 (pixelmap path)
        faint = rows1[:, col("status")] == 1
        rows[faint] = rows1[faint]                                                               # main.py:420-424
    rows[sky_fail] = -99.0
    rows[sky_fail, col("status")] = 1.0
    out = dict(rows=rows, sky=np.where(sky_fail, -99.0, sky), sky_err=np.where(sky_fail, 99.0, sky_err), fwhms=fwhms, theta=theta,
               star_flag=star_flag, masked=masked)
    if want_maps:
        out["segmap"], out["clean"] = segmap_d, final
    if sersic:
        # main.py:480-488: fitSersic(img, apix, fwhms, theta, starMask) on the analysed cutouts (apix = (x, y) = (column, row))
        from .sersic import fitSersic_batch
        ser = np.full((B, 7), -99.0)
        okrow = np.nonzero((rows[:, col("status")] == 0) & ~sky_fail & (rows[:, col("apix_col")] >= 0))[0]
        if okrow.size:
            pick = None if okrow.size == B else torch.as_tensor(okrow, device=dev)
            sel = final if pick is None else final[pick]
            smask = None if star is None else (star if pick is None else star[pick])
            fit = fitSersic_batch(sel, rows[okrow][:, [col("apix_col"), col("apix_row")]], fwhms[okrow], theta[okrow], star_masks=smask)
            good = np.isin(fit["status"], (0, 1))
            ser[okrow[good]] = fit["params"][good]
        out["sersic"] = ser
    return out


def calcMorphology(imageInfos, outfolder, largeImgFactor=None, npix=None, filterSize=3, parallelLibrary="none", cores=1,
                   numberSigmas=5., asymmetry=False, shapeAsymmetry=False, allAsymmetry=True, calculateSersic=False,
                   saveSegMap=False, saveCleanImage=False, imageSource=None, catalogue=None, largeImage=False,
                   paramsaveFile="parameters.csv", occludedSaveFile="occluded-object-locations.csv", segmap=False,
                   starMask=False, CAS=True, *, sky=None, gauss=None, smoothness=False, clean=True, seed=0, root=None,
                   device="cuda:0", chunk=8192):
    """Analyse the frames of `imageInfos` (iterable of (filename, ra, dec), e.g. helpers.getFiles) and write
    `outfolder/paramsaveFile`; returns the list of Result in input order.  The steps of main.py:349-495 per frame: sky
    background, pixelmap, `img -= sky`, cleaning of external sources, the morphology statistics — batched per frame format.

    sky: None — `skyBackground.skybgr` on the device, like upstream (main.py:349; with largeImage=True and largeImgFactor the
    `-li` mode: source check and sky region on the cleaned cutout of npix * largeImgFactor pixels, skyBackground.py:71-76) —,
    "border" / a callable image -> (sky, sky_err) (host; `border_sky` is the stand-in the goldens were made with), or a
    sequence of (sky, sky_err) pairs.  gauss: a sequence of (fwhms, theta) per image: the sky-region statistics for a
    caller-supplied Gaussian fit.
    clean: the maskstarsSEG step between pixelmap and the statistics (main.py:445), seeded by `seed` (the reference's generator
    is unseeded: imageutils.py:86); clean=False analyses the sky-subtracted image as it is.
    root: folder the file names are relative to.  parallelLibrary / cores are accepted and ignored: the batch is the
    parallelism.  npix: None analyses whole frames; a number takes the npix x npix window around (ra, dec) whenever the
    frame has at least npix rows and ra and dec are truthy (Image.py:135 — a declination of exactly 0 gives the whole
    frame, as upstream); a frame that cannot be read, or whose window leaves it (image.PartialOverlapError, the
    `Cutout2D(mode="strict")` error), gets the default Result and the run carries on (main.py:331-339); so does a frame
    whose sky estimate fails (main.py:350-368)."""
    special = bool(catalogue or segmap or starMask or saveSegMap or saveCleanImage or calculateSersic)
    if segmap and (catalogue or starMask):                                    # imganalysis.py:75-83
        raise ValueError("Can't provide both a precomputed segmentation map and a star catalogue / star mask file!")
    if starMask and catalogue:
        raise ValueError("Can't provide both a star catalogue and precomputed star mask file!")
    if special and gauss is not None:
        raise NotImplementedError("catalogue / segmap / starMask / saveSegMap / saveCleanImage / calculateSersic run in the device flow (gauss=None)")
    if calculateSersic and (sky is not None or gauss is not None):
        raise NotImplementedError("calculateSersic needs the widths and the angle of the sky estimate: run with sky=None, like upstream")
    if isinstance(sky, str) and sky == "border":
        sky = border_sky
    from . import batch as _batch                    # needs torch + the CUDA library; fails loudly without them
    infos = [(str(f), ra, dec) for f, ra, dec in imageInfos]
    t0 = time.time()
    frames, groups, failed = [], {}, set()
    device_path = (sky is None and gauss is None) or clean or special
    for i, (f, _, _) in enumerate(infos):
        path = f if root is None or os.path.isabs(f) else os.path.join(str(root), f)
        # main.py:331-339: an image that cannot be read or cut (IOError, AttributeError, PartialOverlapError) gets the
        # default Result and the run carries on with the others
        try:
            with open(path, "rb") as fh:
                raw_file = fh.read()
            payload, hdr = payload_of(raw_file)
            cards_of_frame = header_cards(raw_file) if (saveSegMap or saveCleanImage) else None
            if hdr.get("NAXIS", 2) != 2:
                raise ValueError(f"{f}: not a 2-D image")
            H, W = hdr["NAXIS2"], hdr["NAXIS1"]
            ra, dec = infos[i][1], infos[i][2]
            origin_large, n_large = None, None
            if npix is not None and H >= npix and ra and dec:                      # Image.py:135-136
                n = int(npix)
                # sky_to_pixel returns the true (x = column, y = row) for RA-first and DEC-first headers alike; the
                # reference's position[::-1] (Image.py:179-182) only undoes astropy #10468 and has no counterpart here
                x, y = sky_to_pixel(hdr, ra, dec)
                origin = cutout_origin(x, y, n, (H, W))
            else:
                if H != W:
                    raise ValueError(f"{f}: whole frames must be square, got {H} x {W}")
                n, origin = H, None
            if largeImage and sky is None:                                          # main.py:323-328
                nl = int(npix * largeImgFactor) if (npix is not None and largeImgFactor) else H
                if H >= nl and ra and dec and nl < H:
                    x, y = sky_to_pixel(hdr, ra, dec)
                    origin_large, n_large = cutout_origin(x, y, nl, (H, W)), nl
                else:
                    if H != W:
                        raise ValueError(f"{f}: whole frames must be square, got {H} x {W}")
                    n_large = H                                                     # the whole frame
        except (OSError, AttributeError, PartialOverlapError, KeyError, ValueError) as e:
            print(f"{f}: {type(e).__name__}: {e} -- default Result")
            frames.append([None, -99., 99., None, None, None, None])
            failed.add(i)
            continue
        pixels = None
        if gauss is not None or sky is None:
            s = (float("nan"), float("nan"))                                  # measured on the device below
        elif callable(sky):
            pixels = _host_pixels(payload, hdr)
            if origin is not None:
                pixels = pixels[origin[0]:origin[0] + n, origin[1]:origin[1] + n]
            s = sky(pixels)
        else:
            s = sky[i]
        frames.append([payload, float(s[0]), float(s[1]), origin, origin_large, hdr, cards_of_frame])
        key = (n, H, W, origin is not None, hdr["BITPIX"], float(hdr.get("BZERO", 0.0)), float(hdr.get("BSCALE", 1.0)), n_large, origin_large is not None)
        groups.setdefault(key, []).append(i)
    flags = path_flags(asymmetry, shapeAsymmetry, allAsymmetry, CAS, smoothness)
    rows = np.full((len(infos), _abi.ROW_DOUBLES), -99.0)
    for i in failed:
        rows[i, _abi.ROW_FIELDS.index("status")] = 1.0                         # -> the default Result (main.py:331-339)
    fw_all, th_all = np.full((len(infos), 2), -99.0), np.full(len(infos), -99.0)
    ser_all = np.full((len(infos), 7), -99.0)
    flag_all, masked_all = np.zeros(len(infos), dtype=bool), np.full(len(infos), -99.0)
    saved = [dict() for _ in infos]
    for (n, H, W, windowed, bitpix, bzero, bscale, n_large, large_windowed), idx in groups.items():
        pay = np.empty((len(idx), H * W * BYTES_PER_PIXEL[bitpix]), dtype=np.uint8)
        for k, i in enumerate(idx):
            pay[k] = np.frombuffer(frames[i][0], dtype=np.uint8)
        win = dict(frame=(H, W), origins=np.array([frames[i][3] for i in idx], dtype=np.int32)) if windowed else {}
        if device_path and gauss is None:
            import torch

            from .image import decode_payload, decode_window
            dt = _decode_dtype(bitpix, bzero, bscale)
            pd = torch.from_numpy(pay).to(device)
            if windowed:
                images = decode_window(pd, bitpix, (H, W), win["origins"], n, bzero, bscale, dt)
            else:
                images = decode_payload(pd, bitpix, (len(idx), n, n), bzero, bscale, dt)
            large = None
            if n_large is not None:
                if large_windowed:
                    large = decode_window(pd, bitpix, (H, W), np.array([frames[i][4] for i in idx], dtype=np.int32), n_large, bzero, bscale, dt)
                else:
                    large = decode_payload(pd, bitpix, (len(idx), H, W), bzero, bscale, dt)
            have_sky = sky is not None
            cat = None
            if catalogue:                                                       # main.py:372-395; the frame's own header, like upstream
                cat = dict(path=str(catalogue), radec=[(infos[i][1], infos[i][2]) for i in idx], headers=[frames[i][5] for i in idx],
                           numberSigmas=numberSigmas)
            # precomputed maps of the -segmap / -starmask flows: <outfolder>/segmap_<name>, starmask_<name> (main.py:407-414, :431-438)
            def maps_of(prefix):
                arr = np.empty((len(idx), n, n), dtype=np.float64)
                for k, i in enumerate(idx):
                    with open(os.path.join(str(outfolder), prefix + os.path.basename(infos[i][0])), "rb") as fh:
                        p_, h_ = payload_of(fh.read())
                    a_ = _host_pixels(p_, h_)
                    if a_.shape != (n, n):
                        raise ValueError(f"{prefix}{os.path.basename(infos[i][0])}: {a_.shape}, expected {(n, n)}")
                    arr[k] = a_
                return arr
            seg_in = torch.as_tensor((maps_of("segmap_") != 0).astype(np.uint8), device=device) if segmap else None
            star_in = torch.as_tensor(maps_of("starmask_"), device=device) if starMask else None
            want_maps = bool(saveSegMap or saveCleanImage)
            got = _device_group(images, large, np.array([frames[i][1] for i in idx]) if have_sky else None,
                                np.array([frames[i][2] for i in idx]) if have_sky else None, flags, filterSize, clean,
                                seed, sersic=bool(calculateSersic), catalogue=cat, star_masks=star_in, segmaps=seg_in, want_maps=want_maps)
            r, sk, er, fw, th = got["rows"], got["sky"], got["sky_err"], got["fwhms"], got["theta"]
            if calculateSersic:
                ser_all[idx] = got["sersic"]
            if catalogue:
                flag_all[idx], masked_all[idx] = got["star_flag"], got["masked"]
            if want_maps:                                                       # main.py:447-461 (float64 images under the frame's header)
                from .image import fits_image_bytes
                os.makedirs(str(outfolder), exist_ok=True)
                seg_h = got["segmap"].cpu().numpy() if saveSegMap and got["segmap"] is not None else None
                cln_h = got["clean"].cpu().numpy() if saveCleanImage else None
                for k, i in enumerate(idx):
                    if r[k, _abi.ROW_FIELDS.index("status")] == 1:
                        continue                                                # default Result: upstream returned before writing
                    name = os.path.basename(infos[i][0])
                    cards = frames[i][6]
                    if cln_h is not None:
                        saved[i]["clean"] = os.path.join(str(outfolder), "clean_" + name)
                        with open(saved[i]["clean"], "wb") as fh:
                            fh.write(fits_image_bytes(cln_h[k], cards))
                    if seg_h is not None:
                        saved[i]["segmap"] = os.path.join(str(outfolder), "segmap_" + name)
                        with open(saved[i]["segmap"], "wb") as fh:
                            fh.write(fits_image_bytes(seg_h[k], cards))
            rows[idx] = r
            fw_all[idx], th_all[idx] = fw, th
            for k, i in enumerate(idx):
                frames[i][1], frames[i][2] = float(sk[k]), float(er[k])
            continue
        if gauss is not None:
            sky_rows = np.empty((len(idx), _abi.PM_SKY_COLS))
            win.update(gauss=(np.array([gauss[i][0] for i in idx], dtype=np.float64), np.array([gauss[i][1] for i in idx], dtype=np.float64)),
                       sky_rows_out=sky_rows)
        rows[idx] = _batch.analyse_host_payload(pay, bitpix, n, np.array([frames[i][1] for i in idx]), np.array([frames[i][2] for i in idx]),
                                                bzero=bzero, bscale=bscale, filter_size=filterSize, flags=flags, device=device, chunk=chunk,
                                                **win)
        if gauss is not None:
            for k, i in enumerate(idx):
                if sky_rows[k, 11] != 0:
                    from .skyBackground import _SkyError
                    raise _SkyError(f"Error! Sky region too small {int(sky_rows[k, 0])} ({infos[i][0]})")      # skyBackground.py:147-148
                frames[i][1], frames[i][2] = float(sky_rows[k, 9]), float(sky_rows[k, 10])
                fw_all[i], th_all[i] = np.asarray(gauss[i][0], dtype=np.float64), float(gauss[i][1])
    per_image = (time.time() - t0) / max(1, len(infos))
    results = []
    for i, (f, _, _) in enumerate(infos):
        r = Result.from_row(rows[i], file=os.path.basename(f), sky=frames[i][1], sky_err=frames[i][2])      # main.py:341: file.name
        if r.status != 1 and fw_all[i, 0] != -99.0:
            r.fwhms, r.theta = [float(fw_all[i, 0]), float(fw_all[i, 1])], float(th_all[i])                # main.py:349
        if calculateSersic and r.status != 1:                                                              # main.py:480-488
            (r.sersic_amplitude, r.sersic_r_eff, r.sersic_n, r.sersic_x_0, r.sersic_y_0, r.sersic_ellip,
             r.sersic_theta) = (float(v) for v in ser_all[i])
        if catalogue and r.status != 1:                                                                    # main.py:380-381, :472-474
            r.star_flag, r.maskedPixelFraction = bool(flag_all[i]), float(masked_all[i])
        if "clean" in saved[i]:
            r.cleanImage = saved[i]["clean"]                                                              # main.py:451
        if "segmap" in saved[i]:
            r.pixelMapFile = saved[i]["segmap"]                                                            # main.py:459
        r.outfolder = str(outfolder)
        r.time = per_image
        results.append(r)
    os.makedirs(str(outfolder), exist_ok=True)
    with open(os.path.join(str(outfolder), paramsaveFile), mode="w", newline="") as fh:
        w = csv.writer(fh, delimiter=",")
        w.writerow(CSV_HEADER)
        for r in results:
            r.write(w)
    return results
