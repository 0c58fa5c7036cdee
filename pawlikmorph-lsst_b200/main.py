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
What is not here: `skybgr` (sky / sky_err are inputs, or the clipped border estimate below,
which is also what the committed goldens used), `maskstarsSEG` cleaning (unseeded noise, imageutils.py:86), star
catalogues, Sersic fits, diagnostic plots.  Options that would need them raise NotImplementedError.
"""
import csv
import os
import time

import numpy as np

from . import _abi
from .image import BYTES_PER_PIXEL, PartialOverlapError, cutout_origin, payload_of, sky_to_pixel
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


def calcMorphology(imageInfos, outfolder, largeImgFactor=None, npix=None, filterSize=3, parallelLibrary="none", cores=1,
                   numberSigmas=5., asymmetry=False, shapeAsymmetry=False, allAsymmetry=True, calculateSersic=False,
                   saveSegMap=False, saveCleanImage=False, imageSource=None, catalogue=None, largeImage=False,
                   paramsaveFile="parameters.csv", occludedSaveFile="occluded-object-locations.csv", segmap=False,
                   starMask=False, CAS=True, *, sky=None, gauss=None, smoothness=False, root=None, device="cuda:0", chunk=8192):
    """Analyse the frames of `imageInfos` (iterable of (filename, ra, dec), e.g. helpers.getFiles) and write
    `outfolder/paramsaveFile`; returns the list of Result in input order.

    sky: None (border_sky of every frame), a callable image -> (sky, sky_err), or a sequence of (sky, sky_err) pairs.
    gauss: a sequence of (fwhms, theta) per image — the Gaussian fit of skyBackground.py:167-262, which is the caller's —
    makes sky / sky_err the reference's own estimate: the statistics of the pixels outside the ellipse of semi-axes
    2 min(fwhms), 2 max(fwhms) (skyBackground.py:134-164), measured on the device from the decoded cutout
    (pm_sky_stats_batch); a sky region of fewer than 100 pixels raises skyBackground._SkyError like upstream.
    root: folder the file names are relative to.  parallelLibrary / cores are accepted and ignored: the batch is the
    parallelism.  npix: None analyses whole frames; a number takes the npix x npix window around (ra, dec) whenever the
    frame has at least npix rows and ra and dec are truthy (Image.py:135 — a declination of exactly 0 gives the whole
    frame, as upstream); a frame that cannot be read, or whose window leaves it (image.PartialOverlapError, the
    `Cutout2D(mode="strict")` error), gets the default Result and the run carries on (main.py:331-339)."""
    for name, val in (("calculateSersic", calculateSersic), ("saveSegMap", saveSegMap), ("saveCleanImage", saveCleanImage),
                      ("catalogue", catalogue), ("largeImage", largeImage), ("segmap", segmap), ("starMask", starMask)):
        if val:
            raise NotImplementedError(f"{name}: outside the morphology path this package covers")
    from .batch import analyse_host_payload      # needs torch + the CUDA library; fails loudly without them
    infos = [(str(f), ra, dec) for f, ra, dec in imageInfos]
    t0 = time.time()
    frames, groups, failed = [], {}, set()
    for i, (f, _, _) in enumerate(infos):
        path = f if root is None or os.path.isabs(f) else os.path.join(str(root), f)
        # main.py:331-339: an image that cannot be read or cut (IOError, AttributeError, PartialOverlapError) gets the
        # default Result and the run carries on with the others
        try:
            with open(path, "rb") as fh:
                payload, hdr = payload_of(fh.read())
            if hdr.get("NAXIS", 2) != 2:
                raise ValueError(f"{f}: not a 2-D image")
            H, W = hdr["NAXIS2"], hdr["NAXIS1"]
            ra, dec = infos[i][1], infos[i][2]
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
        except (OSError, AttributeError, PartialOverlapError, KeyError, ValueError) as e:
            print(f"{f}: {type(e).__name__}: {e} -- default Result")
            frames.append([None, -99., 99., None])
            failed.add(i)
            continue
        pixels = None
        if gauss is not None:
            s = (float("nan"), float("nan"))                                  # measured on the device below
        elif sky is None or callable(sky):
            pixels = _host_pixels(payload, hdr)
            if origin is not None:
                pixels = pixels[origin[0]:origin[0] + n, origin[1]:origin[1] + n]
            s = (sky or border_sky)(pixels)
        else:
            s = sky[i]
        frames.append([payload, float(s[0]), float(s[1]), origin])
        key = (n, H, W, origin is not None, hdr["BITPIX"], float(hdr.get("BZERO", 0.0)), float(hdr.get("BSCALE", 1.0)))
        groups.setdefault(key, []).append(i)
    flags = path_flags(asymmetry, shapeAsymmetry, allAsymmetry, CAS, smoothness)
    rows = np.full((len(infos), _abi.ROW_DOUBLES), -99.0)
    for i in failed:
        rows[i, _abi.ROW_FIELDS.index("status")] = 1.0                         # -> the default Result (main.py:331-339)
    for (n, H, W, windowed, bitpix, bzero, bscale), idx in groups.items():
        pay = np.empty((len(idx), H * W * BYTES_PER_PIXEL[bitpix]), dtype=np.uint8)
        for k, i in enumerate(idx):
            pay[k] = np.frombuffer(frames[i][0], dtype=np.uint8)
        win = dict(frame=(H, W), origins=np.array([frames[i][3] for i in idx], dtype=np.int32)) if windowed else {}
        if gauss is not None:
            sky_rows = np.empty((len(idx), _abi.PM_SKY_COLS))
            win.update(gauss=(np.array([gauss[i][0] for i in idx], dtype=np.float64), np.array([gauss[i][1] for i in idx], dtype=np.float64)),
                       sky_rows_out=sky_rows)
        rows[idx] = analyse_host_payload(pay, bitpix, n, np.array([frames[i][1] for i in idx]), np.array([frames[i][2] for i in idx]),
                                         bzero=bzero, bscale=bscale, filter_size=filterSize, flags=flags, device=device, chunk=chunk,
                                         **win)
        if gauss is not None:
            for k, i in enumerate(idx):
                if sky_rows[k, 11] != 0:
                    from .skyBackground import _SkyError
                    raise _SkyError(f"Error! Sky region too small {int(sky_rows[k, 0])} ({infos[i][0]})")      # skyBackground.py:147-148
                frames[i][1], frames[i][2] = float(sky_rows[k, 9]), float(sky_rows[k, 10])
    per_image = (time.time() - t0) / max(1, len(infos))
    results = []
    for i, (f, _, _) in enumerate(infos):
        r = Result.from_row(rows[i], file=os.path.basename(f), sky=frames[i][1], sky_err=frames[i][2])      # main.py:341: file.name
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
