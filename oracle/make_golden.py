"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED
reference hot-path modules (via oracle/ref_import.py) in the build container.

TEST INFRASTRUCTURE.  Run once here (``python -m oracle.make_golden``); the
outputs are committed so that the GPU box, which has no /root/reference, can
check both the CPU port (oracle/pm_oracle.py) and the CUDA path against them.

Per cutout the reference functions are called in the order of main.py:421-495
(cleaning = identity, S by a direct `smoothness` call).  The only
intervention is the canonical tie order for `np.argsort` inside `minapix`
(asymmetry.py:95): the reference's default is unstable and build dependent, so
the run is made twice — native and canonical (stable) — and the file records
where they differ (`native_differs`); the stored values are the canonical ones.

Sets:
  sdss_l.npz : the 69 `sample/sdsslcutout_*` frames (68 listed in images.csv + 1), full frame
  sdss_s.npz : the 69 `sample/sdsscutout_*` 141x141 frames, full frame
  synth.npz  : seeded synthetic Sersic+noise cutouts at 64/128/140/256 (+ one 512)

sky / sky_err for the SDSS frames: 3-sigma-clipped mean / std of the outer
10-pixel border (the real `skybgr` is out of scope) — stored so both
sides use identical numbers.
"""
import glob
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)

from oracle import fitsio, ref_import  # noqa: E402

FIELDS = ("apix_col", "apix_row", "rmax", "r20", "r80", "C", "A", "Abgr", "S", "As",
          "As90", "gini", "m20", "object_edge", "status")


class _StableNP:
    """numpy look-alike whose argsort is stable (canonical tie order)."""

    def __getattr__(self, k):
        return getattr(np, k)

    @staticmethod
    def argsort(a, *args, **kw):
        return np.argsort(a, kind="stable")


def border_sky(img, w=10):
    b = np.concatenate([img[:w].ravel(), img[-w:].ravel(), img[w:-w, :w].ravel(), img[w:-w, -w:].ravel()])
    b = b.astype(np.float64)
    for _ in range(5):
        m, s = b.mean(), b.std()
        keep = np.abs(b - m) <= 3 * s
        if keep.all():
            break
        b = b[keep]
    return float(b.mean()), float(b.std())


def run_reference(ref, img_with_sky, sky, sky_err, fs, canonical):
    out = {k: -99.0 for k in FIELDS}
    out["object_edge"] = 0.0
    out["status"] = 0.0
    img = np.array(img_with_sky, dtype=np.float64)
    n = img.shape[0]
    try:
        seg = ref.pixmap.pixelmap(img, sky + sky_err, fs)
    except AttributeError:
        out["status"] = 1.0
        return out, np.zeros((n, n), np.uint8)
    seg_u8 = seg.astype(np.uint8)
    out["object_edge"] = float(ref.pixmap.checkPixelmapEdges(seg))
    img -= sky
    ones = np.ones_like(img)
    rmax = ref.pixmap.calcRmax(img, seg)
    out["rmax"] = float(rmax)
    ap = ref.apertures.aperpixmap(n, rmax, 9, 0.1)
    saved = ref.asymmetry.np
    try:
        if canonical:
            ref.asymmetry.np = _StableNP()
        try:
            apix = ref.asymmetry.minapix(img, seg, ap, ones)
        except ValueError:                    # np.argmin([]) upstream crash
            out["status"] = 2.0
            return out, seg_u8
    finally:
        ref.asymmetry.np = saved
    out["apix_col"], out["apix_row"] = float(apix[0]), float(apix[1])
    A = ref.asymmetry.calcA(img, seg, ap, apix, 180.0, ones, noisecorrect=True)
    out["A"], out["Abgr"] = float(A[0]), float(A[1])
    out["As"] = float(ref.asymmetry.calcA(seg, seg, ap, apix, 180.0, ones)[0])
    out["As90"] = float(ref.asymmetry.calcA(seg, seg, ap, apix, 90.0, ones)[0])
    try:
        r20, r80 = ref.casgm.calcR20_R80(img, seg, apix, rmax)
        out["r20"], out["r80"] = float(r20), float(r80)
        out["C"] = float(ref.casgm.concentration(r20, r80))
    except (ValueError, RuntimeError):
        out["status"] = 3.0
    out["gini"] = float(ref.casgm.gini(img, seg))
    out["m20"] = float(ref.casgm.m20(img, seg))
    if out["status"] == 0.0 and int(out["r20"]) >= 1:
        out["S"] = float(ref.casgm.smoothness(img, seg, apix, rmax, out["r20"], sky))
    return out, seg_u8


def run_set(ref, name, images, skies, fs_list, extra=None):
    rows, segs, differs = [], [], []
    t0 = time.time()
    for i, (im, (sky, err), fs) in enumerate(zip(images, skies, fs_list)):
        can, seg = run_reference(ref, im, sky, err, fs, canonical=True)
        nat, _ = run_reference(ref, im, sky, err, fs, canonical=False)
        differs.append(any(can[k] != nat[k] and not (np.isnan(can[k]) and np.isnan(nat[k])) for k in FIELDS))
        rows.append([can[k] for k in FIELDS])
        segs.append(np.packbits(seg, axis=None))
        print(f"[{name}] {i + 1}/{len(images)} n={im.shape[0]} fs={fs} segpx={int(seg.sum())} "
              f"status={int(can['status'])} native_differs={differs[-1]}  ({time.time() - t0:.0f}s)", flush=True)
    data = dict(fields=np.array(FIELDS), rows=np.array(rows, dtype=np.float64),
                sky=np.array([s[0] for s in skies]), sky_err=np.array([s[1] for s in skies]),
                filter_size=np.array(fs_list, dtype=np.int32),
                sizes=np.array([im.shape[0] for im in images], dtype=np.int32),
                native_differs=np.array(differs))
    for i, (im, sg) in enumerate(zip(images, segs)):
        data[f"img_{i}"] = im
        data[f"seg_{i}"] = sg
    if extra:
        data.update(extra)
    os.makedirs(os.path.join(REPO, "tests", "golden"), exist_ok=True)
    np.savez_compressed(os.path.join(REPO, "tests", "golden", name + ".npz"), **data)


def load_sdss(pattern):
    files = sorted(glob.glob(os.path.join(ref_import.REF_ROOT, "sample", pattern)))
    imgs, names = [], []
    for f in files:
        a = fitsio.getdata(f)
        if a.dtype.kind == "f" and np.all(a == np.rint(a)) and a.min() >= 0 and a.max() < 65536:
            a = a.astype(np.uint16)           # lossless: the frames hold integer counts
        imgs.append(np.ascontiguousarray(a))
        names.append(os.path.basename(f))
    return imgs, names


def main(which):
    import synth
    ref = ref_import.load()
    print("reference loaded; restated leaves:", ref.shimmed)
    if "sdss_s" in which:
        imgs, names = load_sdss("sdsscutout_*")
        skies = [border_sky(im) for im in imgs]
        run_set(ref, "sdss_s", imgs, skies, [3] * len(imgs), extra=dict(names=np.array(names)))
    if "sdss_l" in which:
        imgs, names = load_sdss("sdsslcutout_*")
        skies = [border_sky(im) for im in imgs]
        run_set(ref, "sdss_l", imgs, skies, [3] * len(imgs), extra=dict(names=np.array(names)))
    if "synth" in which:
        imgs, skies, fss = [], [], []
        for n, cnt, seed in ((64, 10, 11), (128, 8, 12), (140, 8, 13), (256, 8, 14), (512, 1, 15)):
            x = synth.make_batch(cnt, n, seed).numpy()
            for i in range(cnt):
                imgs.append(x[i])
                skies.append((synth.SKY, synth.SKY_ERR))
                fss.append(5 if (n in (128, 256) and i >= cnt - 2) else 3)
        # two faint-centre cases (status 1) and a flat-noise frame
        rng = np.random.default_rng(99)
        imgs.append((100.0 + rng.normal(0, 5, (64, 64))).astype(np.float32)); skies.append((100.0, 50.0)); fss.append(3)
        z = synth.make_batch(1, 128, 77).numpy()[0]
        imgs.append(z); skies.append((100.0, 1e6)); fss.append(3)
        run_set(ref, "synth", imgs, skies, fss)


if __name__ == "__main__":
    main(sys.argv[1:] or ["sdss_s", "sdss_l", "synth"])
