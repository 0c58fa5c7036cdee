"""TEST INFRASTRUCTURE: oracle rows for seeded cutouts whose images are re-made from the seed on the GPU box.

    python oracle/make_seeded_rows.py            (about three minutes on 8 cores)

Writes tests/golden/seeded_rows.npz: for every set (n, flags, count, seed) the rows of oracle/pm_oracle.analyse on
synth.make_exact(count, n, seed) — images that are bit-identical on every machine (integer random numbers, IEEE-exact
operations only) — plus the CRC32 of every pixelmap.  The 512 x 512 set is analysed with SUPPLIED segmaps (the oracle's
own pixelmap, re-made on the GPU box by the pixelmap kernel and checked by CRC first): BASELINE.json configs[4].
tests/test_gpu_seeded.py compares the CUDA path with these rows (>= 1000 cutouts per size)."""
import multiprocessing as mp
import os
import sys
import zlib

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

SETS = (("n140_flags3", 140, 3, 1000, 910140), ("n256_flags15", 256, 15, 1000, 910256), ("n512_segmap", 512, 15, 12, 910512))
FIELDS = ("apix_col", "apix_row", "rmax", "r20", "r80", "C", "A", "Abgr", "S", "As", "As90", "gini", "m20", "object_edge",
          "status", "n_candidates")


def _one(args):
    img, flags, with_segmap = args
    import synth
    from oracle import pm_oracle as O
    x = img.astype(np.float64)
    row, seg, _ = O.analyse(x, synth.SKY, synth.SKY_ERR, 3, flags)
    if with_segmap and row["status"] == 0:
        edge = row["object_edge"]
        row, seg2, _ = O.analyse(x, synth.SKY, synth.SKY_ERR, 3, flags, segmap_in=seg)
        assert np.array_equal(seg, seg2)
        row["object_edge"] = 0.0 * edge          # main.py:427 runs only on the pixelmap path
    return np.array([row[k] for k in FIELDS]), zlib.crc32(np.packbits(seg).tobytes())


def main():
    import synth
    out = {"fields": np.array(FIELDS)}
    with mp.get_context("spawn").Pool(os.cpu_count()) as pool:
        for name, n, flags, count, seed in SETS:
            x = synth.make_exact(count, n, seed)
            res = pool.map(_one, [(x[i], flags, name.endswith("segmap")) for i in range(count)], chunksize=4)
            out[name + "_rows"] = np.stack([r for r, _ in res])
            out[name + "_crc"] = np.array([c for _, c in res], dtype=np.uint32)
            out[name + "_meta"] = np.array([n, flags, count, seed], dtype=np.int64)
            out[name + "_imgcrc"] = np.array([zlib.crc32(x[i].tobytes()) for i in range(count)], dtype=np.uint32)
            print(name, "status counts", np.unique(out[name + "_rows"][:, 14], return_counts=True), flush=True)
    np.savez_compressed(os.path.join(REPO, "tests", "golden", "seeded_rows.npz"), **out)


if __name__ == "__main__":
    main()
