"""Throughput of pm_sky_stats_batch / pm_clipped_stats_batch per cluster size (PM_STAT_CLUSTER) and cutout size (B200):
python profiles/microbench/stats_cluster_bench.py [n ...]"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import pawlikmorph_lsst_b200 as pm  # noqa: E402
import synth  # noqa: E402


def timed(fn, reps=3):
    fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


for n in [int(a) for a in sys.argv[1:]] or [256]:
    B = max(512, min(16384, (1 << 30) // (n * n * 4)))
    x = torch.cat([synth.make_batch(min(4096, B), n, 50 + k, device="cuda:0") for k in range(max(1, B // 4096))])
    B = x.shape[0]
    rng = np.random.default_rng(1)
    fw, th = rng.uniform(4.0, 16.0, (B, 2)), rng.uniform(-3, 3, B)
    for forced in ("default", "0", "1", "2", "4", "8"):
        if forced == "default":
            os.environ.pop("PM_STAT_CLUSTER", None)
        else:
            os.environ["PM_STAT_CLUSTER"] = forced
        cl, res, smem = C.c_int(), C.c_int(), C.c_int()
        if pm._lib.lib().pm_stats_plan(n, pm._abi.PM_F32, 1, C.byref(cl), C.byref(res), C.byref(smem)) != 0:
            continue
        if forced not in ("default", "0") and not res.value:
            continue                      # this cluster size cannot hold the cutout
        sky_ms, rows = timed(lambda: pm.skyBackground.sky_stats_batch(x, fw, th))
        c5_ms, c5 = timed(lambda: pm.skyBackground.clipped_stats_batch(x, 3.0, 5))
        cn_ms, cn = timed(lambda: pm.skyBackground.clipped_stats_batch(x, 3.0, None))
        gb = B * n * n * 4 / 1e9
        print(json.dumps({"n": n, "cutouts": B, "PM_STAT_CLUSTER": forced, "cluster": cl.value, "resident": res.value, "smem": smem.value,
                          "sky_cutouts_per_s": round(B / sky_ms * 1e3), "sky_GBps_once": round(gb / sky_ms * 1e3, 1),
                          "sky_mean_iters": round(float(rows[:, 4].mean().item()), 2),
                          "clip5_cutouts_per_s": round(B / c5_ms * 1e3), "clip5_GBps_once": round(gb / c5_ms * 1e3, 1),
                          "clip5_mean_iters": round(float(c5[:, 4].mean().item()), 2),
                          "clipN_cutouts_per_s": round(B / cn_ms * 1e3), "clipN_mean_iters": round(float(cn[:, 4].mean().item()), 2)}), flush=True)
