"""Where the reference's default flow (main._device_group on resident cutouts) spends its time, call by call (B200):
python profiles/microbench/default_flow_breakdown.py"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import pawlikmorph_lsst_b200 as pm  # noqa: E402
import synth  # noqa: E402

B, n = 8192, 256
x = torch.cat([synth.make_batch(4096, n, 50 + k, device="cuda:0") for k in range(B // 4096)])
A = pm._abi


def timed(name, fn, out):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    out[name] = round((time.perf_counter() - t0) * 1e3, 2)
    return r


t = {}
clip5 = timed("clipped_stats(maxiters=5)", lambda: pm.skyBackground.clipped_stats_batch(x, 3.0, 5), t)
timed("clipped_stats(maxiters=None)", lambda: pm.skyBackground.clipped_stats_batch(x, 3.0, None), t)
timed("detect_sources (threshold statistics + labels)", lambda: pm.imageutils.detect_sources_batch(x), t)
fit = timed("gaussfit (with its clipped statistics)", lambda: pm.skyBackground.gaussfit_batch(x), t)
timed("gaussfit kernel alone", lambda: pm.skyBackground.gaussfit_batch(x, stats=clip5), t)
f = fit.cpu().numpy()
fw = np.abs(f[:, 4:6]) * 2.3548
timed("sky_stats", lambda: pm.skyBackground.sky_stats_batch(x, np.minimum(fw, 1.414 * n), np.abs(f[:, 6])), t)
bgr = timed("skybgr_batch (all of the above + retries + host logic)", lambda: pm.skyBackground.skybgr_batch(x), t)
sky = torch.as_tensor(bgr["sky"], device=x.device)
err = torch.as_tensor(bgr["sky_err"], device=x.device)
first = timed("analyse: pixelmap pass", lambda: pm.batch.analyse_batch(x, sky, err, 3, 15 | A.PM_STOP_AFTER_PIXELMAP, want_segmap=True), t)
sub = timed("img - sky in float64", lambda: x.to(torch.float64) - sky[:, None, None], t)
clean = timed("maskstars_batch", lambda: pm.imageutils.maskstars_batch(sub, seed=1), t)["clean"]
timed("analyse: statistics on the cleaned image (float64, segmap supplied)",
      lambda: pm.batch.analyse_batch(clean, torch.zeros_like(sky), err, 3, 15, segmap_in=first.segmap, smooth_sky_in=sky).rows.cpu(), t)
timed("whole flow (main._device_group)", lambda: pm.main._device_group(x, None, None, None, 15, 3, True, 1), t)
print(json.dumps({"cutouts": B, "n": n, "ms": t}, indent=1))
