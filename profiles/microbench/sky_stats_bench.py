"""Throughput of pm_sky_stats_batch on 256 x 256 float32 cutouts (B200): python profiles/microbench/sky_stats_bench.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import pawlikmorph_lsst_b200 as pm  # noqa: E402
import synth  # noqa: E402

B, n = 16384, 256
x = torch.cat([synth.make_batch(4096, n, 50 + k, device="cuda:0") for k in range(B // 4096)])
rng = np.random.default_rng(1)
fw, th = rng.uniform(4.0, 16.0, (B, 2)), rng.uniform(-3, 3, B)
for _ in range(2):
    rows = pm.skyBackground.sky_stats_batch(x, fw, th)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(3):
    rows = pm.skyBackground.sky_stats_batch(x, fw, th)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
r = rows.cpu().numpy()
print(json.dumps({"kernel": "pm_sky_stats_kernel<float>", "cutouts": B, "n": n, "ms": ms, "cutouts_per_s": B / ms * 1e3,
                  "clipped_fraction": float((r[:, 4] > 0).mean()), "mean_clip_iterations": float(r[:, 4].mean()),
                  "image_GBps_once": B * n * n * 4 / ms / 1e6}))
