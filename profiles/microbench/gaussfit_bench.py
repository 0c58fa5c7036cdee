"""Throughput of pm_gaussfit_batch on 256 x 256 float32 cutouts (B200): python profiles/microbench/gaussfit_bench.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import pawlikmorph_lsst_b200 as pm  # noqa: E402
import synth  # noqa: E402

B, n = 8192, 256
x = torch.cat([synth.make_batch(4096, n, 50 + k, device="cuda:0") for k in range(B // 4096)])
clip = pm.skyBackground.clipped_stats_batch(x, 3.0, 5)
fit = pm.skyBackground.gaussfit_batch(x, stats=clip)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(2):
    fit = pm.skyBackground.gaussfit_batch(x, stats=clip)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 2
f = fit.cpu().numpy()
print(json.dumps({"kernel": "pm_gaussfit_kernel<float>", "variant": os.environ.get("PM_LIB_VARIANT", ""), "cutouts": B, "n": n, "ms": ms,
                  "cutouts_per_s": B / ms * 1e3, "mean_iterations": float(f[:, 8].mean()), "converged": float((f[:, 9] == 0).mean()),
                  "mean_sigma": float(abs(f[:, 4:6]).mean())}))
