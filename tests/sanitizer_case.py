"""Small deterministic workload for compute-sanitizer runs (memcheck / racecheck) on a GPU box:
    compute-sanitizer --tool memcheck python tests/sanitizer_case.py
A few cutout sizes (aligned and odd), both dtypes, all outputs requested, pruned and unpruned minapix."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pawlikmorph_lsst_b200 as pm
import synth

for n, fs, B in ((64, 3, 3), (141, 3, 2), (96, 5, 2), (256, 3, 3)):
    x = synth.make_batch(B, n, 5 + n).cuda()
    r = pm.batch.analyse_batch(x, 100.0, 5.0, fs, 15, want_segmap=True, want_sorted=True, want_apermask=True, cand_cap=64)
    r2 = pm.batch.analyse_batch(x, 100.0, 5.0, fs, 15)
    r3 = pm.batch.analyse_batch(x.double(), 100.0, 5.0, fs, 15 | pm._abi.PM_FORCE_EXACT_FILTER)
    torch.cuda.synchronize()
    assert torch.equal(r.rows, r2.rows), n
    print(n, fs, "ok", r.rows[0, :4].tolist(), flush=True)
print("sanitizer case finished")
