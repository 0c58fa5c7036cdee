"""One decode of 8192 x 256 x 256 BITPIX 16 pixels into float32 (the ncu target for pm_fits_decode_kernel)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import pawlikmorph_lsst_b200 as pm  # noqa: E402

count = 8192 * 256 * 256
raw = torch.randint(0, 256, (count * 2,), dtype=torch.uint8, device="cuda:0")
for _ in range(3):
    out = pm.image.decode_payload(raw, 16, (count,), 32768.0, 1.0, torch.float32)
torch.cuda.synchronize()
print("ok", float(out[:8].sum()))
