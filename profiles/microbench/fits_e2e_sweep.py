"""Tuning aid (B200): (1) pm_fits_decode_batch alone against the measured HBM copy bandwidth; (2) the host payload
call analyse_host_payload against analyse_host over chunk sizes and batch sizes.  Run: python profiles/microbench/fits_e2e_sweep.py"""
import json
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
import pawlikmorph_lsst_b200 as pm  # noqa: E402
import synth  # noqa: E402

dev = torch.device("cuda:0")
peak = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"]
n = 256

# ---- (1) decode kernel alone: inputs larger than L2 (126 MB)
for bitpix, out_dt in ((16, torch.float32), (-32, torch.float32), (16, torch.float64), (8, torch.float32), (-64, torch.float64)):
    count = 8192 * n * n
    bpp = pm.image.BYTES_PER_PIXEL[bitpix]
    raw = torch.randint(0, 256, (count * bpp,), dtype=torch.uint8, device=dev)
    if bitpix < 0:      # keep the floats finite
        raw.view(torch.int16)[(1 if bitpix == -32 else 3)::(2 if bitpix == -32 else 4)] &= 0x3fff
    for _ in range(3):
        out = pm.image.decode_payload(raw, bitpix, (count,), 32768.0 if bitpix == 16 else 0.0, 1.0, out_dt)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    K = 10
    for _ in range(K):
        out = pm.image.decode_payload(raw, bitpix, (count,), 32768.0 if bitpix == 16 else 0.0, 1.0, out_dt)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    gb = count * (bpp + out.element_size()) / 1e9
    print(json.dumps({"kernel": "pm_fits_decode_kernel", "bitpix": bitpix, "out": str(out_dt), "pixels": count, "ms": ms,
                      "GBps": gb / ms * 1e3, "frac_of_hbm_peak": gb / ms * 1e3 / peak}), flush=True)
    del raw, out

# ---- (1b) cut-out decode: 128 x 128 windows out of 231 x 231 BITPIX 16 frames (the SDSS sample geometry)
Bw, Hh, Ww, npix = 65536, 231, 231, 128
raw = torch.randint(0, 256, (Bw * Hh * Ww * 2,), dtype=torch.uint8, device=dev)
org = torch.randint(40, 60, (Bw, 2), dtype=torch.int32)
for _ in range(3):
    out = pm.image.decode_window(raw, 16, (Hh, Ww), org, npix, 32768.0, 1.0, torch.float32)
org_d = org.to(dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(5):
    rc = pm._lib.lib().pm_fits_decode_window_batch(raw.data_ptr(), 16, 32768.0, 1.0, Bw, Hh, Ww, org_d.data_ptr(), npix, out.data_ptr(),
                                                   0, torch.cuda.current_stream().cuda_stream)
    assert rc == 0
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
gb = Bw * npix * npix * 6 / 1e9
print(json.dumps({"kernel": "pm_fits_window16_kernel<float>", "frames": Bw, "frame": [Hh, Ww], "npix": npix, "ms": ms,
                  "algorithmic_GBps": gb / ms * 1e3, "frac_of_hbm_peak": gb / ms * 1e3 / peak,
                  "note": "algorithmic bytes = window pixels x 6; the 2-byte rows start at odd offsets, so DRAM reads whole sectors"}), flush=True)
del raw, out, org_d

# ---- (2) host calls
if "--kernels-only" in sys.argv:
    sys.exit(0)
for E, chunk in ((8192, 512), (8192, 1024), (8192, 2048), (32768, 1024), (32768, 2048), (32768, 4096)):
    x = torch.cat([synth.make_batch(4096, n, 7 + k, device=dev) for k in range(E // 4096)]).round_().clamp_(0, 65535)
    pay = (x.to(torch.int32) - 32768).to(torch.int16).cpu().numpy().astype(">i2")
    pay_h = torch.from_numpy(pay.view(np.uint8).reshape(E, -1)).pin_memory()
    xh = x.cpu().pin_memory()
    del x
    res = {}
    for name, call in (("float32", lambda: pm.batch.analyse_host(xh, synth.SKY, synth.SKY_ERR, 3, 15, device=dev, chunk=chunk)),
                       ("fits16", lambda: pm.batch.analyse_host_payload(pay_h, 16, n, synth.SKY, synth.SKY_ERR, bzero=32768.0,
                                                                         filter_size=3, flags=15, device=dev, chunk=chunk))):
        call()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        res[name] = E * 3 / (time.perf_counter() - t0)
    print(json.dumps({"galaxies": E, "chunk": chunk, "gal_per_s": res}), flush=True)
    del xh, pay_h
    pm.batch._host_cache.clear()
