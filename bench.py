#!/usr/bin/env python
"""bench.py — galaxies/s of the batched `-Aall -cas` (+S, Gini, M20) morphology hot path at 256x256.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path (pm_analyse_batch: pixelmap -> rmax/aperture -> sort -> minapix ->
calcA x3 -> r20/r80/C -> gini/m20 -> S) over the resident synthetic batch of this rank; for N > 1 every
rank owns a contiguous shard of the galaxies (weak scaling: fixed galaxies per GPU) and the step ends
with the all-gather of the [B, 16] result rows.  Prints ONE JSON line (rank 0).

`--impl reference` times the reference algorithm on the host cores: the CPU oracle port
(oracle/pm_oracle.py; the reference itself is pure Python and its third-party leaves are not installed,
DESIGN.md) driven like `-par multi` (main.py:136-148) by a multiprocessing pool.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC = "galaxies/sec for -Aall -cas at 256^2"
UNIT = "galaxies/s"
FLAGS_ALL = 15 | int(os.environ.get("PM_BENCH_EXTRA_FLAGS", "0"), 0)    # tuning only, e.g. 0x2000 = no L2 prefetch


def algorithmic_bytes(n, itemsize=4):
    """image read once + uint8 pixelmap written + one 128-byte result row."""
    return itemsize * n * n + n * n + 128


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------
def _cpu_init():
    import numpy as np

    import synth
    from oracle import pm_oracle as O
    x = synth.make_batch(1, 64, 1).numpy()
    O.analyse(x[0].astype(np.float64), synth.SKY, synth.SKY_ERR, 3)     # numba / import warm-up


def _cpu_one(img):
    import numpy as np

    import synth
    from oracle import pm_oracle as O
    row, _, _ = O.analyse(img.astype(np.float64), synth.SKY, synth.SKY_ERR, 3, FLAGS_ALL)
    return row["status"]


def cpu_pool_rate(images, procs):
    """galaxies/s of the oracle port over `images` with a pool of `procs` workers (warm)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs, initializer=_cpu_init) as pool:
        pool.map(_cpu_one, [images[0]] * procs)                           # every worker warm
        t0 = time.perf_counter()
        pool.map(_cpu_one, list(images), chunksize=1)
        dt = time.perf_counter() - t0
    return len(images) / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import synth
    procs = os.cpu_count() or 1
    S = args.cpu_sample or 8 * procs
    images = synth.make_batch(S, args.n, 20261019 + 3000).numpy()
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    times = []
    with ctx.Pool(procs, initializer=_cpu_init) as pool:
        pool.map(_cpu_one, [images[0]] * procs)
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            pool.map(_cpu_one, list(images), chunksize=1)
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
    total = sum(times)
    value = S * len(times) / total
    sample = f"{S} synthetic {args.n}x{args.n} cutouts per step (seed 20261019+3000), multiprocessing.Pool({procs})"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"synthetic Sersic+noise {args.n}x{args.n} float32 cutouts, -Aall -cas +S +Gini/M20 "
                                   f"(BASELINE.json configs[2]), bounded sample", "filter_size": 3, "flags": FLAGS_ALL},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        try:
            for ln in open(self.path):
                p = [s.strip() for s in ln.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1])); mx.append(float(p[2]))
                except ValueError:
                    continue
                for nm, v in zip(names, p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch

    import pawlikmorph_lsst_b200 as pm
    import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    n = args.n
    # ---- resident synthetic batch (per rank)
    free, _ = torch.cuda.mem_get_info(dev)
    per = n * n * 4
    per_out = n * n + 128
    B = args.batch
    cap = int((free - 6 * 2 ** 30) * 0.9 // (per + per_out))
    if B > cap:
        B = max(1024, cap // 1024 * 1024)
    t0 = time.perf_counter()
    x = synth.make_batch(B, n, 20261019 + 1000 * 2 + rank, device=dev, chunk=2048)
    torch.cuda.synchronize(dev)
    gen_s = time.perf_counter() - t0
    sky = torch.full((B,), synth.SKY, dtype=torch.float64, device=dev)
    err = torch.full((B,), synth.SKY_ERR, dtype=torch.float64, device=dev)
    rows = torch.empty((B, 16), dtype=torch.float64, device=dev)
    segmap = torch.empty((B, n, n), dtype=torch.uint8, device=dev)      # the pixelmap is an output of the path
    table = torch.empty((B * world, 16), dtype=torch.float64, device=dev) if world > 1 else None
    stream = torch.cuda.current_stream(dev)

    def step():
        pm.batch.analyse_batch(x, sky, err, 3, FLAGS_ALL, rows_out=rows, segmap_out=segmap, stream=stream)
        if world > 1:
            dist.all_gather_into_tensor(table, rows)

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    fence()
    if args.phases and rank == 0:
        # tuning aid: per-phase clock64() cycles summed over the CTAs of one extra (untimed) step
        names = ("pixelmap", "list+window", "rmax+aperture", "sort", "gini/m20/cands", "minapix", "calcA:loop", "r20:brentq",
                 "S:bkg", "calcA:dilate", "r20:fold", "r20:bounds", "r20:walk", "S:interior", "S:cut", "S:boxsearch",
                 "sort:bits", "sort:count", "sort:scan", "sort:scatter", "minapix:first", "minapix:stage", "minapix:compact", "p23")
        names += tuple(f"#alive@stage{k}" for k in range(8))
        names += ("gini:sums", "gini:thresh", "gini:moments", "gini:cands", "gini:reverse", "list:rowscan", "list:fill", "r20:total", "rmax2", "minapix:load", "tail:planes", "q43", "q44", "q45", "q46", "q47")
        buf = torch.zeros((pm._abi.PM_PHASE_ROWS, 48), dtype=torch.int64, device=dev)
        pm.batch.analyse_batch(x, sky, err, 3, FLAGS_ALL, rows_out=rows, segmap_out=segmap, stream=stream, phase_cycles_out=buf)
        torch.cuda.synchronize(dev)
        cyc = buf.sum(dim=0).cpu().numpy().astype(float)
        tot = cyc.sum()
        print("phase cycles per galaxy: " + ", ".join((f"{nm} {c / B:.3f}" if nm.startswith("#") else f"{nm} {c / B:.0f} ({100 * c / tot:.1f}%)") for nm, c in zip(names, cyc) if c > 0)
              + f"; total {tot / B:.0f}", file=sys.stderr, flush=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = pm._lib.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record(stream)
    for k in range(args.steps):
        step()
        ev[k + 1].record(stream)
    fence()
    launches = pm._lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[args.steps])
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = world * B * args.steps / (total_ms * 1e-3)
    kernel_ms = total_ms / args.steps          # one kernel launch per step (the gather is not a kernel of ours)
    status = rows[:, 14]
    n_ok = int((status == 0).sum().item())
    mean_cands = float(rows[:, 15].mean().item())
    if args.kernel_only:
        if rank == 0:
            print(json.dumps({"value": value, "unit": UNIT, "ms_per_step": total_ms / args.steps, "galaxies": B, "n": n,
                              "ok_fraction": n_ok / B, "mean_candidates": mean_cands, "launches": int(launches),
                              "flags": FLAGS_ALL, "clocks": clocks}), flush=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- end to end through the public host API: pinned host images -> rows on the host
    E = min(args.e2e_batch, B)
    xh = x[:E].cpu().pin_memory()
    pm.batch.analyse_host(xh, synth.SKY, synth.SKY_ERR, 3, FLAGS_ALL, device=dev, chunk=args.e2e_chunk)   # warm
    fence()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        rows_h = pm.batch.analyse_host(xh, synth.SKY, synth.SKY_ERR, 3, FLAGS_ALL, device=dev, chunk=args.e2e_chunk)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * E * e2e_steps / e2e_s
    assert np.array_equal(rows_h, rows[:E].cpu().numpy(), equal_nan=True), "e2e rows differ from the resident run"

    # ---- the same call fed the way the reference's frames are stored: big-endian BITPIX 16 payloads (BZERO 32768,
    # Image.py:132-134), 2 bytes per pixel over PCIe, decoded on the device (pm_fits_decode_batch).  The cutouts are
    # the same ones rounded to integer counts and clipped to the unsigned-16 range; reported beside e2e, not as it.
    del xh
    xi_h = torch.empty((E, n, n), dtype=torch.float32).pin_memory()
    pay_h = torch.empty((E, n * n * 2), dtype=torch.uint8).pin_memory()
    for s0 in range(0, E, 4096):
        xc = x[s0:min(E, s0 + 4096)].round().clamp_(0, 65535)
        xi_h[s0:s0 + xc.shape[0]].copy_(xc)
        be = (xc.to(torch.int32) - 32768).to(torch.int16)
        pay_h[s0:s0 + xc.shape[0]].copy_(be.view(torch.uint8).reshape(-1, 2).flip(1).reshape(xc.shape[0], -1))   # big-endian
        del xc, be
    want16 = pm.batch.analyse_host(xi_h, synth.SKY, synth.SKY_ERR, 3, FLAGS_ALL, device=dev, chunk=args.e2e_chunk)
    del xi_h
    kw16 = dict(bzero=32768.0, bscale=1.0, filter_size=3, flags=FLAGS_ALL, device=dev, chunk=args.e2e_payload_chunk)
    pm.batch.analyse_host_payload(pay_h, 16, n, synth.SKY, synth.SKY_ERR, **kw16)                            # warm
    fence()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        rows16 = pm.batch.analyse_host_payload(pay_h, 16, n, synth.SKY, synth.SKY_ERR, **kw16)
    torch.cuda.synchronize(dev)
    f16_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([f16_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        f16_s = float(t.item())
    f16_value = world * E * e2e_steps / f16_s
    assert np.array_equal(rows16, want16, equal_nan=True), "payload rows differ from the float32 host call"
    # ---- the decode kernel alone (HBM-bound byte work): CUDA events on the launching stream, payload > L2
    D = min(E, 8192)
    pay_d = pay_h[:D].to(dev)
    for _ in range(3):
        img_d = pm.image.decode_payload(pay_d, 16, (D, n, n), 32768.0, 1.0, torch.float32)
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    d0.record()
    for _ in range(10):
        img_d = pm.image.decode_payload(pay_d, 16, (D, n, n), 32768.0, 1.0, torch.float32)
    d1.record()
    torch.cuda.synchronize(dev)
    dec_ms = d0.elapsed_time(d1) / 10
    dec_bytes = D * n * n * 6
    del pay_d, img_d
    # ---- sky-region statistics (the per-pixel half of skybgr) on resident cutouts: CUDA events, 3 launches
    S_ = min(B, 16384)
    rng_ = np.random.default_rng(1)
    fw_, th_ = rng_.uniform(4.0, 16.0, (S_, 2)), rng_.uniform(-3.0, 3.0, S_)
    for _ in range(2):
        sky_rows = pm.skyBackground.sky_stats_batch(x[:S_], fw_, th_)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    s0.record()
    for _ in range(3):
        sky_rows = pm.skyBackground.sky_stats_batch(x[:S_], fw_, th_)
    s1.record()
    torch.cuda.synchronize(dev)
    sky_ms = s0.elapsed_time(s1) / 3
    sky_iters = float(sky_rows[:, 4].mean().item())

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    # ---- roofline (HBM) for the one kernel of the step
    peaks_path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    abytes = algorithmic_bytes(n)
    achieved = B * abytes / (kernel_ms * 1e-3) / 1e9
    traffic = dec_traffic = None
    tpath = os.path.join(REPO, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj.get("dram_bytes_per_galaxy") and tj["dram_bytes_per_galaxy"] * B
            dec_traffic = tj.get("decode16_f32_dram_bytes_per_pixel") and tj["decode16_f32_dram_bytes_per_pixel"] * D * n * n
        except Exception:
            traffic = None
    # ---- CPU baseline (bounded sample, N = 1 only)
    cpu = None
    if world == 1 and not args.no_cpu:
        procs = os.cpu_count() or 1
        S = args.cpu_sample or 8 * procs
        sample_imgs = x[:S].cpu().numpy()
        rate, dt = cpu_pool_rate(sample_imgs, procs)
        cpu = {"value": rate, "unit": UNIT, "cores": procs, "kind": "port",
               "sample": f"first {S} cutouts of the same synthetic batch, oracle port, multiprocessing.Pool({procs}), {dt:.1f} s"}
    line = {"metric": METRIC if args.n == 256 else METRIC.replace("256^2", f"{args.n}^2"), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{B} synthetic Sersic+noise {n}x{n} float32 cutouts per GPU resident in HBM "
                                   f"({B * per / 2 ** 30:.1f} GiB > L2, no flush needed), -Aall -cas +S +Gini/M20 "
                                   f"(BASELINE.json configs[2]); one CTA per galaxy, one launch per step",
                       "galaxies_per_gpu": B, "n": n, "filter_size": 3, "flags": FLAGS_ALL,
                       "parallelism": f"galaxy-sharded x{world}" + (" + all-gather of rows" if world > 1 else ""),
                       "status_ok_fraction": n_ok / B, "mean_candidates": mean_cands, "gen_seconds": gen_s,
                       "step_ms": step_ms},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_galaxy": abytes,
                         "kernel": "pm_analyse_kernel<float>", "kernel_ms": kernel_ms},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": E * per + 16 * E, "d2h_bytes_per_step": E * 128,
                    "galaxies_per_step": E, "chunk": args.e2e_chunk, "api": "pawlikmorph_lsst_b200.batch.analyse_host"},
            "e2e_fits16": {"value": f16_value, "unit": UNIT, "h2d_bytes_per_step": E * n * n * 2 + 16 * E, "d2h_bytes_per_step": E * 128,
                           "galaxies_per_step": E, "chunk": args.e2e_payload_chunk, "api": "pawlikmorph_lsst_b200.batch.analyse_host_payload",
                           "note": "same cutouts as integer counts in the FITS BITPIX 16 / BZERO 32768 wire format, decoded on the device"},
            "roofline_decode": {"bound": "hbm", "achieved": dec_bytes / (dec_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                "frac": dec_bytes / (dec_ms * 1e-3) / 1e9 / peak, "kernel": "pm_fits_decode_kernel<16, float>",
                                "kernel_ms": dec_ms, "algorithmic_bytes_per_pixel": 6, "pixels": D * n * n, "traffic": dec_traffic},
            "sky_stats": {"value": S_ / (sky_ms * 1e-3), "unit": "cutouts/s", "kernel": "pm_sky_stats_kernel<float>", "kernel_ms": sky_ms,
                          "cutouts": S_, "mean_clip_iterations": sky_iters, "per_gpu": True,
                          "note": "sky-region statistics of skyBackground._calcSkybgr with random fitted ellipses; L2/issue bound, see profiles/r1i_sky_ncu_summary.txt"},
            "gpu_launches": int(launches), "clocks": clocks}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--batch", type=int, default=int(os.environ.get("PM_BENCH_BATCH", 262144)),
                    help="galaxies resident per GPU (BASELINE.json configs[2]: 262144; reduced automatically if HBM is short)")
    ap.add_argument("--e2e-batch", type=int, default=32768)
    ap.add_argument("--e2e-chunk", type=int, default=1024)
    ap.add_argument("--e2e-payload-chunk", type=int, default=4096)
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--kernel-only", action="store_true", help="tuning: stop after the resident timed steps (short JSON line)")
    ap.add_argument("--phases", action="store_true", help="print per-phase cycle shares to stderr (tuning aid)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
