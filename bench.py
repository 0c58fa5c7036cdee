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
    """One cutout through the oracle port: (row vector over ROW_FIELDS, packed pixelmap)."""
    import numpy as np

    import synth
    from oracle import pm_oracle as O
    row, seg, _ = O.analyse(img.astype(np.float64), synth.SKY, synth.SKY_ERR, 3, FLAGS_ALL & 15)
    return np.array([row[k] for k in ROW_FIELDS]), np.packbits(seg)


ROW_FIELDS = ("apix_col", "apix_row", "rmax", "r20", "r80", "C", "A", "Abgr", "S", "As", "As90", "gini", "m20", "object_edge",
              "status", "n_candidates")


def cpu_pool_run(images, procs):
    """The oracle port over `images` with a pool of `procs` workers (warm): galaxies/s, seconds, rows, packed pixelmaps."""
    import multiprocessing as mp

    import numpy as np
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs, initializer=_cpu_init) as pool:
        pool.map(_cpu_one, [images[0]] * procs)                           # every worker warm
        t0 = time.perf_counter()
        res = pool.map(_cpu_one, list(images), chunksize=1)
        dt = time.perf_counter() - t0
    return len(images) / dt, dt, np.stack([r for r, _ in res]), [s for _, s in res]


def sample_size(procs, requested=0):
    """Cutouts of the CPU sample: the first 512 of the resident batch when the box has the cores to do them in
    ~20 s, fewer on small hosts."""
    return requested or (512 if procs >= 12 else 32 * procs)


def first_galaxies(S, n, device_index=0):
    """The first S cutouts of rank 0's resident batch (same generator, seed and chunking as run_ours), on the host."""
    import torch

    import synth
    seed = 20261019 + 1000 * 2 + 0
    if torch.cuda.is_available():
        dev = torch.device("cuda", device_index)
        x = synth.make_batch(max(S, 2048), n, seed, device=dev, chunk=2048)[:S].cpu().numpy()     # data generation only
        return x, "first %d cutouts of the resident batch of the GPU arm (seed 20261019+2000, generated on the device)" % S, True
    return synth.make_batch(S, n, seed).numpy(), "%d cutouts of the same generator on the host (no CUDA device: other random stream)" % S, False


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    procs = os.cpu_count() or 1
    S = sample_size(procs, args.cpu_sample)
    images, sample_desc, same = first_galaxies(S, args.n)
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
    sample = f"{sample_desc}, every step, multiprocessing.Pool({procs})"
    line = {"impl": "reference", "metric": metric_name(args.n), "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.n) + "; bounded sample per step", "n": args.n, "filter_size": 3, "flags": FLAGS_ALL & 15,
                       "same_galaxies_as_gpu_arm": same},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def metric_name(n):
    return METRIC if n == 256 else METRIC.replace("256^2", f"{n}^2")


def workload_name(n):
    cfg = {256: "BASELINE.json configs[2]", 140: "the 140^2 size of BASELINE.json configs[1] with every statistic", 512: "the 512^2 size of BASELINE.json configs[4]"}
    return (f"synthetic Sersic+noise {n}x{n} float32 cutouts, -Aall -cas +S +Gini/M20 (flags 15)"
            + (f" = {cfg[n]}" if n in cfg else ""))


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
    # the kernel writes its rows straight into this rank's slot of the gather buffer: in-place all-gather
    table = torch.empty((B * world, 16), dtype=torch.float64, device=dev)
    rows = table[rank * B:(rank + 1) * B]
    segmap = torch.empty((B, n, n), dtype=torch.uint8, device=dev)      # the pixelmap is an output of the path
    stream = torch.cuda.current_stream(dev)

    def step():
        pm.batch.analyse_batch(x, sky, err, 3, FLAGS_ALL, rows_out=rows, segmap_out=segmap, stream=stream)
        if world > 1:
            dist.all_gather_into_tensor(table, rows)           # in place: `rows` is this rank's slot of `table`

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

    # ---- multi-GPU: the gathered table holds every rank's rows in rank order, bit for bit, and a fixed global batch
    # sharded over the ranks gives the same table whatever the world size (no cross-galaxy term: main.py:161-174)
    dist_info = None
    if world > 1:
        mine = rows.clone()
        table.zero_()
        rows.copy_(mine)
        dist.all_gather_into_tensor(table, rows)
        torch.cuda.synchronize(dev)
        assert torch.equal(table[rank * B:(rank + 1) * B], mine), "gathered table: own slot differs"
        chk = torch.stack([table[r * B:(r + 1) * B].nan_to_num(nan=-7.0).sum(dim=0) for r in range(world)])   # [W, 16] per-slot checksums
        ref = chk.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(chk, ref), "gathered table differs between ranks"
    G = 8192                                             # fixed global batch: identical table for N = 1, 2, 4, 8
    if G % world == 0:
        import hashlib
        xg = synth.make_batch(G, n, 20261019 + 7777, device=dev, chunk=2048)
        lo, hi = pm.batch.shard_bounds(G, world, rank)
        tg = torch.empty((G, 16), dtype=torch.float64, device=dev)
        pm.batch.analyse_batch(xg[lo:hi].contiguous(), synth.SKY, synth.SKY_ERR, 3, FLAGS_ALL, rows_out=tg[lo:hi], stream=stream)
        if world > 1:
            dist.all_gather_into_tensor(tg, tg[lo:hi])
        torch.cuda.synchronize(dev)
        dist_info = {"fixed_batch": G, "seed": "20261019+7777", "table_sha256": hashlib.sha256(tg.cpu().numpy().tobytes()).hexdigest(),
                     "note": "rows of the same 8192 cutouts sharded over the ranks and all-gathered: the hash must not depend on --gpus"}
        del xg, tg
    # ---- BASELINE.json configs[3]: 4 194 304 cutouts of 256^2 strong-scaled over the ranks, generated on the device in
    # chunks (never materialised: SURVEY 8d); only the analysis of each chunk is timed (CUDA events), generation is not
    strong = None
    if world > 1 and n == 256 and not args.no_strong:
        total_g = 4194304
        per_rank = total_g // world
        done_g, ms = 0, 0.0
        k = 0
        while done_g < per_rank:
            nb = min(B, per_rank - done_g)
            if k > 0:
                x[:nb] = synth.make_batch(nb, n, 20261019 + 4000 + 64 * rank + k, device=dev, chunk=2048)
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            pm.batch.analyse_batch(x[:nb], sky[:nb], err[:nb], 3, FLAGS_ALL, rows_out=rows[:nb], segmap_out=segmap[:nb], stream=stream)
            b_.record(stream)
            torch.cuda.synchronize(dev)
            ms += a.elapsed_time(b_)
            done_g += nb
            k += 1
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        strong = {"galaxies_total": per_rank * world, "galaxies_per_gpu": per_rank, "chunks_per_gpu": k, "value": per_rank * world / (float(t.item()) * 1e-3),
                  "unit": UNIT, "seconds": float(t.item()) * 1e-3, "scaling": "strong",
                  "note": "BASELINE.json configs[3]; chunks regenerated on the device between timed launches, rows stay on the device"}
        x = synth.make_batch(B, n, 20261019 + 1000 * 2 + rank, device=dev, chunk=2048)      # the resident batch again
        pm.batch.analyse_batch(x, sky, err, 3, FLAGS_ALL, rows_out=rows, segmap_out=segmap, stream=stream)
        torch.cuda.synchronize(dev)

    # ---- pinned host -> device bandwidth of this rank (what bounds the end-to-end numbers below)
    probe_h = [torch.empty(512 * 2 ** 20, dtype=torch.uint8).pin_memory() for _ in range(2)]
    probe_d = [torch.empty_like(t, device=dev) for t in probe_h]
    pstreams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]          # two copy streams, like analyse_host
    for k in range(2):
        with torch.cuda.stream(pstreams[k]):
            probe_d[k].copy_(probe_h[k], non_blocking=True)
    fence()
    t0p = time.perf_counter()
    for it in range(4):
        for k in range(2):
            with torch.cuda.stream(pstreams[k]):
                probe_d[k].copy_(probe_h[k], non_blocking=True)
    for st_ in pstreams:
        st_.synchronize()
    h2d_gbs = 8 * probe_h[0].numel() / (time.perf_counter() - t0p) / 1e9
    h2d_all = torch.tensor([h2d_gbs], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(h2d_all, op=dist.ReduceOp.SUM)      # all ranks copy at the same time: the aggregate the host sustains
    h2d_sum = float(h2d_all.item())
    del probe_h, probe_d

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
    # ---- the pre-processing rows of SURVEY.md §8(f) on the same resident cutouts (rank 0; CUDA events, 1 warm-up + 2 timed)
    prep = None
    if rank == 0 and not args.no_configs:
        def timed(fn, reps=2):
            fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            a.record()
            for _ in range(reps):
                out = fn()
            b.record()
            torch.cuda.synchronize(dev)
            return a.elapsed_time(b) / reps, out
        P_ = min(B, 8192)
        xp = x[:P_]
        img_bytes = P_ * n * n * 4
        clip_ms, clip = timed(lambda: pm.skyBackground.clipped_stats_batch(xp, 3.0, 5))
        fit_ms, fit = timed(lambda: pm.skyBackground.gaussfit_batch(xp, stats=clip))
        bgr_ms, bgr = timed(lambda: pm.skyBackground.skybgr_batch(xp))
        star_ms, star = timed(lambda: pm.imageutils.maskstars_batch(xp, seed=1))
        flow_ms, flow = timed(lambda: pm.main._device_group(xp, None, None, None, FLAGS_ALL, 3, True, 1), reps=1)
        flow_ok = float((flow["rows"][:, pm._abi.ROW_FIELDS.index("status")] == 0).mean())
        # fitSersic on the sky-subtracted cutouts with the widths / angle of the Gaussian fit and the apix of the rows
        S_n = min(P_, 2048)
        fitn = fit[:S_n].cpu().numpy()
        ser_in = xp[:S_n] - float(synth.SKY)
        ser_cen = rows[:S_n, :2].cpu().numpy()
        ser_fw = 2.3548200450309493 * np.abs(fitn[:, 4:6])
        ser_ms, ser = timed(lambda: pm.sersic.fitSersic_batch(ser_in, ser_cen, ser_fw, np.abs(fitn[:, 6])), reps=1)
        gbs = lambda nbytes, ms: nbytes / (ms * 1e-3) / 1e9
        prep = {"cutouts": P_, "n": n, "per_gpu": True,
                "clipped_stats": {"value": P_ / (clip_ms * 1e-3), "unit": "cutouts/s", "kernel": "pm_clipstats_kernel<float>", "kernel_ms": clip_ms,
                                  "algorithmic_gbs": gbs(img_bytes, clip_ms), "mean_iterations": float(clip[:, 4].mean().item())},
                "gaussfit": {"value": P_ / (fit_ms * 1e-3), "unit": "cutouts/s", "kernel": "pm_gaussfit_kernel<float>", "kernel_ms": fit_ms,
                             "algorithmic_gbs": gbs(img_bytes, fit_ms), "mean_iterations": float(fit[:, 8].mean().item()),
                             "converged_fraction": float((fit[:, 9] == 0).double().mean().item())},
                "skybgr": {"value": P_ / (bgr_ms * 1e-3), "unit": "cutouts/s", "ms": bgr_ms, "algorithmic_gbs": gbs(img_bytes, bgr_ms),
                           "ok_fraction": float((np.asarray(bgr["status"]) == 0).mean()),
                           "note": "whole skyBackground.skybgr flow: clipped statistics, Gaussian fit, source check, sky-region statistics"},
                "maskstarsSEG": {"value": P_ / (star_ms * 1e-3), "unit": "cutouts/s", "ms": star_ms, "algorithmic_gbs": gbs(2 * img_bytes, star_ms),
                                 "mean_segments": float(torch.as_tensor(star["info"])[:, 0].double().mean().item()),
                                 "note": "two clipped-statistics launches + the smooth/threshold/label/replace kernel (reads and writes the cutout)"}}
        prep["fitSersic"] = {"value": S_n / (ser_ms * 1e-3), "unit": "cutouts/s", "ms": ser_ms, "cutouts": S_n,
                             "ok_fraction": float(np.isin(ser["status"], (0, 1)).mean()), "mean_iterations": float(np.nanmean(ser["iterations"])),
                             "note": "exact elliptical aperture sums (walk + bisection in lock-step, one launch per step) + Sersic2D fit"}
        prep["reference_default_flow"] = {"value": P_ / (flow_ms * 1e-3), "unit": "galaxies/s", "ms": flow_ms, "ok_fraction": flow_ok,
                                          "note": "main.py:349-495 as upstream runs it by default, on resident cutouts: skybgr, pixelmap with the sky in, "
                                                  "img - sky, maskstarsSEG, the statistics on the cleaned image (rows read back to the host)"}
        del clip, fit, bgr, star, flow, ser, ser_in

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
            per_n = tj.get("dram_bytes_per_galaxy_by_n", {}).get(str(n))          # ncu --set full captures, per cutout size
            traffic = per_n * B if per_n else None
            dec_traffic = tj.get("decode16_f32_dram_bytes_per_pixel") and tj["decode16_f32_dram_bytes_per_pixel"] * D * n * n
        except Exception:
            traffic = None
    # ---- CPU baseline + parity sample (N = 1 only): the oracle port over the FIRST S cutouts of the resident batch, timed
    # on the host cores, and its rows / pixelmaps compared with what the kernel wrote for the same cutouts
    cpu = parity = None
    if world == 1 and not args.no_cpu:
        from tests.golden_util import FLOATS, compare_row
        procs = os.cpu_count() or 1
        S = min(B, sample_size(procs, args.cpu_sample))
        sample_imgs = x[:S].cpu().numpy()
        rate, dt, want_rows, want_segs = cpu_pool_run(sample_imgs, procs)
        cpu = {"value": rate, "unit": UNIT, "cores": procs, "kind": "port",
               "sample": f"first {S} cutouts of the resident batch, oracle port, multiprocessing.Pool({procs}), {dt:.1f} s"}
        got_rows, got_segs = rows[:S].cpu().numpy(), segmap[:S].cpu().numpy()
        seg_equal = all(np.array_equal(np.packbits(got_segs[i]), want_segs[i]) for i in range(S))
        errs, max_rel = [], 0.0
        for i in range(S):
            g_, w_ = dict(zip(ROW_FIELDS, got_rows[i])), dict(zip(ROW_FIELDS, want_rows[i]))
            errs += compare_row(g_, w_, 1e-9, f"[{i}] ")
            if g_["n_candidates"] != w_["n_candidates"]:
                errs.append(f"[{i}] n_candidates")
            for k_ in FLOATS:
                if w_[k_] != -99.0 and np.isfinite(w_[k_]) and np.isfinite(g_[k_]) and w_[k_] != 0.0:
                    max_rel = max(max_rel, abs(g_[k_] - w_[k_]) / abs(w_[k_]))
        parity = {"n": S, "segmap_equal": bool(seg_equal), "rows_within_1e-9": not errs, "max_rel": max_rel,
                  "mismatches": errs[:5], "checker": "oracle/pm_oracle.py (pinned by the goldens of the unmodified reference)"}
    # ---- the other BASELINE.json configurations that fit one GPU (info keys; N = 1 only, after the headline run)
    other = {}
    if world == 1 and not args.no_configs and n == 256:
        del x, segmap, table, rows
        torch.cuda.empty_cache()

        def run_cfg(nn, Bc, flags, seed, with_segmaps):
            xc = synth.make_batch(Bc, nn, seed, device=dev, chunk=1024)
            rc_ = torch.empty((Bc, 16), dtype=torch.float64, device=dev)
            sc_ = torch.empty((Bc, nn, nn), dtype=torch.uint8, device=dev)
            kw = {}
            if with_segmaps:        # configs[4]: precomputed segmaps are an INPUT (main.py:428-439); made here by the pixelmap pass
                pm.batch.analyse_batch(xc, synth.SKY, synth.SKY_ERR, 3, flags | pm._abi.PM_STOP_AFTER_PIXELMAP, rows_out=rc_, segmap_out=sc_, stream=stream)
                kw = dict(segmap_in=sc_)
            else:
                kw = dict(segmap_out=sc_)
            for _ in range(3):
                pm.batch.analyse_batch(xc, synth.SKY, synth.SKY_ERR, 3, flags, rows_out=rc_, stream=stream, **kw)
            c0_, c1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            c0_.record(stream)
            for _ in range(3):
                pm.batch.analyse_batch(xc, synth.SKY, synth.SKY_ERR, 3, flags, rows_out=rc_, stream=stream, **kw)
            c1_.record(stream)
            torch.cuda.synchronize(dev)
            ms = c0_.elapsed_time(c1_) / 3
            ab = algorithmic_bytes(nn)
            ach = Bc * ab / (ms * 1e-3) / 1e9
            tn = None
            try:
                tn = json.load(open(tpath)).get("dram_bytes_per_galaxy_by_n", {}).get(str(nn))
            except Exception:
                pass
            ok = float((rc_[:, 14] == 0).double().mean().item())
            return {"value": Bc / (ms * 1e-3), "unit": UNIT, "galaxies": Bc, "n": nn, "flags": flags, "ms_per_step": ms, "status_ok_fraction": ok,
                    "segmap": "supplied (u8 input)" if with_segmaps else "computed (u8 output)",
                    "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                 "algorithmic_bytes_per_galaxy": ab, "traffic": tn * Bc if tn else None}}
        other["config2_140"] = dict(run_cfg(140, 65536, 3, 20261019 + 1000 * 1, False),
                                    workload="BASELINE.json configs[1]: 65536 cutouts of 140^2, -A -As (flags 3)")
        other["config5_512"] = dict(run_cfg(512, 8192, 15, 20261019 + 1000 * 4, True),
                                    workload="BASELINE.json configs[4]: 512^2 cutouts with precomputed segmaps, every statistic (flags 15), "
                                             "8192 resident cutouts (8 GiB); one CTA per cutout, the same kernel (no tiled variant)")
    line = {"metric": metric_name(n), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{B} per GPU resident in HBM ({B * per / 2 ** 30:.1f} GiB > L2, no flush needed): " + workload_name(n)
                                   + "; one CTA per galaxy, one launch per step",
                       "galaxies_per_gpu": B, "n": n, "filter_size": 3, "flags": FLAGS_ALL,
                       "parallelism": f"galaxy-sharded x{world}" + (" + all-gather of rows" if world > 1 else ""),
                       "status_ok_fraction": n_ok / B, "mean_candidates": mean_cands, "gen_seconds": gen_s,
                       "step_ms": step_ms},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_galaxy": abytes,
                         "kernel": "pm_analyse_kernel<float>", "kernel_ms": kernel_ms},
            "cpu_baseline": cpu, "parity_sample": parity,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": E * per + 16 * E, "d2h_bytes_per_step": E * 128,
                    "galaxies_per_step": E, "chunk": args.e2e_chunk, "api": "pawlikmorph_lsst_b200.batch.analyse_host",
                    "h2d_gbs_per_rank": h2d_gbs, "h2d_gbs_all_ranks": h2d_sum,
                    "fraction_of_h2d_bound": e2e_value * (per + 16) / 1e9 / h2d_sum if h2d_sum > 0 else None,
                    "note": "pinned host -> device copy bandwidth measured in this run with every rank copying at once; the float32 "
                            "payload call is bound by it"},
            "e2e_fits16": {"value": f16_value, "unit": UNIT, "h2d_bytes_per_step": E * n * n * 2 + 16 * E, "d2h_bytes_per_step": E * 128,
                           "galaxies_per_step": E, "chunk": args.e2e_payload_chunk, "api": "pawlikmorph_lsst_b200.batch.analyse_host_payload",
                           "fraction_of_h2d_bound": f16_value * (n * n * 2 + 16) / 1e9 / h2d_sum if h2d_sum > 0 else None,
                           "note": "same cutouts as integer counts in the FITS BITPIX 16 / BZERO 32768 wire format, decoded on the device"},
            "roofline_decode": {"bound": "hbm", "achieved": dec_bytes / (dec_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                "frac": dec_bytes / (dec_ms * 1e-3) / 1e9 / peak, "kernel": "pm_fits_decode_kernel<16, float>",
                                "kernel_ms": dec_ms, "algorithmic_bytes_per_pixel": 6, "pixels": D * n * n, "traffic": dec_traffic},
            "sky_stats": {"value": S_ / (sky_ms * 1e-3), "unit": "cutouts/s", "kernel": "pm_sky_stats_kernel<float>", "kernel_ms": sky_ms,
                          "cutouts": S_, "mean_clip_iterations": sky_iters, "per_gpu": True,
                          "note": "sky-region statistics of skyBackground._calcSkybgr with random fitted ellipses; instruction-issue bound, passes through L2, see profiles/r2u_sky_ncu_summary.txt"},
            "gpu_launches": int(launches), "clocks": clocks}
    if prep:
        line["prep"] = prep
    if dist_info:
        line["distributed_check"] = dist_info
    if strong:
        line["config4_strong_4M"] = strong
    line.update(other)
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
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaled 4 M stream of --gpus N > 1")
    ap.add_argument("--no-configs", action="store_true", help="skip the 140^2 / 512^2 configuration runs")
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
