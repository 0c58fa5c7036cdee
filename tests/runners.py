"""Two ways of calling the C ABI with numpy data, with one result format, so that every parity
check in tests/parity_checks.py runs unchanged against
  * the CUDA library on a B200 (`-m gpu`), and
  * the emulator build of the same kernel sources on the CPU (`-m "not gpu"`, tests/emul)."""
import numpy as np


def emul_runner(images, sky, sky_err, fs=3, flags=15, **ov):
    from tests import emul_harness as H
    return H.analyse(images, sky, sky_err, fs, flags, **ov)


def emul_runner_pruned(images, sky, sky_err, fs=3, flags=15, **ov):
    """Without the per-candidate output: minapix then takes its exact-pruning path."""
    from tests import emul_harness as H
    return H.analyse(images, sky, sky_err, fs, flags, want=("segmap",), **ov)


def gpu_runner_pruned(images, sky, sky_err, fs=3, flags=15, **ov):
    return gpu_runner(images, sky, sky_err, fs, flags, _cand_cap=0, **ov)


def gpu_runner(images, sky, sky_err, fs=3, flags=15, _cand_cap=512, **ov):
    import torch

    import pawlikmorph_lsst_b200 as pm
    x = torch.from_numpy(np.ascontiguousarray(images)).cuda()
    conv = {}
    for k, v in ov.items():
        if v is not None:
            conv[k] = torch.from_numpy(np.ascontiguousarray(np.asarray(v))).cuda()
    B = x.shape[0]
    sky = torch.from_numpy(np.broadcast_to(np.asarray(sky, np.float64), (B,)).copy())
    sky_err = torch.from_numpy(np.broadcast_to(np.asarray(sky_err, np.float64), (B,)).copy())
    r = pm.batch.analyse_batch(x, sky, sky_err, fs, flags, want_segmap=True, want_sorted=True, want_apermask=True,
                               cand_cap=_cand_cap, **conv)
    torch.cuda.synchronize()
    return dict(rows=r.rows.cpu().numpy(), segmap=r.segmap.cpu().numpy(), sorted_idx=r.sorted_idx.cpu().numpy(),
                n_positive=r.n_positive.cpu().numpy(), apermask=r.apermask.cpu().numpy(),
                cand_a=r.cand_a.cpu().numpy() if r.cand_a is not None else None)
