"""ctypes view of include/pawlik_b200.h (constants, structs, prototypes).

Pure declarations: nothing is loaded here.  `_lib.py` binds these to the CUDA
library (the product); the CPU test-suite binds them to the emulator build of the
same sources (tests/emul) to exercise the kernel logic without a GPU.
"""
import ctypes as C

PM_ABI_VERSION = 2
PM_MAX_N = 640
PM_F32, PM_F64 = 0, 1
PM_OK, PM_EINVAL, PM_ENOSPC, PM_ECUDA, PM_ENODEV = 0, -1, -2, -3, -4
PM_DO_A, PM_DO_AS, PM_DO_CAS, PM_DO_S, PM_DO_ALL = 1, 2, 4, 8, 15
PM_FORCE_EXACT_FILTER = 0x100
PM_STOP_AFTER_PIXELMAP = 0x200
PM_STOP_AFTER_RMAX = 0x400
PM_A_ANGLE90 = 0x800
PM_NO_NOISECORRECT = 0x1000
PM_L2_PREFETCH = 0x2000
PM_NO_PRUNE = 0x4000
PM_PATH_SPLIT = 0x8000
PM_TAIL_WARP = 0x10000
PM_RADIX_SORT = 0x20000
PM_PHASE_ROWS = 8192
PM_ST_OK, PM_ST_FAINT_CENTRE, PM_ST_NO_CANDIDATE, PM_ST_BAD_BRACKET, PM_ST_EMPTY_SEGMAP, PM_ST_WORKSPACE = range(6)

ROW_FIELDS = ("apix_col", "apix_row", "rmax", "r20", "r80", "C", "A", "Abgr", "S", "As",
              "As90", "gini", "m20", "object_edge", "status", "n_candidates")
ROW_DOUBLES = len(ROW_FIELDS)


class Overrides(C.Structure):
    _fields_ = [("segmap_in", C.c_void_p), ("starmask", C.c_void_p), ("apermask_in", C.c_void_p),
                ("centroid_in", C.c_void_p), ("radius_in", C.c_void_p), ("r20_in", C.c_void_p),
                ("threshold_in", C.c_void_p), ("smooth_sky_in", C.c_void_p)]


class Outputs(C.Structure):
    _fields_ = [("segmap_out", C.c_void_p), ("sorted_idx_out", C.c_void_p), ("n_positive_out", C.c_void_p),
                ("apermask_out", C.c_void_p), ("cand_a_out", C.c_void_p), ("cand_cap", C.c_int32),
                ("phase_cycles_out", C.c_void_p)]


EXPORTS = ("pm_abi_version", "pm_device_info", "pm_last_cuda_error", "pm_launch_count",
           "pm_workspace_bytes", "pm_analyse_batch", "pm_aperpixmap_batch", "pm_fits_decode_batch",
           "pm_fits_decode_window_batch", "pm_sky_stats_batch", "pm_clipped_stats_batch", "pm_clipped_stats2_batch", "pm_stats_plan", "pm_gaussfit_batch", "pm_ellipse_sum_batch", "pm_sersicfit_batch",
           "pm_maskstars_workspace_bytes", "pm_maskstars_batch")
PM_CLIP_COLS, PM_GAUSS_COLS, PM_STAR_COLS = 10, 10, 4
CLIP_FIELDS = ("count", "mean", "median", "std", "clip_iterations", "clipped_count", "clipped_mean", "clipped_median", "clipped_std", "max")
SERSIC_FIELDS = ("amplitude", "r_eff", "n", "x_0", "y_0", "ellip", "theta", "cost", "iterations", "status")
GAUSS_FIELDS = ("height", "amplitude", "xo", "yo", "sigma_x", "sigma_y", "theta", "cost", "iterations", "status")
PM_SKY_COLS = 12
SKY_FIELDS = ("count", "mean", "median", "std", "clip_iterations", "clipped_count", "clipped_mean", "clipped_median", "clipped_std",
              "sky", "sky_err", "status")


def declare(lib):
    """Attach prototypes to a loaded library and return it."""
    lib.pm_abi_version.restype = C.c_int
    lib.pm_abi_version.argtypes = []
    lib.pm_device_info.restype = C.c_int
    lib.pm_device_info.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_size_t)]
    lib.pm_last_cuda_error.restype = C.c_char_p
    lib.pm_last_cuda_error.argtypes = []
    lib.pm_launch_count.restype = C.c_uint64
    lib.pm_launch_count.argtypes = []
    lib.pm_workspace_bytes.restype = C.c_size_t
    lib.pm_workspace_bytes.argtypes = [C.c_int64, C.c_int, C.c_int]
    lib.pm_analyse_batch.restype = C.c_int
    lib.pm_analyse_batch.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                     C.c_uint32, C.POINTER(Overrides), C.POINTER(Outputs), C.c_void_p,
                                     C.c_void_p, C.c_size_t, C.c_void_p]
    lib.pm_aperpixmap_batch.restype = C.c_int
    lib.pm_aperpixmap_batch.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
    lib.pm_fits_decode_batch.restype = C.c_int
    lib.pm_fits_decode_batch.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int64, C.c_void_p, C.c_int, C.c_void_p]
    lib.pm_sky_stats_batch.restype = C.c_int
    lib.pm_sky_stats_batch.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pm_fits_decode_window_batch.restype = C.c_int
    lib.pm_fits_decode_window_batch.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int64, C.c_int, C.c_int, C.c_void_p,
                                                C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    lib.pm_clipped_stats_batch.restype = C.c_int
    lib.pm_clipped_stats_batch.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
    lib.pm_ellipse_sum_batch.restype = C.c_int
    lib.pm_ellipse_sum_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pm_sersicfit_batch.restype = C.c_int
    lib.pm_sersicfit_batch.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pm_clipped_stats2_batch.restype = C.c_int
    lib.pm_clipped_stats2_batch.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pm_stats_plan.restype = C.c_int
    lib.pm_stats_plan.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.pm_gaussfit_batch.restype = C.c_int
    lib.pm_gaussfit_batch.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pm_maskstars_workspace_bytes.restype = C.c_size_t
    lib.pm_maskstars_workspace_bytes.argtypes = [C.c_int64, C.c_int]
    lib.pm_maskstars_batch.restype = C.c_int
    lib.pm_maskstars_batch.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    return lib


class PawlikError(RuntimeError):
    pass


def check(lib, rc, what):
    if rc != PM_OK:
        msg = lib.pm_last_cuda_error().decode("utf-8", "replace")
        names = {PM_EINVAL: "PM_EINVAL", PM_ENOSPC: "PM_ENOSPC", PM_ECUDA: "PM_ECUDA", PM_ENODEV: "PM_ENODEV"}
        if rc == PM_EINVAL:
            raise ValueError(f"{what}: {msg}")
        raise PawlikError(f"{what} failed with {names.get(rc, rc)}: {msg}")
