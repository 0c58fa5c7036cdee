"""In-tree build of libpawlik_b200.so with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
# PM_LIB_VARIANT=<name> selects an experimental build (tuning runs only) compiled with the extra defines of
# PM_NVCC_DEFINES (e.g. "PM_FRONT_MIN_CTAS=3 PM_TAILCTA_MIN_CTAS=3"); the library is then libpawlik_b200_<name>.so
VARIANT = os.environ.get("PM_LIB_VARIANT", "")
SO = os.path.join(HERE, "libpawlik_b200.so" if not VARIANT else f"libpawlik_b200_{VARIANT}.so")
SOURCES = [os.path.join(HERE, "csrc", f) for f in ("pm_capi.cu", "pm_tail.cu", "pm_k_mono.cu", "pm_k_front.cu", "pm_k_minapix.cu", "pm_k_tailcta.cu", "pm_k_prep.cu", "pm_k_stats.cu", "pm_stats.cuh", "pm_prep.cuh", "pm_launch.cuh", "pm_kernels.cuh", "pm_block.cuh", "pm_platform.cuh", "pm_fits.cuh")] + \
    [os.path.join(REPO, "include", "pawlik_b200.h")]


def nvcc_path():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def up_to_date():
    return os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(s) for s in SOURCES)


def build(force=False, verbose=False):
    """The translation units (pm_capi.cu: the C ABI; pm_k_*.cu: one CTA-group kernel each; pm_tail.cu: the warp-group tail
    kernel) are compiled side by side, then linked into one shared object."""
    if not force and up_to_date():
        return SO
    from concurrent.futures import ThreadPoolExecutor
    objdir = os.path.join(HERE, "csrc", "_obj" + (("_" + VARIANT) if VARIANT else ""))
    os.makedirs(objdir, exist_ok=True)
    common = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-I" + os.path.join(REPO, "include")]
    for d in os.environ.get("PM_NVCC_DEFINES", "").split():
        common.insert(1, "-D" + d)
    if verbose:
        common[1:1] = ["-Xptxas", "-v"]
    units = ["pm_capi.cu", "pm_k_mono.cu", "pm_k_front.cu", "pm_k_minapix.cu", "pm_k_tailcta.cu", "pm_tail.cu", "pm_k_prep.cu", "pm_k_stats.cu"]

    def one(u):
        o = os.path.join(objdir, u.replace(".cu", ".o"))
        r = subprocess.run(common + ["-c", os.path.join(HERE, "csrc", u), "-o", o], capture_output=True, text=True)
        return o, r

    with ThreadPoolExecutor(len(units)) as ex:
        res = list(ex.map(one, units))
    for o, r in res:
        if verbose or r.returncode:
            print(r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError(f"nvcc failed for {o}")
    subprocess.run([nvcc_path(), "-shared", "-o", SO] + [o for o, _ in res], check=True)
    return SO


if __name__ == "__main__":
    print(build(force=True, verbose=True))
