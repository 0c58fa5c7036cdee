"""In-tree build of libpawlik_b200.so with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
# PM_LIB_VARIANT=<threads>x<ctas> selects an experimental build (tuning runs only), e.g. 256x2
VARIANT = os.environ.get("PM_LIB_VARIANT", "")
SO = os.path.join(HERE, "libpawlik_b200.so" if not VARIANT else f"libpawlik_b200_{VARIANT}.so")
SOURCES = [os.path.join(HERE, "csrc", f) for f in ("pm_capi.cu", "pm_kernels.cuh", "pm_block.cuh", "pm_platform.cuh", "pm_fits.cuh", "pm_sky.cuh")] + \
    [os.path.join(REPO, "include", "pawlik_b200.h")]


def nvcc_path():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def up_to_date():
    return os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(s) for s in SOURCES)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return SO
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-I" + os.path.join(REPO, "include"),
           os.path.join(HERE, "csrc", "pm_capi.cu"), "-o", SO]
    if VARIANT:
        t, k = VARIANT.split("x")
        cmd[1:1] = [f"-DPM_THREADS={int(t)}", f"-DPM_MIN_CTAS={int(k)}"]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.run(cmd, check=True)
    return SO


if __name__ == "__main__":
    print(build(force=True, verbose=True))
