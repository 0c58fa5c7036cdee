"""The C ABI: the built CUDA library loads here (no GPU needed to load it) and exports every symbol
include/pawlik_b200.h declares; host-only entry points behave; the product never imports the oracle."""
import ctypes
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(REPO, "include", "pawlik_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pm_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_bound_functions():
    import pawlikmorph_lsst_b200 as pm
    assert set(_header_functions()) == set(pm._abi.EXPORTS)


def test_library_builds_loads_and_exports_every_symbol():
    import pawlikmorph_lsst_b200 as pm
    so = pm.build.build()
    L = ctypes.CDLL(so)
    for name in _header_functions():
        assert hasattr(L, name), name
    L = pm._lib.lib()
    assert L.pm_abi_version() == 2
    assert L.pm_workspace_bytes(0, 256, 0) == 0

    def want(B, n):          # pm_capi.cu make_plan(): counters | big list | per-group scratch | hand-over slots of a chunk
        up = lambda v: (v + 255) // 256 * 256
        stride = up(256 + 2 * n * ((n + 31) // 32) * 4 + 4 * (n * n // 2 + 2048))
        chunk = min(B, 49152, max(1024, (4 << 30) // stride))
        scratch = max(min(B, 592) * (64 * n * n + 65536), min(B, 640) * (up(8 * n * n) + 131072), min(B, 5120) * (up(6 * n * n) + 131072))
        return up(256 + 4 * chunk) + up(scratch) + chunk * stride
    for B, n in ((1, 256), (10 ** 6, 256), (5000, 140), (3, 512), (10 ** 5, 640)):
        assert L.pm_workspace_bytes(B, n, 0) == want(B, n), (B, n)
    assert L.pm_workspace_bytes(1, 10 ** 4, 0) == 0          # n > PM_MAX_N


def test_row_struct_is_16_doubles():
    import pawlikmorph_lsst_b200 as pm
    src = open(os.path.join(REPO, "include", "pawlik_b200.h")).read()
    body = re.search(r"typedef struct \{(.*?)\} pm_row_t;", src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = [n.strip() for decl in re.findall(r"double ([^;]+);", body) for n in decl.split(",")]
    assert tuple(names) == pm._abi.ROW_FIELDS and len(names) == 16


def test_product_never_imports_the_oracle():
    pkg = os.path.join(REPO, "pawlikmorph-lsst_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "oracle" not in text.replace("the oracle", ""), os.path.join(root, f)


def test_no_cpu_fallback_without_cuda():
    import torch

    import pawlikmorph_lsst_b200 as pm
    if torch.cuda.is_available():
        pytest.skip("GPU box")
    with pytest.raises(TypeError):
        pm.batch.analyse_batch(torch.zeros((1, 64, 64)), 0.0, 0.0, 3)
    with pytest.raises(RuntimeError):
        pm.pixmap.calcRmax(__import__("numpy").ones((16, 16)), __import__("numpy").ones((16, 16)))
