"""CPU-side checks of the drop-in boundary: the CUDA library builds/loads here (nvcc cross-compiles),
exports every symbol include/slb.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import os
import re

import pytest

from slam_localization_b200 import build as slb_build
from slam_localization_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "slb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(slb_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    slb_build.build()
    L = engine.lib()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), "libslb.so does not export %s declared in include/slb.h" % n
    assert sorted(engine.EXPORTS) == names
    assert L.slb_version() == 100


def test_product_library_does_not_link_the_oracle():
    import subprocess
    out = subprocess.run(["ldd", engine.LIB_PATH], capture_output=True, text=True).stdout
    assert "libslo" not in out
    syms = subprocess.run(["nm", "-D", engine.LIB_PATH], capture_output=True, text=True).stdout
    assert "slo_" not in syms


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(engine.SlbError):
        engine.Ukf(8)
    # the raw C ABI refuses too
    import ctypes as C
    cfg = engine.SlbConfig()
    cfg.kind, cfg.layout, cfg.batch = engine.KIND_UKF, engine.LAYOUT_MTK9, 8
    h = C.c_void_p()
    rc = engine.lib().slb_create(C.byref(cfg), C.byref(h))
    assert rc == -2 and b"no CPU fallback" in engine.lib().slb_last_error()
