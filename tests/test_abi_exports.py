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


def test_next_row_entry_points_validate_arguments_without_a_gpu():
    """Argument validation of the handle-less entry points happens before any device work: bad arguments are
    SLB_ERR_INVALID, an empty batch is a no-op, and real work without a GPU is SLB_ERR_NO_DEVICE (never a CPU fallback)."""
    import ctypes as C
    import numpy as np
    import torch
    L = engine.lib()
    buf = np.zeros(45 * 45 + 64)
    p = buf.ctypes.data_as(C.c_void_p)
    acc = (C.c_int32 * 4)()
    assert L.slb_ekf_update(0, 3, p, p, p, p, p, 0, p, acc, None) == 0                    # empty batch
    assert L.slb_ekf_update(1, 4, p, p, p, p, p, 0, p, acc, None) == -1                   # m != 3 is not built
    assert b"m = 3" in L.slb_last_error()
    assert L.slb_ekf_predict(1, None, p, p, p, None) == -1                                 # null pointer
    assert L.slb_datamodel_safe_fuse(0, p, p, p, p, p, p, None) == 0
    assert L.slb_transform_compose(-1, p, p, p, p, p, p, None) == -1
    if not torch.cuda.is_available():
        assert L.slb_ekf_predict(1, p, p, p, p, None) == -2
        assert L.slb_datamodel_safe_fuse(1, p, p, p, p, p, p, None) == -2
        assert L.slb_deadreckon_update_pose(1, C.c_double(0.01), p, p, p, p, p, p, p, p, p, None) == -2
        assert b"no CPU fallback" in L.slb_last_error()


def test_round2_entry_points_validate_arguments_without_a_gpu():
    """The handle-taking entry points added in round 2 reject a null handle before touching the device, and the NCCL helpers
    validate their arguments (NCCL itself is resolved with dlopen only when a communicator is actually requested)."""
    import ctypes as C
    L = engine.lib()
    buf = (C.c_ubyte * 128)()
    comm = C.c_void_p()
    assert L.slb_wait(None, None) == -1
    assert L.slb_set_output_slice(None, 0, 1) == -1
    assert L.slb_check_sigma_points(None, None, None, None) == -1
    assert L.slb_gather_stats(None, None, None, None) == -1
    assert L.slb_nccl_unique_id(None) == -1
    assert L.slb_nccl_comm_init(C.byref(comm), 0, buf, 0, 0) == -1          # nranks < 1
    assert L.slb_nccl_comm_init(C.byref(comm), 2, buf, 2, 0) == -1          # rank out of range
    assert L.slb_nccl_comm_destroy(None) == 0                                # nothing to destroy
    assert L.slb_status_ex(None, None, 5, None) == -1
    assert L.slb_usckf_step_host_async(None, 4, 102, None, C.c_double(0.01), None, None, None, 0, None, None) == -1


def test_usckf_shapes_are_validated_at_create_without_a_gpu():
    """slb_create checks the device first (no CPU fallback); on a GPU box the shape check is covered by the GPU tests."""
    import ctypes as C
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = engine.SlbConfig()
    cfg.kind, cfg.batch, cfg.nk, cfg.nl = engine.KIND_USCKF, 8, 3, 12
    h = C.c_void_p()
    assert engine.lib().slb_create(C.byref(cfg), C.byref(h)) == -2
