"""GPU parity: DataModel covariance fusion (config 5 of BASELINE.json) through the C ABI vs the oracle.
The kernel mirrors the oracle operation by operation without FMA contraction, so the bar is bit-exact,
even on the cond ~1e7 covariances the config prescribes."""
import numpy as np
import pytest

from slam_localization_b200 import engine, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("d", [3, 6])
@pytest.mark.parametrize("n", [1, 31, 32, 33, 1000, 4097])
def test_fusion_bit_exact(slo, d, n):
    sc = synth.fusion_scenario(n, d=d)
    xo, Co = engine.DataModel.fuse(sc["x1"], sc["C1"], sc["x2"], sc["C2"])
    xr, Cr = slo.datamodel(0, sc["x1"], sc["C1"], sc["x2"], sc["C2"])
    np.testing.assert_array_equal(xo.numpy(), xr)
    np.testing.assert_array_equal(Co.numpy(), Cr)


def test_fusion_reference_fixture(slo):
    fx = synth.datamodel_fixture()              # test/DataModelUnitTest.cpp:32-35,66-67
    xo, Co = engine.DataModel.fuse_host(fx["x1"], fx["C1"], fx["x2"], fx["C2"])
    np.testing.assert_allclose(xo[0], [0.0146635, 0.0011758085, -0.0187294], rtol=1e-9)
    np.testing.assert_allclose(Co, fx["C_expected"], rtol=1e-12, atol=1e-26)
    xr, Cr = slo.datamodel(0, fx["x1"], fx["C1"], fx["x2"], fx["C2"])
    np.testing.assert_array_equal(xo, xr)
    np.testing.assert_array_equal(Co, Cr)


@pytest.mark.parametrize("d", [3, 6])
def test_addsub(slo, d):
    sc = synth.fusion_scenario(777, d=d)
    for sign in (+1, -1):
        xo, Co = engine.DataModel.addsub(sign, sc["x1"], sc["C1"], sc["x2"], sc["C2"])
        xr, Cr = slo.datamodel(sign, sc["x1"], sc["C1"], sc["x2"], sc["C2"])
        np.testing.assert_array_equal(xo.numpy(), xr)
        np.testing.assert_array_equal(Co.numpy(), Cr)       # operator- adds covariances too


def test_fusion_full_size_properties():
    """1M 6-dof fusions (BASELINE config 5): fusing an estimate with itself halves the covariance and
    keeps the mean; fusion is symmetric in its arguments.  Explicit inverses (DataModel.hpp:54) lose
    ~cond*eps, and among 1M random covariances a few are nearly singular, so the bound is per instance:
    1e-12 * cond."""
    n = 1 << 20
    sc = synth.fusion_scenario(n, d=6, log_spread=0.5)
    kap = np.maximum(np.linalg.cond(sc["C1"]), np.linalg.cond(sc["C2"]))
    x1, C1, x2, C2 = (engine.DeviceArray(sc[k]) for k in ("x1", "C1", "x2", "C2"))

    def close(a, b, scale):
        err = np.max(np.abs(a - b).reshape(n, -1), axis=1)
        assert np.all(err <= 1e-12 * kap * scale), float(np.max(err / (kap * scale)))

    xs, Cs = engine.DataModel.fuse(x1, C1, x1, C1)
    close(xs.numpy(), sc["x1"], 1.0 + np.max(np.abs(sc["x1"]), axis=1))
    close(Cs.numpy(), 0.5 * sc["C1"], np.max(np.abs(sc["C1"]).reshape(n, -1), axis=1))
    xa, Ca = engine.DataModel.fuse(x1, C1, x2, C2)
    xb, Cb = engine.DataModel.fuse(x2, C2, x1, C1)
    close(xa.numpy(), xb.numpy(), 1.0 + np.max(np.abs(xb.numpy()), axis=1))
    close(Ca.numpy(), Cb.numpy(), np.max(np.abs(Cb.numpy()).reshape(n, -1), axis=1))
    # in-place form, like data1.fusion(data2)
    engine.DataModel.fuse(x1, C1, x2, C2, out=(x1, C1))
    np.testing.assert_array_equal(x1.numpy(), xa.numpy())
    np.testing.assert_array_equal(C1.numpy(), Ca.numpy())


@pytest.mark.parametrize("d", [3, 6])
def test_fusion_golden_fixture(d):
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "fusion_d%d.npz" % d))
    xo, Co = engine.DataModel.fuse(g["x1"], g["C1"], g["x2"], g["C2"])
    np.testing.assert_array_equal(xo.numpy(), g["xo"])
    np.testing.assert_array_equal(Co.numpy(), g["Co"])


def test_fuse_host_zero_copy_with_pinned_buffers(slo):
    """Page-locked host arrays: slb_datamodel_fuse_host lets the kernel read and write them in place (mapped memory)."""
    import torch
    sc = synth.fusion_scenario(5000, d=6)
    pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
    h = [pin(sc[k]) for k in ("x1", "C1", "x2", "C2")]
    xo, Co = torch.empty((5000, 6), dtype=torch.float64).pin_memory(), torch.empty((5000, 6, 6), dtype=torch.float64).pin_memory()
    engine.check(engine.lib().slb_datamodel_fuse_host(6, 5000, *[engine._ptr_of(t) for t in h], engine._ptr_of(xo), engine._ptr_of(Co)))
    xr, Cr = slo.datamodel(0, sc["x1"], sc["C1"], sc["x2"], sc["C2"])
    np.testing.assert_array_equal(xo.numpy(), xr)
    np.testing.assert_array_equal(Co.numpy(), Cr)


def test_device_api_accepts_pinned_and_cuda_tensors(slo):
    """engine.dev() wraps torch tensors (CUDA or page-locked host) without a copy: the C ABI only sees pointers."""
    import torch
    sc = synth.fusion_scenario(777, d=3)
    xr, Cr = slo.datamodel(0, sc["x1"], sc["C1"], sc["x2"], sc["C2"])
    pinned = [torch.from_numpy(sc[k]).pin_memory() for k in ("x1", "C1", "x2", "C2")]
    xo, Co = engine.DataModel.fuse(*pinned)
    np.testing.assert_array_equal(xo.numpy(), xr)
    on_gpu = [t.cuda() for t in pinned]
    xo, Co = engine.DataModel.fuse(*on_gpu)
    np.testing.assert_array_equal(Co.numpy(), Cr)
    with pytest.raises(engine.SlbError):
        engine.DataModel.fuse(torch.from_numpy(sc["x1"]), *on_gpu[1:])      # pageable host tensor: refused, not copied silently
