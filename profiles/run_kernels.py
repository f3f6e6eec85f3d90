"""Small driver for ncu: launches each hot kernel a few times on BASELINE-shaped batches.
Usage (under gpurun):  python profiles/run_kernels.py [ukf] [usckf] [fusion] [msckf]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from slam_localization_b200 import engine, synth  # noqa: E402

which = sys.argv[1:] or ["ukf", "usckf", "fusion"]
REP = 3
if "ukf" in which:
    B = 65536
    sc = synth.ukfom_scenario(B, seed=1, p_scale=1e-4)
    f = engine.Ukf(B)
    f.set_state(sc["mu"], sc["P"])
    u, z, Q, R = (engine.DeviceArray(sc[k]) for k in ("u", "z", "Q", "R"))
    for _ in range(REP):
        f.step(engine.PM_UKFOM_IMU, engine.MM_GPS_POS, u, sc["dt"], Q, z, R)
    torch.cuda.synchronize()
    assert sum(f.status_counts()) == 0
if "usckf" in which:
    B = 65536
    sc = synth.usckf_scenario(2048, seed=2)
    f = engine.Usckf(B)
    f.set_state(sc["mu"], sc["P"], replicate=True)
    rep = B // 2048
    u, z = engine.DeviceArray(np.tile(sc["u"], (rep, 1))), engine.DeviceArray(np.tile(sc["z"], (rep, 1)))
    Q, R = engine.DeviceArray(sc["Q"]), engine.DeviceArray(sc["R"])
    for _ in range(REP):
        f.step(engine.PM_USCKF_TEST, engine.MM_USCKF_VO, u, sc["dt"], Q, z, R)
    torch.cuda.synchronize()
    assert sum(f.status_counts()) == 0
if "fusion" in which:
    n = 1 << 20
    sc = synth.fusion_scenario(n, d=6)
    a = [engine.DeviceArray(sc[k]) for k in ("x1", "C1", "x2", "C2")]
    out = (engine.DeviceArray(shape=(n, 6)), engine.DeviceArray(shape=(n, 6, 6)))
    for _ in range(REP):
        engine.DataModel.fuse(*a, out=out)
    torch.cuda.synchronize()
if "msckf" in which:
    B = 4096
    sc = synth.msckf_scenario(256, seed=3)
    f = engine.Msckf(B, nclones=10)
    f.set_state(sc["mu"], sc["P"], replicate=True)
    rep = B // 256
    u, z = engine.DeviceArray(np.tile(sc["u"], (rep, 1))), engine.DeviceArray(np.tile(sc["z"], (rep, 1)))
    Q, R, lm = engine.DeviceArray(sc["Q"]), engine.DeviceArray(sc["R"]), engine.DeviceArray(sc["landmarks"])
    for _ in range(REP):
        f.predict(engine.PM_MSCKF_DELTAPOSE, u, 0.0, Q)
        f.update(engine.MM_MSCKF_REPROJ, lm, z, R)
    torch.cuda.synchronize()
if "msckf_ekf" in which:
    B = 2048
    sc = synth.msckf_scenario(256, seed=3)
    f = engine.Msckf(B, nclones=10)
    f.set_state(sc["mu"], sc["P"], replicate=True)
    rep = B // 256
    z = engine.DeviceArray(np.tile(sc["z"], (rep, 1)))
    R, lm = engine.DeviceArray(sc["R"]), engine.DeviceArray(sc["landmarks"])
    for _ in range(2):
        f.update_ekf(engine.MM_MSCKF_REPROJ, lm, z, R)
    torch.cuda.synchronize()
if "ekf" in which:
    n = 65536
    sc = synth.ekf_scenario(1024, seed=1)
    rep = n // 1024
    f = engine.ErrorStateEkf(np.tile(sc["mu"], (rep, 1)), np.tile(sc["err"], (rep, 1)), np.tile(sc["P"], (rep, 1, 1)))
    F, Q, z, H, R = (engine.DeviceArray(x) for x in (np.tile(sc["F"], (rep, 1, 1)), sc["Q"], np.tile(sc["z"], (rep, 1)), sc["H"], sc["R"]))
    for _ in range(2):
        f.ekf_predict(F, Q)
        f.ekf_update(z, H, R, gate=False)
    torch.cuda.synchronize()
if "next" in which:
    n = 1 << 20
    sf = synth.safe_fusion_scenario(n, log_spread=1.5)
    a = [engine.DeviceArray(sf[k]) for k in ("x1", "C1", "x2", "C2")]
    for _ in range(2):
        engine.DataModel.safe_fuse(*a)
    dr = synth.deadreckon_scenario(n)
    for _ in range(2):
        engine.DeadReckon.update_pose(dr["dt"], dr["vel0"], dr["vel1"], dr["velcov"], dr["prev_pose"], dr["prev_cov"])
    torch.cuda.synchronize()
print("ok")
