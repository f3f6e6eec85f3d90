set -x
cat > /tmp/san.py <<'PY'
import numpy as np, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch
from slam_localization_b200 import engine, synth
# UKF
sc = synth.ukfom_scenario(300, seed=1); f = engine.Ukf(300); f.set_state(sc["mu"], sc["P"])
f.step(engine.PM_UKFOM_IMU, engine.MM_GPS_POS, sc["u"], sc["dt"], sc["Q"], sc["z"], sc["R"])
pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
out = torch.empty((300, 10), dtype=torch.float64).pin_memory()
f.step_host(engine.PM_UKFOM_IMU, engine.MM_GPS_POS, pin(sc["u"]), sc["dt"], pin(sc["Q"]), pin(sc["z"]), pin(sc["R"]), mu_out=out)
# USCKF
sc = synth.usckf_scenario(70, seed=2); f = engine.Usckf(70); f.set_state(sc["mu"], sc["P"])
f.step(engine.PM_USCKF_TEST, engine.MM_USCKF_VO, sc["u"], sc["dt"], sc["Q"], sc["z"], sc["R"])
out = torch.empty((70, 51), dtype=torch.float64).pin_memory()
f.step_host(engine.PM_USCKF_TEST, engine.MM_USCKF_VO, pin(sc["u"]), sc["dt"], pin(sc["Q"]), pin(sc["z"]), pin(sc["R"]), mu_out=out)
f.cloning(engine.STATEK_I); f.cloning(engine.STATEK_L)
# MSCKF both flavours, k = 10 and k = 9, with outliers
for k in (10, 9):
    sc = synth.msckf_scenario(5, seed=3, k=k, nfeat=50, outlier_frac=0.08)
    for upd in ("update", "update_ekf"):
        f = engine.Msckf(5, nclones=k); f.set_state(sc["mu"], sc["P"])
        f.predict(engine.PM_MSCKF_DELTAPOSE, sc["u"], 0.0, sc["Q"])
        getattr(f, upd)(engine.MM_MSCKF_REPROJ, sc["landmarks"], sc["z"], sc["R"], gate=True)
        f.mu()
# fusion, next rows
fs = synth.fusion_scenario(1000, d=6); engine.DataModel.fuse(fs["x1"], fs["C1"], fs["x2"], fs["C2"])
fs = synth.fusion_scenario(77, d=3); engine.DataModel.fuse(fs["x1"], fs["C1"], fs["x2"], fs["C2"]); engine.DataModel.safe_fuse(fs["x1"], fs["C1"], fs["x2"], fs["C2"])
sc = synth.ekf_scenario(37, seed=4, outlier_frac=0.3); e = engine.ErrorStateEkf(sc["mu"], sc["err"], sc["P"])
e.ekf_predict(sc["F"], sc["Q"]); e.ekf_update(sc["z"], sc["H"], sc["R"], gate=True); e.ekf_single_update(sc["zs"], sc["Hs"], sc["R"], gate=True); e.cloning()
dr = synth.deadreckon_scenario(100); engine.DeadReckon.update_pose(dr["dt"], dr["vel0"], dr["vel1"], dr["velcov"], dr["prev_pose"], dr["prev_cov"])
torch.cuda.synchronize(); print("sanitizer driver ok")
PY
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/san.py 2>&1 | tail -25 | tee gpurun_out/r02_sanitizer_memcheck.log
