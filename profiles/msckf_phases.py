"""Phase timeline of msckf_update_kernel (clock64 stamps of CTA 0, one instance per CTA round, BASELINE config-3 shapes).
On the GPU box:
    make -C slam-localization_b200/csrc phases            (-> libslb_phases.so, an instrumented copy, not the product library)
    SLB_LIB=$PWD/slam-localization_b200/csrc/libslb_phases.so python profiles/msckf_phases.py"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import torch  # noqa: E402
from slam_localization_b200 import engine, synth  # noqa: E402

NAMES = ["load record", "chol(P) [skipped: parked by the previous instance]", "sigma points through h", "zbar / innovation", "W, centre Z",
         "covXZ = L W (TRMM)", "S = 0.5 Zc^T Zc + R", "gate", "compaction of S / covXZ", "chol(S') + Y + w", "delta = Y w (+ next record)",
         "P_new = P - Y Y^T", "chol_dual(P_new, next P)", "sigma points X", "manifold mean", "deviations", "P = 0.5 D^T D", "mean out"]
B = 148 * 4
sc = synth.msckf_scenario(148, seed=3)
f = engine.Msckf(B, nclones=10)
f.set_state(sc["mu"], sc["P"], replicate=True)
z = engine.DeviceArray(np.tile(sc["z"], (4, 1)))
R, lm = engine.DeviceArray(sc["R"]), engine.DeviceArray(sc["landmarks"])
f.update(engine.MM_MSCKF_REPROJ, lm, z, R)
torch.cuda.synchronize()
lib = ctypes.CDLL(os.environ["SLB_LIB"])
out = np.zeros(4 * 32, dtype=np.int64)
assert lib.slb_debug_msckf_phases(out.ctypes.data_as(ctypes.c_void_p)) == 0
d = out.reshape(4, 32)
for it in (1, 2):
    t = d[it]
    print("instance round %d of CTA 0: %d cycles" % (it, t[18] - t[0]))
    for k in range(18):
        if t[k + 1] and t[k]:
            print("  %-52s %7d  (%4.1f %%)" % (NAMES[k], t[k + 1] - t[k], 100.0 * (t[k + 1] - t[k]) / (t[18] - t[0])))
