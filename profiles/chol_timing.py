"""Per-panel timeline of one chol_blocked call (clock64 stamps of CTA 0; one instance per SM so nothing else competes).
On the GPU box:
    make -C slam-localization_b200/csrc timing CALL=<k>     (k-th chol_blocked call of the kernel: UKF flavour 0 = P, 1 = S'+Y,
                                                            2 = P_new; EKF flavour 0 = gate S + L^-1, 1 = S'+Y)
    cp slam-localization_b200/csrc/libslb_timing.so slam-localization_b200/csrc/libslb.so      (the box's copy is scratch)
    python profiles/chol_timing.py ukf|ekf slam-localization_b200/csrc/libslb.so
Columns: end of the panel solve (slowest warp), everybody past the barrier, warp 0 after its tile / after factoring the next
diagonal block, when the other warps finish their tasks (min / median / max), end of the panel.  Cycles from the panel start."""
import ctypes
import sys

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import torch  # noqa: E402
from slam_localization_b200 import engine, synth  # noqa: E402

which, libpath = sys.argv[1], sys.argv[2]
B = 148
sc = synth.msckf_scenario(B, seed=3)
f = engine.Msckf(B, nclones=10)
f.set_state(sc["mu"], sc["P"])
u, z = engine.DeviceArray(sc["u"]), engine.DeviceArray(sc["z"])
Q, R, lm = engine.DeviceArray(sc["Q"]), engine.DeviceArray(sc["R"]), engine.DeviceArray(sc["landmarks"])
f.predict(engine.PM_MSCKF_DELTAPOSE, u, 0.0, Q)
(f.update if which == "ukf" else f.update_ekf)(engine.MM_MSCKF_REPROJ, lm, z, R)
torch.cuda.synchronize()
lib = ctypes.CDLL(libpath)
out = np.zeros(16 * 16 * 8, dtype=np.int64)
assert lib.slb_debug_chol(out.ctypes.data_as(ctypes.c_void_p)) == 0
d = out.reshape(16, 16, 8)
tot = 0
for p in range(16):
    if d[p, 0, 0] == 0:
        break
    t0 = d[p, :, 0].min()
    others = d[p, 1:, 4] - t0
    end = d[p, :, 5].max() - t0
    tot += end
    print(f"panel {p:2d}: solve {d[p, :, 1].max() - t0:5d}  past barrier {d[p, :, 2].max() - t0:5d}  warp0 tile {max(d[p, 0, 3] - t0, 0):5d} "
          f"factor {d[p, 0, 4] - t0:5d}  tiles {others.min():5d}/{int(np.median(others)):5d}/{others.max():5d}  end {end:5d}")
print("total", tot)
