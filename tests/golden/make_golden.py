"""Generates the committed golden fixtures under tests/golden/ from the CPU oracle.

The reference cannot be built or imported in this environment (Eigen/MTK/Boost absent), so these
vectors are ORACLE-generated (the oracle itself is pinned by tests/test_oracle_*.py); they freeze
its outputs so that a later change to oracle/ or to the kernels cannot silently move both.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import slo  # noqa: E402
from slam_localization_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def ukf():
    sc = synth.ukfom_scenario(48, seed=101)
    mu1, P1, st, _ = slo.ukf_step(9, slo.PM_UKFOM_IMU, slo.MM_GPS_POS, sc["mu"], sc["P"], sc["u"], sc["dt"], sc["Q"],
                                  sc["z"], sc["R"])
    assert not st.any()
    np.savez_compressed(os.path.join(OUT, "ukf_mtk9.npz"), mu0=sc["mu"], P0=sc["P"], u=sc["u"], z=sc["z"], dt=sc["dt"],
                        Q=sc["Q"], R=sc["R"], mu1=mu1, P1=P1)


def usckf():
    sc = synth.usckf_scenario(12, seed=102)
    nk, nl = sc["nk"], sc["nl"]
    mu1, P1, st, _ = slo.usckf_step(slo.PM_USCKF_TEST, slo.MM_USCKF_VO, nk, nl, sc["mu"], sc["P"], sc["u"], sc["dt"],
                                    sc["Q"], None, None, update=False)
    mu2, P2, st2, _ = slo.usckf_step(slo.PM_USCKF_TEST, slo.MM_USCKF_VO, nk, nl, mu1, P1, None, 0.0, None, sc["z"],
                                     sc["R"], predict=False)
    assert not st.any() and not st2.any()
    np.savez_compressed(os.path.join(OUT, "usckf_n48.npz"), mu0=sc["mu"], P0=sc["P"], u=sc["u"], z=sc["z"], dt=sc["dt"],
                        Q=sc["Q"], R=sc["R"], mu1=mu1, P1=P1, mu2=mu2, P2=P2)


def msckf():
    sc = synth.msckf_scenario(4, seed=103, k=10, nfeat=50)
    mu1, P1, st = slo.msckf_predict(slo.PM_MSCKF_DELTAPOSE, 10, sc["mu"], sc["P"], sc["u"], 0.0, sc["Q"])
    mu2, P2, out, st2, _ = slo.msckf_update(slo.MM_MSCKF_REPROJ, 10, sc["mu"], sc["P"], sc["landmarks"], sc["z"], sc["R"])
    assert not st.any() and not st2.any()
    np.savez_compressed(os.path.join(OUT, "msckf_k10_f50.npz"), mu0=sc["mu"], P0=sc["P"], u=sc["u"], z=sc["z"],
                        Q=sc["Q"], R=sc["R"], landmarks=sc["landmarks"], mu_pred=mu1, P_pred=P1, mu_upd=mu2, P_upd=P2,
                        outliers=out)


def fusion():
    for d in (3, 6):
        sc = synth.fusion_scenario(64, d=d)
        xo, Co = slo.datamodel(0, sc["x1"], sc["C1"], sc["x2"], sc["C2"])
        np.savez_compressed(os.path.join(OUT, "fusion_d%d.npz" % d), xo=xo, Co=Co, **sc)


if __name__ == "__main__":
    ukf()
    usckf()
    msckf()
    fusion()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
