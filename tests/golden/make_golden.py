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


def usckf_shapes():
    """One fused predict+update step for feature sizes other than the bench's (3, 9), and checkSigmaPoints of the result."""
    out = {}
    for nk, nl in ((6, 6), (9, 3), (3, 0)):
        sc = synth.usckf_scenario(4, seed=106 + nk, nk=nk, nl=nl)
        mu2, P2, st, _ = slo.usckf_step(slo.PM_USCKF_TEST, slo.MM_USCKF_VO, nk, nl, sc["mu"], sc["P"], sc["u"], sc["dt"], sc["Q"],
                                        sc["z"], sc["R"])
        assert not st.any()
        fl, diff = slo.check_sigma_points(2, mu2, P2, nk=nk, nl=nl)
        assert not fl.any()
        tag = "_%d_%d" % (nk, nl)
        for k, v in dict(mu0=sc["mu"], P0=sc["P"], u=sc["u"], z=sc["z"], Q=sc["Q"], R=sc["R"], mu2=mu2, P2=P2, diff=diff).items():
            out[k + tag] = v
        out["dt"] = sc["dt"]
    np.savez_compressed(os.path.join(OUT, "usckf_shapes.npz"), **out)


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


def next_rows():
    """error-state EKF, safeFusion, dead reckoning with uncertainty."""
    sc = synth.ekf_scenario(6, seed=104, outlier_frac=0.3)
    err1, P1 = slo.ekf_predict(sc["err"], sc["P"], sc["F"], sc["Q"])
    P2, ret, acc = slo.ekf_update(sc["mu"], P1, sc["z"], sc["H"], sc["R"], gate=1)
    mu3, P3, acc3 = slo.ekf_single_update(sc["mu"], err1, P2, sc["zs"], sc["Hs"], sc["R"], gate=1)
    np.savez_compressed(os.path.join(OUT, "ekf_n45.npz"), mu0=sc["mu"], err0=sc["err"], P0=sc["P"], F=sc["F"], Q=sc["Q"],
                        H=sc["H"], R=sc["R"], z=sc["z"], Hs=sc["Hs"], zs=sc["zs"], err1=err1, P1=P1, P2=P2, ret=ret, acc=acc,
                        mu3=mu3, P3=P3, acc3=acc3)
    sf = synth.safe_fusion_scenario(64, log_spread=1.5)
    xo, Co = slo.safe_fusion(sf["x1"], sf["C1"], sf["x2"], sf["C2"])
    np.savez_compressed(os.path.join(OUT, "safe_fusion_d3.npz"), xo=xo, Co=Co, **sf)
    dr = synth.deadreckon_scenario(32, seed=105)
    post, pcov, dpose, dcov = slo.dr_update_pose(dr["dt"], dr["vel0"], dr["vel1"], dr["velcov"], dr["prev_pose"], dr["prev_cov"])
    np.savez_compressed(os.path.join(OUT, "deadreckon.npz"), post=post, pcov=pcov, dpose=dpose, dcov=dcov, **dr)


if __name__ == "__main__":
    if len(sys.argv) > 1:                      # e.g. `make_golden.py usckf_shapes`: (re)generate the named fixtures only
        for name in sys.argv[1:]:
            globals()[name]()
        sys.exit(0)
    ukf()
    usckf_shapes()
    usckf()
    msckf()
    fusion()
    next_rows()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
