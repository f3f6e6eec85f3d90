"""Third, precision-independent pin of the oracle: one Usckf::update and one Msckf::update (with
applyDelta) evaluated in 50-digit arithmetic (tests/mp_ref.py, written from the reference's formulas) against the
double-precision C++ oracle on the same inputs.  Agreement far below the 1e-9 parity tolerance means the tolerance is
spent on the GPU kernels' summation order, not on the oracle's own rounding."""
import numpy as np

import mp_ref
from slam_localization_b200 import synth

AUG = [0, 1, 0, 0] * 3


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))


def test_usckf_update_matches_50_digit_arithmetic(slo):
    nk, nl = 3, 9
    sc = synth.usckf_scenario(2, seed=501, nk=nk, nl=nl)
    mu_o, P_o, st, _ = slo.usckf_step(slo.PM_USCKF_TEST, slo.MM_USCKF_VO, nk, nl, sc["mu"], sc["P"], None, 0.0, None, sc["z"],
                                      sc["R"], predict=False)
    assert not st.any()
    for i in range(2):
        mu_m, P_m = mp_ref.usckf_update(AUG, sc["mu"][i], sc["P"][i], sc["z"][i], lambda a: mp_ref.mm_usckf_vo(a, nk),
                                        sc["R"], nk + nl)
        mu_m, P_m = mp_ref.to_float(mu_m), mp_ref.to_float(P_m)
        d = slo.boxminus(AUG, mu_o[i], mu_m, nk + nl)
        assert np.max(np.abs(d)) <= 1e-12, np.max(np.abs(d))
        assert _rel(np.tril(P_o[i]), np.tril(P_m)) <= 1e-11


def test_msckf_update_with_apply_delta_matches_50_digit_arithmetic(slo):
    k, nfeat = 3, 6                                       # N = 30, 61 sigma points, m = 12: seconds in mpmath
    sc = synth.msckf_scenario(1, seed=502, k=k, nfeat=nfeat)
    blocks = [0, 1, 0, 0] + [0, 1] * k
    mu_o, P_o, out, st, it = slo.msckf_update(slo.MM_MSCKF_REPROJ, k, sc["mu"], sc["P"], sc["landmarks"], sc["z"], sc["R"],
                                              gate=False)
    assert not st.any() and not out.any()
    lm = sc["landmarks"]
    mu_m, P_m = mp_ref.msckf_update(blocks, sc["mu"][0], sc["P"][0], sc["z"][0], lambda s: mp_ref.mm_msckf_reproj(s, k, lm),
                                    sc["R"])
    mu_m, P_m = mp_ref.to_float(mu_m), mp_ref.to_float(P_m)
    d = slo.boxminus(blocks, mu_o[0], mu_m)
    assert np.max(np.abs(d)) <= 1e-12, np.max(np.abs(d))
    assert _rel(P_o[0], P_m) <= 1e-10
