"""Pins the oracle's manifold algebra against the ONLY numeric assertions the reference holds:
test/MsckfUnitTest.cpp:61,62,66,71,110,113.  `==` on wrapped states means
"boxminus is below 1e-12 in every component" (MtkWrap.hpp:104-108,231-235)."""
import numpy as np

D2R = np.pi / 180.0
STATE = [0, 1, 0, 0]


def multi(k):
    return STATE + [0, 1] * k


def ident(blocks):
    x = []
    for s in blocks:
        x += [1.0, 0, 0, 0] if s else [0.0, 0, 0]
    return np.array(x)


def eq(slo, blocks, a, b):
    return np.all(np.abs(slo.boxminus(blocks, a, b)) <= 1e-12)


def test_dof_equals_vectorized_size(slo):          # MsckfUnitTest.cpp:61
    mstate = ident(multi(0))
    assert slo.get_vectorized(multi(0), mstate).size == 12


def test_state_equals_itself(slo):                 # :62
    mstate = ident(multi(0))
    assert eq(slo, multi(0), mstate, mstate)


def test_set_of_vectorized_roundtrip(slo):         # :64-66
    mstate = ident(multi(0))
    bis = slo.set_from_vector(multi(0), slo.get_vectorized(multi(0), mstate))
    assert eq(slo, multi(0), mstate, bis)


def test_reduced_multistate_dof(slo):              # :68-71 ReducedState = pos + orient
    assert 3 * len([0, 1]) == 6


def _bis(slo):
    bis = ident(multi(0))
    bis[0:3] = [1, 2.0, -3.0]
    bis[3:7] = slo.boxplus([1], bis[3:7], np.array([1.0, 1.0, 1.0]) * D2R)   # orient.boxplus(euler) :87
    return bis


def test_set_equals_boxplus(slo):                  # :98-110: resstate.set(v) == mstate + v
    b = multi(0)
    mstate, bis = ident(b), _bis(slo)
    v = slo.boxminus(b, mstate, bis)
    res = slo.set_from_vector(b, v)
    summ = slo.boxplus(b, mstate, v)
    assert eq(slo, b, res, summ)
    # and the convention it pins: exp(v) is a rotation of |v| rad about v/|v|
    w = np.array([0.3, -0.2, 0.5])
    q = slo.so3_exp(w)
    th = np.linalg.norm(w)
    np.testing.assert_allclose(q, np.r_[np.cos(th / 2), np.sin(th / 2) * w / th], rtol=0, atol=1e-15)


def test_boxplus_inverts_boxminus(slo):            # :111-113: mstate + (-(mstate - bis)) == bis
    b = multi(0)
    mstate, bis = ident(b), _bis(slo)
    v = slo.boxminus(b, mstate, bis)
    assert eq(slo, b, bis, slo.boxplus(b, mstate, -v))


def test_exp_log_roundtrip_and_small_angle_branch(slo):
    rng = np.random.default_rng(0)
    for scale in (1e-9, 1e-5, 1e-3, 0.02, 0.5, 2.0, 3.1):
        v = rng.normal(size=3)
        v *= scale / np.linalg.norm(v)
        np.testing.assert_allclose(slo.so3_log(slo.so3_exp(v)), v, rtol=1e-12, atol=1e-18)
    # log of the identity is exactly zero (|qv| clamp to 1e-11)
    assert np.all(slo.so3_log(np.array([1.0, 0, 0, 0])) == 0.0)
    # +-q identified (atan, not atan2)
    q = slo.so3_exp(np.array([0.1, 0.2, -0.3]))
    np.testing.assert_allclose(slo.so3_log(q), slo.so3_log(-q), rtol=0, atol=1e-15)


def test_multistate_with_sensor_poses(slo):
    b = multi(4)
    rng = np.random.default_rng(1)
    x = ident(b)
    d = rng.normal(size=36) * 0.3
    y = slo.boxplus(b, x, d)
    np.testing.assert_allclose(slo.boxminus(b, y, x), d, rtol=1e-12, atol=1e-14)
    assert eq(slo, b, slo.boxplus(b, x, slo.boxminus(b, y, x)), y)


def test_augmented_state_features(slo):
    b = STATE * 3
    rng = np.random.default_rng(2)
    x = np.r_[ident(b), rng.normal(size=12)]
    d = rng.normal(size=48) * 0.2
    y = slo.boxplus(b, x, d, nfeat=12)
    np.testing.assert_allclose(y[39:], x[39:] + d[36:], rtol=0, atol=1e-15)
    np.testing.assert_allclose(slo.boxminus(b, y, x, nfeat=12), d, rtol=1e-12, atol=1e-14)


def test_chi2_gate_table(slo):                     # Usckf.hpp:794-855
    th = [3.84, 5.99, 7.81, 9.49, 11.07, 12.59, 14.07, 15.51, 16.92]
    for dof, t in enumerate(th, start=1):
        assert slo.accept_mahalanobis(t - 1e-9, dof)
        assert not slo.accept_mahalanobis(t, dof)
    assert not slo.accept_mahalanobis(0.0, 10)
    assert not slo.accept_mahalanobis(0.0, 0)
