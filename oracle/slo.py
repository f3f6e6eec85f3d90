"""ctypes binding of the CPU oracle (oracle/_build/libslo.so) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The oracle is a dependency-free restatement of the reference's Eigen/MTK
algorithm (see oracle/slo_core.hpp for what is and is not pinned by the reference's tests).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libslo.so")

# ids shared with include/slb.h
LAYOUT_POSE6, LAYOUT_MTK9, LAYOUT_STATE12 = 6, 9, 12
PM_UKFOM_IMU, PM_UKFOM_IMU_REFBUG, PM_POSE6_ODOM, PM_USCKF_TEST, PM_MSCKF_DELTAPOSE = 1, 2, 3, 4, 5
MM_GPS_POS, MM_USCKF_VO, MM_MSCKF_REPROJ = 101, 102, 103
STATEK, STATEK_L, STATEK_I = 1, 2, 3
ST_CHOL_FAIL, ST_MEAN_NOCONV, ST_GATE_REJECT, ST_NONFINITE = 1, 2, 4, 8

LAYOUT_BLOCKS = {LAYOUT_POSE6: [0, 1], LAYOUT_MTK9: [0, 1, 0], LAYOUT_STATE12: [0, 1, 0, 0]}


def build(force=False):
    """Compile the oracle with the committed Makefile (g++ only, no dependencies)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []),
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def hardware_threads():
    return int(lib().slo_hardware_threads())


# ---- primitives --------------------------------------------------------------------------------
def so3_exp(v, scale=1.0):
    v = _d(v)
    q = np.empty(4)
    lib().slo_so3_exp(_p(v), C.c_double(scale), _p(q))
    return q


def so3_log(q):
    q = _d(q)
    v = np.empty(3)
    lib().slo_so3_log(_p(q), _p(v))
    return v


def _lay(blocks):
    b = _i(blocks)
    return b, len(blocks)


def qdim(blocks, nfeat=0):
    return sum(4 if s else 3 for s in blocks) + nfeat


def boxplus(blocks, x, d, nfeat=0):
    b, nb = _lay(blocks)
    x, d = _d(x), _d(d)
    y = np.empty_like(x)
    lib().slo_boxplus(_p(b), nb, nfeat, _p(x), _p(d), _p(y))
    return y


def boxminus(blocks, a, bb, nfeat=0):
    b, nb = _lay(blocks)
    a, bb = _d(a), _d(bb)
    d = np.empty(3 * nb + nfeat)
    lib().slo_boxminus(_p(b), nb, nfeat, _p(a), _p(bb), _p(d))
    return d


def set_from_vector(blocks, v, nfeat=0):
    b, nb = _lay(blocks)
    v = _d(v)
    x = np.empty(qdim(blocks, nfeat))
    lib().slo_set_from_vector(_p(b), nb, nfeat, _p(v), _p(x))
    return x


def get_vectorized(blocks, x, nfeat=0):
    b, nb = _lay(blocks)
    x = _d(x)
    v = np.empty(3 * nb + nfeat)
    lib().slo_get_vectorized(_p(b), nb, nfeat, _p(x), _p(v))
    return v


def llt(P):
    P = _d(P)
    n = P.shape[0]
    L = np.empty_like(P)
    info = lib().slo_llt(n, _p(P), _p(L))
    return L, info


def inverse(A, fixed=False):
    A = _d(A)
    Ai = np.empty_like(A)
    lib().slo_inverse(A.shape[0], _p(A), _p(Ai), int(fixed))
    return Ai


def accept_mahalanobis(m2, dof):
    lib().slo_accept_mahalanobis.argtypes = [C.c_double, C.c_int]
    return bool(lib().slo_accept_mahalanobis(float(m2), int(dof)))


# ---- filters (batched, instance-major; return new arrays) ----------------------------------------
def ukf_step(layout, pm, mm, mu, P, u, dt, Q, z, R, gate_dof=0, predict=True, update=True, nthreads=1):
    mu, P = _d(mu).copy(), _d(P).copy()
    B = mu.shape[0]
    u = _d(u) if u is not None else np.zeros((B, 6))
    z = _d(z) if z is not None else np.zeros((B, 3))
    n = P.shape[1]
    Q = _d(Q) if Q is not None else np.zeros((n, n))
    R = _d(R) if R is not None else np.zeros((3, 3))
    st = np.zeros(B, np.int32)
    it = np.zeros(B, np.int32)
    rc = lib().slo_ukf_step(layout, pm, mm, B, _p(mu), _p(P), _p(u), C.c_double(dt), _p(Q), _p(z), _p(R),
                            gate_dof, int(predict), int(update), _p(st), _p(it), nthreads)
    assert rc == 0
    return mu, P, st, it


def usckf_step(pm, mm, nk, nl, mu, P, u, dt, Q, z, R, gate_dof=0, predict=True, update=True, nthreads=1):
    mu, P = _d(mu).copy(), _d(P).copy()
    B = mu.shape[0]
    u = _d(u) if u is not None else np.zeros((B, 6))
    z = _d(z) if z is not None else np.zeros((B, max(nk, 1)))
    Q = _d(Q) if Q is not None else np.zeros((12, 12))
    R = _d(R) if R is not None else np.zeros((max(nk, 1), max(nk, 1)))
    st = np.zeros(B, np.int32)
    it = np.zeros(B, np.int32)
    rc = lib().slo_usckf_step(pm, mm, B, nk, nl, _p(mu), _p(P), _p(u), C.c_double(dt), _p(Q), _p(z), _p(R),
                              gate_dof, int(predict), int(update), _p(st), _p(it), nthreads)
    assert rc == 0
    return mu, P, st, it


def usckf_clone(mode, nk, nl, mu, P):
    mu, P = _d(mu).copy(), _d(P).copy()
    lib().slo_usckf_clone(mode, mu.shape[0], nk, nl, _p(mu), _p(P), 1)
    return mu, P


def usckf_ctor_single(mu_single, P_single):
    mu_single, P_single = _d(mu_single), _d(P_single)
    B = mu_single.shape[0]
    mu = np.empty((B, 39))
    P = np.empty((B, 36, 36))
    lib().slo_usckf_ctor_single(B, _p(mu_single), _p(P_single), _p(mu), _p(P))
    return mu, P


def usckf_set_measurement(mode, nk, nl, mu, P, z, R):
    mu, P, z, R = _d(mu), _d(P), _d(z), _d(R)
    B, ln = z.shape
    nk2 = ln if mode == STATEK else nk
    nl2 = ln if mode == STATEK_L else nl
    mu2 = np.empty((B, 39 + nk2 + nl2))
    P2 = np.empty((B, 36 + nk2 + nl2, 36 + nk2 + nl2))
    lib().slo_usckf_set_measurement(mode, B, nk, nl, _p(mu), _p(P), ln, _p(z), _p(R), _p(mu2), _p(P2))
    return mu2, P2


def msckf_predict(pm, k, mu, P, u, dt, Q, nthreads=1):
    mu, P, u, Q = _d(mu).copy(), _d(P).copy(), _d(u), _d(Q)
    B = mu.shape[0]
    st = np.zeros(B, np.int32)
    rc = lib().slo_msckf_predict(pm, B, k, _p(mu), _p(P), _p(u), C.c_double(dt), _p(Q), _p(st), nthreads)
    assert rc == 0
    return mu, P, st


def msckf_update(mm, k, mu, P, landmarks, z, R, gate=True, nthreads=1):
    mu, P, landmarks, z, R = _d(mu).copy(), _d(P).copy(), _d(landmarks), _d(z), _d(R)
    B, m = z.shape
    out = np.zeros(B, np.int32)
    st = np.zeros(B, np.int32)
    it = np.zeros(B, np.int32)
    rc = lib().slo_msckf_update(mm, B, k, _p(mu), _p(P), _p(landmarks), m, _p(z), _p(R), int(gate), _p(out),
                                _p(st), _p(it), nthreads)
    assert rc == 0
    return mu, P, out, st, it


def check_sigma_points(kind, mu, P, nk=0, nl=0, k=0, nthreads=1):
    """checkSigmaPoints() of Usckf (kind 2) / Msckf (kind 3): returns (flags, diff[B,2])."""
    mu, P = _d(mu), _d(P)
    B = mu.shape[0]
    flags = np.zeros(B, np.int32)
    diff = np.zeros((B, 2))
    rc = lib().slo_check_sigma_points(kind, B, nk, nl, k, _p(mu), _p(P), _p(flags), _p(diff), nthreads)
    assert rc == 0
    return flags, diff


def msckf_remove_outliers(innov, S, N=12):
    innov, S = _d(innov), _d(S)
    m = innov.shape[0]
    kept = np.zeros(m, np.int32)
    nk = C.c_int(0)
    o = lib().slo_msckf_remove_outliers(m, N, _p(innov), _p(S), _p(kept), C.byref(nk))
    return int(o), kept[: nk.value].copy()


def datamodel(op, x1, C1, x2, C2, nthreads=1):
    x1, C1, x2, C2 = _d(x1), _d(C1), _d(x2), _d(C2)
    n, d = x1.shape
    xo, Co = np.empty_like(x1), np.empty_like(C1)
    lib().slo_datamodel(int(op), d, C.c_longlong(n), _p(x1), _p(C1), _p(x2), _p(C2), _p(xo), _p(Co), nthreads)
    return xo, Co


def datamodel_default(d):
    x, Cv = np.empty(d), np.empty((d, d))
    lib().slo_datamodel_default(d, _p(x), _p(Cv))
    return x, Cv


# ---- SURVEY 8(f) next rows (oracle/slo_next.hpp) ----------------------------------------------
def ekf_predict(err, P, F, Q, nthreads=1):
    err, P, F, Q = _d(err).copy(), _d(P).copy(), _d(F), _d(Q)
    lib().slo_ekf_predict(C.c_longlong(err.shape[0]), _p(err), _p(P), _p(F), _p(Q), nthreads)
    return err, P


def ekf_update(mu, P, z, H, R, gate=0, nthreads=1):
    mu, P, z, H, R = _d(mu), _d(P).copy(), _d(z), _d(H), _d(R)
    n, m = z.shape
    ret, acc = np.empty((n, m)), np.empty(n, dtype=np.int32)
    lib().slo_ekf_update(C.c_longlong(n), m, _p(mu), _p(P), _p(z), _p(H), _p(R), int(gate), _p(ret), _p(acc), nthreads)
    return P, ret, acc


def ekf_single_update(mu, err, P, z, H, R, gate=0, nthreads=1):
    mu, err, P, z, H, R = _d(mu).copy(), _d(err), _d(P).copy(), _d(z), _d(H), _d(R)
    n, m = z.shape
    acc = np.empty(n, dtype=np.int32)
    lib().slo_ekf_single_update(C.c_longlong(n), m, _p(mu), _p(err), _p(P), _p(z), _p(H), _p(R), int(gate), _p(acc), nthreads)
    return mu, P, acc


def ekf_clone(mu, err, P):
    mu, err, P = _d(mu).copy(), _d(err).copy(), _d(P).copy()
    lib().slo_ekf_clone(C.c_longlong(mu.shape[0]), _p(mu), _p(err), _p(P))
    return mu, err, P


def safe_fusion(x1, C1, x2, C2, nthreads=1):
    x1, C1, x2, C2 = _d(x1), _d(C1), _d(x2), _d(C2)
    xo, Co = np.empty_like(x1), np.empty_like(C1)
    lib().slo_safe_fusion(C.c_longlong(x1.shape[0]), _p(x1), _p(C1), _p(x2), _p(C2), _p(xo), _p(Co), nthreads)
    return xo, Co


def jacobi_svd(A):
    A = _d(A)
    n = A.shape[0]
    U, sv = np.empty((n, n)), np.empty(n)
    lib().slo_jacobi_svd(n, _p(A), _p(U), _p(sv))
    return U, sv


def transform_compose(pose2, cov2, pose1, cov1, nthreads=1):
    pose2, cov2, pose1, cov1 = _d(pose2), _d(cov2), _d(pose1), _d(cov1)
    po, co = np.empty_like(pose2), np.empty_like(cov2)
    lib().slo_transform_compose(C.c_longlong(pose2.shape[0]), _p(pose2), _p(cov2), _p(pose1), _p(cov1), _p(po), _p(co), nthreads)
    return po, co


def dr_update_pose(dt, vel0, vel1, velcov, prev_pose, prev_cov, nthreads=1):
    vel0, vel1, velcov, prev_pose, prev_cov = _d(vel0), _d(vel1), _d(velcov), _d(prev_pose), _d(prev_cov)
    n = vel0.shape[0]
    post, pcov, dpose, dcov = np.empty((n, 7)), np.empty((n, 6, 6)), np.empty((n, 7)), np.empty((n, 6, 6))
    lib().slo_dr_update_pose.argtypes = [C.c_longlong, C.c_double] + [C.c_void_p] * 9 + [C.c_int]
    lib().slo_dr_update_pose(n, float(dt), _p(vel0), _p(vel1), _p(velcov), _p(prev_pose), _p(prev_cov), _p(post), _p(pcov),
                             _p(dpose), _p(dcov), nthreads)
    return post, pcov, dpose, dcov


def dr_update_attitude(dt, w0, w1):
    w0, w1 = _d(w0), _d(w1)
    dq = np.empty(4)
    lib().slo_dr_update_attitude.argtypes = [C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    lib().slo_dr_update_attitude(float(dt), _p(w0), _p(w1), _p(dq))
    return dq


def msckf_update_ekf(mm, k, mu, P, landmarks, z, R, gate=True, nthreads=1):
    """Msckf::update, EKF flavour (Msckf.hpp:297-349).  Returns mu, P, outliers, status."""
    mu, P, landmarks, z, R = _d(mu).copy(), _d(P).copy(), _d(landmarks), _d(z), _d(R)
    B, m = z.shape
    out, st = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32)
    rc = lib().slo_msckf_update_ekf(int(mm), B, int(k), _p(mu), _p(P), _p(landmarks), m, _p(z), _p(R), 1 if gate else 0,
                                    _p(out), _p(st), nthreads)
    assert rc == 0
    return mu, P, out, st


def msckf_reproj_jac(k, mu, landmarks):
    mu, landmarks = _d(mu), _d(landmarks)
    nfeat = landmarks.shape[0]
    z, H = np.empty(2 * nfeat), np.empty((2 * nfeat, 12 + 6 * k))
    lib().slo_msckf_reproj_jac(int(k), _p(mu), _p(landmarks), nfeat, _p(z), _p(H))
    return z, H


def householder_qr(A):
    A = _d(A)
    rows, cols = A.shape
    QR, tau, Q = np.empty((rows, cols)), np.empty(min(rows, cols)), np.empty((rows, cols))
    lib().slo_householder_qr(rows, cols, _p(A), _p(QR), _p(tau), _p(Q))
    return QR, tau, Q
