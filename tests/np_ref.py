"""Independent numpy/scipy re-implementation of the sigma-point filters, used ONLY to cross-check
the C++ oracle (two independent restatements agreeing to ~1e-12 is the best pin available: the
reference cannot be built here and asserts no filter output -- SURVEY.md 4.4).

Deliberately different from oracle/: SO(3) exp/log through scipy's Rotation (rotation vectors),
Cholesky through LAPACK (numpy.linalg.cholesky), inverses through numpy.linalg.inv/solve, and the
augmented-state deltas applied WITHOUT the exp/log round trips of quirk Q10.
"""
import numpy as np
from scipy.spatial.transform import Rotation as Rot


def q_to_rot(q):  # (w,x,y,z) -> Rotation
    return Rot.from_quat([q[1], q[2], q[3], q[0]])


def rot_to_q(r):
    x, y, z, w = r.as_quat()
    return np.array([w, x, y, z])


def blocks_qoff(blocks):
    offs, o = [], 0
    for s in blocks:
        offs.append(o)
        o += 4 if s else 3
    return offs, o


def boxplus(blocks, x, d, nfeat=0):
    y = np.array(x, dtype=float)
    offs, qn = blocks_qoff(blocks)
    for b, s in enumerate(blocks):
        o = offs[b]
        if s:
            y[o:o + 4] = rot_to_q(q_to_rot(x[o:o + 4]) * Rot.from_rotvec(d[3 * b:3 * b + 3]))
        else:
            y[o:o + 3] = x[o:o + 3] + d[3 * b:3 * b + 3]
    if nfeat:
        y[qn:qn + nfeat] = x[qn:qn + nfeat] + d[3 * len(blocks):]
    return y


def boxminus(blocks, a, b_, nfeat=0):
    d = np.zeros(3 * len(blocks) + nfeat)
    offs, qn = blocks_qoff(blocks)
    for b, s in enumerate(blocks):
        o = offs[b]
        if s:
            d[3 * b:3 * b + 3] = (q_to_rot(b_[o:o + 4]).inv() * q_to_rot(a[o:o + 4])).as_rotvec()
        else:
            d[3 * b:3 * b + 3] = a[o:o + 3] - b_[o:o + 3]
    if nfeat:
        d[3 * len(blocks):] = a[qn:qn + nfeat] - b_[qn:qn + nfeat]
    return d


def sigma_points(blocks, mu, delta, P, nfeat=0):
    L = np.linalg.cholesky(P)
    n = P.shape[0]
    X = [boxplus(blocks, mu, delta, nfeat)]
    for j in range(n):
        X.append(boxplus(blocks, mu, delta + L[:, j], nfeat))
        X.append(boxplus(blocks, mu, delta - L[:, j], nfeat))
    return X


def mean_manifold(blocks, X, nfeat=0):
    ref = X[0].copy()
    it = 0
    while True:
        md = np.mean([boxminus(blocks, x, ref, nfeat) for x in X], axis=0)
        ref = boxplus(blocks, ref, md, nfeat)
        it += 1
        if not (np.linalg.norm(md) > 1e-6 and it < 10000):
            break
    return ref, it


def cov_manifold(blocks, mean, X, nfeat=0):
    D = np.array([boxminus(blocks, x, mean, nfeat) for x in X])
    return 0.5 * D.T @ D


def ukf_predict(blocks, mu, P, g, Q):
    X = [g(x) for x in sigma_points(blocks, mu, np.zeros(P.shape[0]), P)]
    m, _ = mean_manifold(blocks, X)
    return m, cov_manifold(blocks, m, X) + Q


def ukf_apply_delta(blocks, mu, P, delta, nfeat=0):
    X = sigma_points(blocks, mu, delta, P, nfeat)
    m, _ = mean_manifold(blocks, X, nfeat)
    return m, cov_manifold(blocks, m, X, nfeat)


def ukf_update(blocks, mu, P, z, h, R, nfeat=0, apply_delta=True):
    n = P.shape[0]
    X = sigma_points(blocks, mu, np.zeros(n), P, nfeat)
    Z = np.array([h(x) for x in X])
    zb = Z.mean(axis=0)
    DZ = Z - zb
    S = 0.5 * DZ.T @ DZ + R
    DX = np.array([boxminus(blocks, x, mu, nfeat) for x in X])
    Pxz = 0.5 * DX.T @ DZ
    K = np.linalg.solve(S.T, Pxz.T).T
    nu = z - zb
    P2 = P - K @ S @ K.T
    if apply_delta:
        return ukf_apply_delta(blocks, mu, P2, K @ nu, nfeat)
    return boxplus(blocks, mu, K @ nu, nfeat), P2


# ---- models (independent formulations through rotation matrices) --------------------------------
def pm_ukfom_imu(s, u, dt, refbug=False):
    o = np.zeros(10)
    R = q_to_rot(s[3:7])
    base = Rot.identity() if refbug else R
    o[3:7] = rot_to_q(base * Rot.from_rotvec(u[3:6] * dt))
    o[7:10] = s[7:10] + (R.apply(u[0:3]) + np.array([0, 0, 9.81])) * dt
    o[0:3] = s[0:3] + s[7:10] * dt
    return o


def pm_pose6_odom(s, u, dt):
    o = np.zeros(7)
    R = q_to_rot(s[3:7])
    o[0:3] = s[0:3] + R.apply(u[0:3]) * dt
    o[3:7] = rot_to_q(R * Rot.from_rotvec(u[3:6] * dt))
    return o


def pm_usckf_test(s, u, dt):
    o = np.zeros(13)
    o[3:7] = rot_to_q(q_to_rot(s[3:7]) * Rot.from_rotvec(u[3:6] * dt))
    o[10:13] = u[3:6]
    o[7:10] = u[0:3]
    o[0:3] = s[0:3] + s[7:10] * dt
    return o


def pm_msckf_deltapose(s, u):
    o = np.zeros(13)
    Rn = q_to_rot(s[3:7]) * q_to_rot(u[3:7])
    o[3:7] = rot_to_q(Rn)
    o[0:3] = s[0:3] + Rn.apply(u[0:3])
    o[7:10] = u[7:10]
    o[10:13] = u[10:13]
    return o


def mm_usckf_vo(a, nk):
    Rk, Ri = q_to_rot(a[3:7]), q_to_rot(a[29:33])
    Rd = Ri.inv() * Rk
    dp = a[0:3] - a[26:29]
    z = np.zeros(nk)
    for i in range(0, nk, 3):
        z[i:i + 3] = Rd.apply(a[39 + i:39 + i + 3]) + dp
    return z


def mm_msckf_reproj(s, k, lm):
    nfeat = lm.shape[0]
    z = np.zeros(2 * nfeat)
    for f in range(nfeat):
        j = f % k
        p, q = s[13 + 7 * j:16 + 7 * j], s[16 + 7 * j:20 + 7 * j]
        pc = q_to_rot(q).inv().apply(lm[f] - p)
        z[2 * f:2 * f + 2] = pc[0:2] / pc[2]
    return z


AUG_BLOCKS = [0, 1, 0, 0] * 3
STATE_BLOCKS = [0, 1, 0, 0]


def multi_blocks(k):
    return STATE_BLOCKS + [0, 1] * k


def usckf_predict(mu, P, nk, nl, f, Q):
    """Usckf::predict, written from the algebra (Fk = Pxy^T Pii^-1, blocks of row/col i scaled)."""
    mu, P = mu.copy(), P.copy()
    si = mu[26:39].copy()
    Pii = P[24:36, 24:36].copy()
    X0 = sigma_points(STATE_BLOCKS, si, np.zeros(12), Pii)
    X = [f(x) for x in X0]
    m, _ = mean_manifold(STATE_BLOCKS, X)
    DX0 = np.array([boxminus(STATE_BLOCKS, x, si) for x in X0])
    DX = np.array([boxminus(STATE_BLOCKS, x, m) for x in X])
    Pxy = 0.5 * DX0.T @ DX
    Fk = Pxy.T @ np.linalg.inv(Pii)
    mu[26:39] = m
    N = P.shape[0]
    others = [i for i in range(N) if not (24 <= i < 36)]
    P[np.ix_(range(24, 36), others)] = Fk @ P[np.ix_(range(24, 36), others)]
    P[np.ix_(others, range(24, 36))] = P[np.ix_(others, range(24, 36))] @ Fk.T
    P[24:36, 24:36] = 0.5 * DX.T @ DX + Q
    return mu, P


def usckf_update(mu, P, nk, nl, z, R):
    return ukf_update(AUG_BLOCKS, mu, P, z, lambda a: mm_usckf_vo(a, nk), R, nfeat=nk + nl, apply_delta=False)


def fusion(x1, C1, x2, C2):
    I1, I2 = np.linalg.inv(C1), np.linalg.inv(C2)
    P = np.linalg.inv(I1 + I2)
    return P @ (I1 @ x1 + I2 @ x2), P
