"""Multi-precision (mpmath, 50 digits) restatement of one Usckf::update (Usckf.hpp:260-308) and one Msckf::update
with applyDelta (Msckf.hpp:220-277,659-666), used ONLY as a third, precision-independent check of the C++ oracle:
the oracle (double, Eigen-style operation order) and tests/np_ref.py (double, LAPACK) share a floating-point format,
this file does not.  Agreement of the oracle with it at ~1e-12 says the oracle's rounding behaviour is benign at the
1e-9 parity tolerance the GPU tests use.  Written from the reference's formulas, not from oracle/ code."""
import mpmath as mp

mp.mp.dps = 50


def _v(x):
    return [mp.mpf(float(t)) for t in x]


def q_mul(a, b):
    return [a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
            a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
            a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3],
            a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1]]


def q_conj(a):
    return [a[0], -a[1], -a[2], -a[3]]


def so3_exp(v):
    th = mp.sqrt(v[0] ** 2 + v[1] ** 2 + v[2] ** 2)
    if th == 0:
        return [mp.mpf(1), mp.mpf(0), mp.mpf(0), mp.mpf(0)]
    s = mp.sin(th / 2) / th
    return [mp.cos(th / 2), s * v[0], s * v[1], s * v[2]]


def so3_log(q):
    n = mp.sqrt(q[1] ** 2 + q[2] ** 2 + q[3] ** 2)
    if n == 0:
        return [mp.mpf(0)] * 3
    s = 2 * mp.atan(n / q[0]) / n        # MTK: atan, not atan2 (q and -q are the same rotation)
    return [s * q[1], s * q[2], s * q[3]]


def rot_apply(q, v):
    """R(q) v through the rotation matrix (Eigen Quaternion::toRotationMatrix)."""
    w, x, y, z = q
    R = [[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
         [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
         [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]]
    return [sum(R[i][k] * v[k] for k in range(3)) for i in range(3)]


def qoffs(blocks):
    offs, o = [], 0
    for s in blocks:
        offs.append(o)
        o += 4 if s else 3
    return offs, o


def boxplus(blocks, x, d, nfeat=0):
    y = list(x)
    offs, qn = qoffs(blocks)
    for b, s in enumerate(blocks):
        o = offs[b]
        if s:
            y[o:o + 4] = q_mul(x[o:o + 4], so3_exp(d[3 * b:3 * b + 3]))
        else:
            for i in range(3):
                y[o + i] = x[o + i] + d[3 * b + i]
    for f in range(nfeat):
        y[qn + f] = x[qn + f] + d[3 * len(blocks) + f]
    return y


def boxminus(blocks, a, b_, nfeat=0):
    d = [mp.mpf(0)] * (3 * len(blocks) + nfeat)
    offs, qn = qoffs(blocks)
    for b, s in enumerate(blocks):
        o = offs[b]
        if s:
            d[3 * b:3 * b + 3] = so3_log(q_mul(q_conj(b_[o:o + 4]), a[o:o + 4]))
        else:
            for i in range(3):
                d[3 * b + i] = a[o + i] - b_[o + i]
    for f in range(nfeat):
        d[3 * len(blocks) + f] = a[qn + f] - b_[qn + f]
    return d


def _sigma(blocks, mu, delta, P, nfeat):
    n = P.rows
    L = mp.cholesky(P)
    X = [boxplus(blocks, mu, delta, nfeat)]
    for j in range(n):                                    # unscaled sigma points (quirk Q1)
        X.append(boxplus(blocks, mu, [delta[i] + L[i, j] for i in range(n)], nfeat))
        X.append(boxplus(blocks, mu, [delta[i] - L[i, j] for i in range(n)], nfeat))
    return X


def _gain(blocks, mu, P, z, h, R, nfeat):
    n, m = P.rows, len(z)
    X = _sigma(blocks, mu, [mp.mpf(0)] * n, P, nfeat)
    Z = [h(x) for x in X]
    ns = len(X)
    zb = [sum(Zi[c] for Zi in Z) / ns for c in range(m)]
    S = mp.matrix(m, m)
    Pxz = mp.matrix(n, m)
    for x, Zi in zip(X, Z):
        dz = [Zi[c] - zb[c] for c in range(m)]
        dx = boxminus(blocks, x, mu, nfeat)
        for r in range(m):
            for c in range(m):
                S[r, c] += dz[r] * dz[c] / 2
        for r in range(n):
            for c in range(m):
                Pxz[r, c] += dx[r] * dz[c] / 2
    S += R
    K = Pxz * (S ** -1)
    nu = mp.matrix([z[c] - zb[c] for c in range(m)])
    return K, S, K * nu


def usckf_update(blocks, mu, P, z, h, R, nfeat):
    """mu = mu [+] K nu, Pk -= K S K^T (no applyDelta: quirk Q3)."""
    mu, z = _v(mu), _v(z)
    P, R = mp.matrix(P.tolist()), mp.matrix(R.tolist())
    K, S, d = _gain(blocks, mu, P, z, h, R, nfeat)
    return boxplus(blocks, mu, [d[i] for i in range(P.rows)], nfeat), P - K * S * K.T


def msckf_update(blocks, mu, P, z, h, R):
    """Msckf UKF-flavoured update without the outlier gate: Pk -= K S K^T, then applyDelta(K nu) re-estimates the mean
    (iterative manifold mean, 1e-6 stop) and REPLACES Pk by the sigma-point covariance around it."""
    mu, z = _v(mu), _v(z)
    P, R = mp.matrix(P.tolist()), mp.matrix(R.tolist())
    n = P.rows
    K, S, d = _gain(blocks, mu, P, z, h, R, 0)
    P2 = P - K * S * K.T
    X = _sigma(blocks, mu, [d[i] for i in range(n)], P2, 0)
    ref = list(X[0])
    while True:
        md = [sum(boxminus(blocks, x, ref)[i] for x in X) / len(X) for i in range(n)]
        ref = boxplus(blocks, ref, md)
        if not mp.sqrt(sum(t * t for t in md)) > mp.mpf("1e-6"):
            break
    Pn = mp.matrix(n, n)
    for x in X:
        dx = boxminus(blocks, x, ref)
        for r in range(n):
            for c in range(n):
                Pn[r, c] += dx[r] * dx[c] / 2
    return ref, Pn


# ---- models ---------------------------------------------------------------------------------------------------
def mm_usckf_vo(a, nk):
    """test/UsckfUnitTest.cpp:62-86: delta = statek [-] statek_i as a transform, applied to the 3-D features."""
    dq = q_mul(q_conj(a[29:33]), a[3:7])
    dp = [a[i] - a[26 + i] for i in range(3)]
    z = []
    for i in range(0, nk, 3):
        r = rot_apply(dq, a[39 + i:39 + i + 3])
        z += [r[c] + dp[c] for c in range(3)]
    return z


def mm_msckf_reproj(s, k, lm):
    z = []
    for f in range(len(lm)):
        j = f % k
        p, q = s[13 + 7 * j:16 + 7 * j], s[16 + 7 * j:20 + 7 * j]
        pc = rot_apply(q_conj(q), [mp.mpf(float(lm[f][c])) - p[c] for c in range(3)])
        z += [pc[0] / pc[2], pc[1] / pc[2]]
    return z


def to_float(x):
    import numpy as np
    if isinstance(x, mp.matrix):
        return np.array([[float(x[r, c]) for c in range(x.cols)] for r in range(x.rows)])
    return np.array([float(t) for t in x])
