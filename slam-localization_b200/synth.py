"""Seeded synthetic workloads for the five BASELINE.json configs.

Everything is plain numpy, instance-major, in the q-vector convention of include/slb.h, so the
same arrays feed the CUDA engine (through the C ABI) and the CPU oracle (through ctypes).
"""
import numpy as np

D2R = np.pi / 180.0
STATE_BLOCKS = [0, 1, 0, 0]
LAYOUT_BLOCKS = {6: [0, 1], 9: [0, 1, 0], 12: [0, 1, 0, 0]}


def random_unit_quat(rng, B, max_angle=np.pi):
    """(B,4) unit quaternions (w,x,y,z) with w >= 0, rotation angle uniform in [0,max_angle)."""
    ax = rng.normal(size=(B, 3))
    ax /= np.linalg.norm(ax, axis=1, keepdims=True)
    th = rng.uniform(0, max_angle, size=(B, 1))
    return np.concatenate([np.cos(th / 2), np.sin(th / 2) * ax], axis=1)


def random_spd(rng, B, n, scale=1.0, cond=1e3):
    """(B,n,n) SPD matrices scale * Q diag(10^U(-log10(cond),0)) Q^T, exactly symmetric."""
    A = rng.normal(size=(B, n, n))
    Qm, _ = np.linalg.qr(A)
    ev = 10.0 ** rng.uniform(-np.log10(cond), 0.0, size=(B, n))
    P = np.einsum("bij,bj,bkj->bik", Qm, ev, Qm) * scale
    return 0.5 * (P + P.transpose(0, 2, 1))


def identity_q(blocks, nfeat=0):
    x = []
    for s in blocks:
        x += [1.0, 0.0, 0.0, 0.0] if s else [0.0, 0.0, 0.0]
    return np.array(x + [0.0] * nfeat)


def random_q(rng, B, blocks, nfeat=0, vec_scale=1.0, max_angle=np.pi):
    cols = []
    for s in blocks:
        cols.append(random_unit_quat(rng, B, max_angle) if s else rng.normal(size=(B, 3)) * vec_scale)
    if nfeat:
        cols.append(rng.normal(size=(B, nfeat)) * vec_scale)
    return np.concatenate(cols, axis=1)


# ---- config 2: batched UKFoM (test/UKFoMUnitTest.cpp) ---------------------------------------------
def ukfom_process_noise(dt, layout=9):
    """process_noise_cov(dt), test/UKFoMUnitTest.cpp:73-80 (pos 0, orient 1e-4 dt, vel 2e-4 dt)."""
    n = layout
    Q = np.zeros((n, n))
    Q[3:6, 3:6] = 1e-4 * dt * np.eye(3)
    if n >= 9:
        Q[6:9, 6:9] = 2e-4 * dt * np.eye(3)
    if n == 6:
        Q[0:3, 0:3] = 1e-4 * dt * np.eye(3)
    return Q


def ukfom_fixture():
    """The reference's own UKFOM test inputs (test/UKFoMUnitTest.cpp:95-114), batch of one."""
    mu = identity_q(LAYOUT_BLOCKS[9])[None, :]
    P = 0.001 * np.eye(9)[None, :, :]
    u = np.array([[0.0, 0.0, 0.0, 10.0 * D2R, 0.0, 0.0]])
    z = np.array([[1.0, 0.0, 0.0]])
    return dict(mu=mu, P=P, u=u, z=z, dt=0.01, Q=ukfom_process_noise(0.01), R=1e-8 * np.eye(3))


def ukfom_scenario(B, seed=0, layout=9, dt=0.01, p_scale=1e-3, cond=1e3, r_sigma=1e-2):
    rng = np.random.default_rng(seed)
    blocks = LAYOUT_BLOCKS[layout]
    mu = random_q(rng, B, blocks, vec_scale=1.0)
    P = random_spd(rng, B, layout, scale=p_scale, cond=cond)
    u = np.concatenate([rng.normal(size=(B, 3)), rng.normal(size=(B, 3)) * 0.2], axis=1)
    z = mu[:, 0:3] + rng.normal(size=(B, 3)) * r_sigma
    return dict(mu=mu, P=P, u=u, z=z, dt=dt, Q=ukfom_process_noise(dt, layout), R=(r_sigma ** 2) * np.eye(3))


def ukfom_inputs(B, step, seed=0, r_sigma=1e-2, truth_pos=None):
    """Per-step control inputs / measurements (the only host->device traffic of a step)."""
    rng = np.random.default_rng([seed, step])
    u = np.concatenate([rng.normal(size=(B, 3)), rng.normal(size=(B, 3)) * 0.2], axis=1)
    base = truth_pos if truth_pos is not None else np.zeros((B, 3))
    z = base + rng.normal(size=(B, 3)) * r_sigma
    return u, z


def _qmul(a, b):
    w = a[:, 0] * b[:, 0] - a[:, 1] * b[:, 1] - a[:, 2] * b[:, 2] - a[:, 3] * b[:, 3]
    x = a[:, 0] * b[:, 1] + a[:, 1] * b[:, 0] + a[:, 2] * b[:, 3] - a[:, 3] * b[:, 2]
    y = a[:, 0] * b[:, 2] + a[:, 2] * b[:, 0] + a[:, 3] * b[:, 1] - a[:, 1] * b[:, 3]
    z = a[:, 0] * b[:, 3] + a[:, 3] * b[:, 0] + a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1]
    return np.stack([w, x, y, z], axis=1)


def _qexp(v):
    th = np.linalg.norm(v, axis=1, keepdims=True)
    s = np.where(th > 1e-12, np.sin(th / 2) / np.maximum(th, 1e-300), 0.5)
    return np.concatenate([np.cos(th / 2), s * v], axis=1)


def _qrot(q, v, inverse=False):
    w, u = q[:, 0:1], (-q[:, 1:4] if inverse else q[:, 1:4])
    t = 2.0 * np.cross(u, v)
    return v + w * t + np.cross(u, t)


class UkfomTruthRun:
    """A well-posed long replay for the UKFoM layout (pos, SO3, vel): B simulated vehicles fly a bounded Lissajous
    trajectory; each step yields the IMU sample u = (body acceleration, body rate) that drives the truth through the
    very process model the filter uses (test/UKFoMUnitTest.cpp:45-70: vel += (R(q) a + (0,0,9.81)) dt, so a carries
    the -9.81 that cancels it) plus sensor noise, and a GPS fix z = true position + noise.  The state is observable and
    bounded, so an estimator's rounding differences contract instead of growing -- unlike a replay whose measurements
    are drawn around the estimator's own mean, whose position runs away with the uncompensated 9.81 term and which is
    exponentially sensitive to 1e-15 perturbations (measured: two runs of the same CPU oracle differing by 1e-15 at
    step 0 are 5e-4 apart after 10k such steps)."""

    def __init__(self, B, seed=0, dt=0.01, r_sigma=1e-2, acc_sigma=2e-2, gyro_sigma=2e-3):
        self.B, self.dt, self.k = B, dt, 0
        self.rng = np.random.default_rng([seed, 777])
        self.r_sigma, self.acc_sigma, self.gyro_sigma = r_sigma, acc_sigma, gyro_sigma
        self.p = self.rng.normal(size=(B, 3))
        self.q = random_unit_quat(self.rng, B, max_angle=0.5)
        self.v = 0.3 * self.rng.normal(size=(B, 3))
        self.amp = 1.0 + self.rng.uniform(size=(B, 3))           # world-frame acceleration amplitudes, m/s^2
        self.om = 0.5 + 1.5 * self.rng.uniform(size=(B, 3))      # rad/s
        self.ph = 2 * np.pi * self.rng.uniform(size=(B, 3))
        self.wamp = 0.3 * self.rng.normal(size=(B, 3))           # body-rate amplitudes, rad/s

    def initial(self, p_scale=1e-4):
        """Prior mean = truth + a draw from the prior covariance p_scale * I."""
        sd = np.sqrt(p_scale)
        mu = np.concatenate([self.p + sd * self.rng.normal(size=(self.B, 3)),
                             _qmul(self.q, _qexp(sd * self.rng.normal(size=(self.B, 3)))),
                             self.v + sd * self.rng.normal(size=(self.B, 3))], axis=1)
        return mu, np.tile(p_scale * np.eye(9), (self.B, 1, 1))

    def step(self):
        t = self.k * self.dt
        a_w = self.amp * np.sin(self.om * t + self.ph) - 0.05 * self.p - 0.2 * self.v   # bounded: weak spring + damper
        a_b = _qrot(self.q, a_w - np.array([0.0, 0.0, 9.81]), inverse=True)
        w_b = self.wamp * np.cos(0.7 * self.om * t + self.ph)
        # truth through the filter's own process model (old orientation rotates a, old velocity moves pos)
        p = self.p + self.v * self.dt
        v = self.v + (_qrot(self.q, a_b) + np.array([0.0, 0.0, 9.81])) * self.dt
        q = _qmul(self.q, _qexp(w_b * self.dt))
        self.p, self.v, self.q = p, v, q / np.linalg.norm(q, axis=1, keepdims=True)
        self.k += 1
        u = np.concatenate([a_b + self.acc_sigma * self.rng.normal(size=(self.B, 3)),
                            w_b + self.gyro_sigma * self.rng.normal(size=(self.B, 3))], axis=1)
        z = self.p + self.r_sigma * self.rng.normal(size=(self.B, 3))
        return u, z


# ---- configs 1 / 4: USCKF (test/UsckfUnitTest.cpp) -----------------------------------------------
def usckf_process_noise(dt):
    """processNoiseCov(dt), test/UsckfUnitTest.cpp:51-60."""
    return 0.1 * dt * np.eye(12)


def usckf_scenario(B, seed=0, nk=3, nl=9, dt=0.01, rho=0.5, p0=0.0025, feat_cov=0.008, jitter=0.2):
    """SPD start per SURVEY 8d config 1: rho-coupled clones of a (perturbed) P0 = p0*I block, VO /
    ICP feature vectors as in test/UsckfUnitTest.cpp:198-208, measurement near h(mu)."""
    rng = np.random.default_rng(seed)
    N = 36 + nk + nl
    mu = np.concatenate([random_q(rng, B, STATE_BLOCKS * 3, vec_scale=0.5, max_angle=0.5),
                         3.34 + 0.1 * rng.normal(size=(B, nk)), 1.34 + 0.1 * rng.normal(size=(B, nl))], axis=1)
    Ps = random_spd(rng, B, 12, scale=p0, cond=1.0 / max(1e-9, 1 - jitter) if jitter > 0 else 1.0)
    Cl = np.array([[1, rho, rho * rho], [rho, 1, rho], [rho * rho, rho, 1.0]])
    P = np.zeros((B, N, N))
    P[:, :36, :36] = np.einsum("ij,bkl->bikjl", Cl, Ps).reshape(B, 36, 36)
    F = random_spd(rng, B, nk + nl, scale=feat_cov, cond=2.0)
    P[:, 36:, 36:] = F
    # weak state<->feature coupling that keeps P SPD
    Cx = rng.normal(size=(B, 36, nk + nl)) * (0.05 * np.sqrt(p0 * feat_cov))
    P[:, :36, 36:] = Cx
    P[:, 36:, :36] = Cx.transpose(0, 2, 1)
    u = np.concatenate([rng.normal(size=(B, 3)), rng.normal(size=(B, 3)) * 0.2], axis=1)
    z = mu[:, 39:39 + nk] + rng.normal(size=(B, nk)) * 0.1
    return dict(mu=mu, P=P, u=u, z=z, dt=dt, Q=usckf_process_noise(dt), R=0.01 * np.eye(nk), nk=nk, nl=nl)


def usckf_unit_test_fixture():
    """Inputs of USCKF_DYNAMIC (test/UsckfUnitTest.cpp:175-284): values only, the test asserts nothing."""
    return dict(P0_single=0.0025 * np.eye(12), dt=0.01,
                featuresVO=np.full(3, 3.34), featuresVOCov=0.008 * np.eye(3),
                featuresICP=np.full(9, 1.34), featuresICPCov=0.008 * np.eye(9),
                featuresVO2=np.full(3, 3.35), featuresVO2Cov=0.05 * np.eye(3),
                velo=np.array([100.0, 0, 0]), angvelo=np.full(3, 100.0 * D2R),
                z=np.array([2.33, 3.35, 3.35]), R=0.01 * np.eye(3))


# ---- config 3: MSCKF ------------------------------------------------------------------------------
def msckf_landmarks(nfeat=50, seed=7):
    rng = np.random.default_rng(seed)
    lm = np.empty((nfeat, 3))
    lm[:, 0:2] = rng.uniform(-2.0, 2.0, size=(nfeat, 2))
    lm[:, 2] = rng.uniform(4.0, 8.0, size=nfeat)
    return lm


def msckf_process_noise():
    """cov_process, test/MsckfUnitTest.cpp:172-177."""
    return 0.01 * np.eye(12)


def msckf_scenario(B, seed=0, k=10, nfeat=50, p_scale=1e-4, sigma_px=0.01, outlier_frac=0.0):
    """Clone poses scattered around the origin looking down +z at the shared landmarks."""
    rng = np.random.default_rng(seed)
    N = 12 + 6 * k
    blocks = STATE_BLOCKS + [0, 1] * k
    mu = random_q(rng, B, blocks, vec_scale=0.2, max_angle=0.2)
    P = random_spd(rng, B, N, scale=p_scale, cond=50.0)
    lm = msckf_landmarks(nfeat)
    # ideal measurement from the mean + pixel noise
    z = np.empty((B, 2 * nfeat))
    for f in range(nfeat):
        j = f % k
        p = mu[:, 13 + 7 * j:16 + 7 * j]
        q = mu[:, 16 + 7 * j:20 + 7 * j]
        d = lm[f][None, :] - p
        w, v = q[:, 0:1], -q[:, 1:4]          # conjugate
        t = 2.0 * np.cross(v, d)
        pc = d + w * t + np.cross(v, t)
        z[:, 2 * f] = pc[:, 0] / pc[:, 2]
        z[:, 2 * f + 1] = pc[:, 1] / pc[:, 2]
    z += rng.normal(size=z.shape) * sigma_px
    if outlier_frac > 0:
        mask = rng.uniform(size=(B, nfeat)) < outlier_frac
        z += np.repeat(mask, 2, axis=1) * rng.normal(size=z.shape) * 1.0
    dq = random_unit_quat(rng, B, max_angle=0.05)
    u = np.concatenate([rng.normal(size=(B, 3)) * 0.1, dq, rng.normal(size=(B, 3)) * 0.1,
                        rng.normal(size=(B, 3)) * 0.1], axis=1)
    return dict(mu=mu, P=P, u=u, z=z, landmarks=lm, Q=msckf_process_noise(), R=(sigma_px ** 2) * np.eye(2 * nfeat),
                k=k, nfeat=nfeat, dt=0.0)


# ---- config 5: DataModel fusion -------------------------------------------------------------------
def fusion_scenario(n, d=6, seed=99, log_spread=3.0):
    """x ~ N(0,I), C = A A^T + 1e-6 I with A ~ N(0,1)^{dxd} diag(10^U(-log_spread,0))  (SURVEY 8d
    config 5 uses log_spread = 3, i.e. cond(C) up to ~1e7: the explicit-inverse fusion of
    DataModel.hpp:54 then carries a forward error of ~cond*eps whatever the implementation)."""
    rng = np.random.default_rng(seed)

    def one():
        x = rng.normal(size=(n, d))
        A = rng.normal(size=(n, d, d)) * (10.0 ** rng.uniform(-log_spread, 0, size=(n, 1, d)))
        Cm = A @ A.transpose(0, 2, 1) + 1e-6 * np.eye(d)
        return x, 0.5 * (Cm + Cm.transpose(0, 2, 1))

    x1, C1 = one()
    x2, C2 = one()
    return dict(x1=x1, C1=C1, x2=x2, C2=C2)


def datamodel_fixture():
    """test/DataModelUnitTest.cpp:32-35 inputs and the analytic fusion answer (equal covariances)."""
    x1 = np.array([[0.0124889, 0.00171945, -0.0138983]])
    x2 = np.array([[0.0168381, 0.000632167, -0.0235605]])
    Cm = 1e-10 * np.eye(3)[None]
    return dict(x1=x1, C1=Cm, x2=x2, C2=Cm, x_expected=0.5 * (x1 + x2), C_expected=0.5 * Cm)


# ---- SURVEY 8(f) "next" rows ------------------------------------------------------------------------
def ekf_vectorize(mu):
    """ERROR_QUATERNION vectorisation of an augmented error-state mean (n x 48 -> n x 45): per single
    state pos vel (qx qy qz) gbias abias (UsckfError.hpp:521-531)."""
    mu = np.asarray(mu)
    parts = []
    for s in range(3):
        q = mu[:, 16 * s:16 * s + 16]
        parts += [q[:, 0:6], q[:, 7:10], q[:, 10:16]]
    return np.concatenate(parts, axis=1)


def ekf_scenario(n, seed=0, dt=0.01, p_scale=1e-2, cond=1e3, r_sigma=0.05, outlier_frac=0.0):
    """Row f2: error-state EKF (3 x 15-DOF augmented state).  F = I + dt * A with a random strictly
    local A per instance (what a linearised IMU error model looks like), delay-position measurement
    H = [0 .. -I_3(statek_l.pos) .. +I_3(statek_i.pos)], velocity measurement for the single update."""
    rng = np.random.default_rng(seed)
    mu = np.zeros((n, 48))
    for s in range(3):
        mu[:, 16 * s:16 * s + 6] = rng.normal(size=(n, 6))
        mu[:, 16 * s + 6:16 * s + 10] = random_unit_quat(rng, n, max_angle=1.0)
        mu[:, 16 * s + 10:16 * s + 16] = 0.01 * rng.normal(size=(n, 6))
    err = 1e-3 * rng.normal(size=(n, 45))
    P = random_spd(rng, n, 45, scale=p_scale, cond=cond)
    F = np.eye(15)[None] + dt * rng.normal(size=(n, 15, 15))
    Q = 1e-4 * dt * np.eye(15)
    H = np.zeros((3, 45))
    H[:, 15:18] = -np.eye(3)
    H[:, 30:33] = np.eye(3)
    R = (r_sigma ** 2) * np.eye(3)
    z = ekf_vectorize(mu) @ H.T + r_sigma * rng.normal(size=(n, 3))
    if outlier_frac > 0:
        bad = rng.uniform(size=n) < outlier_frac
        z[bad] += 50.0 * r_sigma * np.sign(rng.normal(size=(int(bad.sum()), 3)))
    Hs = np.zeros((3, 15))
    Hs[:, 3:6] = np.eye(3)
    zs = 1e-2 * rng.normal(size=(n, 3))
    return dict(mu=mu, err=err, P=P, F=F, Q=Q, H=H, R=R, z=z, Hs=Hs, zs=zs, dt=dt)


def safe_fusion_scenario(n, seed=5, log_spread=2.0):
    """Row f3: pairs of 3-D estimates (safeFusion is d = 3 only, DataModel.hpp:104)."""
    sc = fusion_scenario(n, d=3, seed=seed, log_spread=log_spread)
    return sc


def safe_fusion_fixture():
    """test/DataModelUnitTest.cpp:66-74: data3 = data1 (x1, 1e-10 I), data3.safeFusion(data2)."""
    fx = datamodel_fixture()
    return dict(x1=fx["x1"], C1=fx["C1"], x2=fx["x2"], C2=fx["C2"])


def deadreckon_scenario(n, seed=11, dt=0.01):
    """Row f4: body velocities at two consecutive samples (linear 3, angular 3), their shared 6x6
    covariance, and a previous pose with uncertainty (pos(3) quat(w,x,y,z); cov over [r t])."""
    rng = np.random.default_rng(seed)
    vel0 = np.concatenate([rng.normal(size=(n, 3)), 0.5 * rng.normal(size=(n, 3))], axis=1)
    vel1 = vel0 + 0.05 * rng.normal(size=(n, 6))
    A = rng.normal(size=(6, 6))
    velcov = np.zeros((6, 6))
    velcov[:3, :3] = 1e-2 * (A[:3, :3] @ A[:3, :3].T + np.eye(3))
    velcov[3:, 3:] = 1e-3 * (A[3:, 3:] @ A[3:, 3:].T + np.eye(3))
    prev_pose = np.concatenate([rng.normal(size=(n, 3)) * 5.0, random_unit_quat(rng, n, max_angle=np.pi)], axis=1)
    flip = rng.uniform(size=n) < 0.3          # both quaternion signs reach the matrix->quaternion branch logic
    prev_pose[flip, 3:] *= -1.0
    prev_cov = random_spd(rng, n, 6, scale=1e-3, cond=1e2)
    return dict(dt=dt, vel0=vel0, vel1=vel1, velcov=velcov, prev_pose=prev_pose, prev_cov=prev_cov)
