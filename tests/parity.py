"""Parity metric shared by the GPU tests.  north_star: relative error <= 1e-9 on state means and
covariances per step (<= 1e-6 after 10k free-running steps)."""
import numpy as np

STEP_TOL = 1e-9
LONG_TOL = 1e-6


def mean_error(slo, blocks, mu_a, mu_b, nfeat=0):
    """max over instances of |a [-] b|_inf / max(1, |b|_inf): tangent-space distance, relative."""
    worst = 0.0
    for a, b in zip(mu_a, mu_b):
        d = slo.boxminus(blocks, a, b, nfeat)
        worst = max(worst, float(np.max(np.abs(d))) / max(1.0, float(np.max(np.abs(b)))))
    return worst


def cov_error(P_a, P_b):
    """max over instances of |Pa - Pb|_max / |Pb|_max."""
    num = np.max(np.abs(P_a - P_b), axis=(1, 2))
    den = np.max(np.abs(P_b), axis=(1, 2))
    return float(np.max(num / den))


def assert_parity(slo, blocks, mu_gpu, P_gpu, mu_ref, P_ref, nfeat=0, tol=STEP_TOL, mask=None):
    if mask is not None:
        mu_gpu, P_gpu, mu_ref, P_ref = mu_gpu[mask], P_gpu[mask], mu_ref[mask], P_ref[mask]
    em = mean_error(slo, blocks, mu_gpu, mu_ref, nfeat)
    ec = cov_error(P_gpu, P_ref)
    assert em <= tol, "mean parity %.3e > %.1e" % (em, tol)
    assert ec <= tol, "covariance parity %.3e > %.1e" % (ec, tol)
    return em, ec


def symmetrize_lower(P):
    """The engine stores the lower triangle (what the reference's LLT reads, Q8); compare like with like."""
    L = np.tril(P)
    return L + np.tril(P, -1).transpose(0, 2, 1)
