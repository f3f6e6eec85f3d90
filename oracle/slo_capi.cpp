// oracle/slo_capi.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// extern "C" surface of the CPU oracle (restatement of the reference's Eigen/MTK algorithm;
// PARITY UNPINNED for filter outputs, see slo_core.hpp).  Batches are instance-major dense
// arrays; every entry point loops the single-instance restatement over the batch, optionally
// on `nthreads` std::threads (instances are independent -- this is also how the CPU baseline
// in bench.py uses all host cores).
#include <algorithm>
#include <thread>

#include "slo_models.hpp"
#include "slo_next.hpp"

using namespace slo;

namespace {

template <typename F>
void parallel_for(int B, int nthreads, F body) {
    if (nthreads <= 1 || B <= 1) {
        for (int i = 0; i < B; ++i) body(i);
        return;
    }
    nthreads = std::min(nthreads, B);
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t)
        th.emplace_back([=]() {
            const int lo = (int)((long long)B * t / nthreads), hi = (int)((long long)B * (t + 1) / nthreads);
            for (int i = lo; i < hi; ++i) body(i);
        });
    for (auto &t : th) t.join();
}

Mat load_mat(const double *p, int r, int c) {
    Mat m(r, c);
    std::copy(p, p + (size_t)r * c, m.a.begin());
    return m;
}
void store_mat(const Mat &m, double *p) { std::copy(m.a.begin(), m.a.end(), p); }

Layout ukf_layout(int id) {
    switch (id) {
        case SLB_LAYOUT_POSE6: return Layout::pose6();
        case SLB_LAYOUT_MTK9: return Layout::mtk9();
        default: return Layout::state12();
    }
}
int pm_nu(int pm) { return pm == SLB_PM_MSCKF_DELTAPOSE ? 13 : 6; }

}  // namespace

extern "C" {

// ---- primitives (exposed for the algebra / linalg pin tests) --------------------------------
void slo_so3_exp(const double *v, double scale, double *q) { so3_exp(v, scale, q); }
void slo_so3_log(const double *q, double *v) { so3_log(q, v); }
// layout given as an array of nblk flags (1 = SO3) plus nfeat
static Layout mk(const int *so3, int nblk, int nfeat) {
    Layout l;
    for (int i = 0; i < nblk; ++i) l.so3.push_back((uint8_t)so3[i]);
    l.nfeat = nfeat;
    return l;
}
void slo_boxplus(const int *so3, int nblk, int nfeat, const double *x, const double *d, double *y) {
    Layout l = mk(so3, nblk, nfeat);
    Vec r = boxplus(l, Vec(x, x + l.qdim()), Vec(d, d + l.dof()));
    std::copy(r.begin(), r.end(), y);
}
void slo_boxminus(const int *so3, int nblk, int nfeat, const double *a, const double *b, double *d) {
    Layout l = mk(so3, nblk, nfeat);
    Vec r = boxminus(l, Vec(a, a + l.qdim()), Vec(b, b + l.qdim()));
    std::copy(r.begin(), r.end(), d);
}
void slo_set_from_vector(const int *so3, int nblk, int nfeat, const double *v, double *x) {
    Layout l = mk(so3, nblk, nfeat);
    Vec r = set_from_vector(l, Vec(v, v + l.dof()));
    std::copy(r.begin(), r.end(), x);
}
void slo_get_vectorized(const int *so3, int nblk, int nfeat, const double *x, double *v) {
    Layout l = mk(so3, nblk, nfeat);
    Vec r = get_vectorized(l, Vec(x, x + l.qdim()));
    std::copy(r.begin(), r.end(), v);
}
int slo_llt(int n, const double *P, double *L) {
    Mat l;
    int info = llt_lower(load_mat(P, n, n), l);
    store_mat(l, L);
    return info;
}
void slo_inverse(int n, const double *A, double *Ai, int fixed) {
    Mat m = load_mat(A, n, n);
    store_mat(fixed ? inverse_fixed(m) : inverse_lu(m), Ai);
}
int slo_accept_mahalanobis(double m2, int dof) { return accept_mahalanobis_distance(m2, dof) ? 1 : 0; }

// ---- ukfom::ukf -----------------------------------------------------------------------------
// mu: B x q, P: B x n x n, u: B x 6, z: B x 3; Q (n x n) and R (3 x 3) shared.
int slo_ukf_step(int layout, int pm, int mm, int B, double *mu, double *P, const double *u, double dt,
                 const double *Q, const double *z, const double *R, int gate_dof, int do_predict,
                 int do_update, int *status, int *mean_iters, int nthreads) {
    const Layout l = ukf_layout(layout);
    const int n = l.dof(), q = l.qdim(), nu = pm_nu(pm), m = 3;
    if (do_update && mm != SLB_MM_GPS_POS) return -1;
    parallel_for(B, nthreads, [&](int i) {
        Ukf f(l, Vec(mu + (size_t)i * q, mu + (size_t)(i + 1) * q), load_mat(P + (size_t)i * n * n, n, n));
        if (do_predict) f.predict(make_process_model(pm, u + (size_t)i * nu, dt), load_mat(Q, n, n));
        if (do_update) f.update(Vec(z + (size_t)i * m, z + (size_t)(i + 1) * m), mm_gps_pos, load_mat(R, m, m), gate_dof);
        std::copy(f.mu.begin(), f.mu.end(), mu + (size_t)i * q);
        store_mat(f.sigma, P + (size_t)i * n * n);
        if (status) status[i] = f.status;
        if (mean_iters) mean_iters[i] = f.last_mean_iters;
    });
    return 0;
}

// ---- localization::Usckf --------------------------------------------------------------------
// mu: B x (39+nk+nl), P: B x N x N with N = 36+nk+nl
int slo_usckf_step(int pm, int mm, int B, int nk, int nl, double *mu, double *P, const double *u,
                   double dt, const double *Q, const double *z, const double *R, int gate_dof,
                   int do_predict, int do_update, int *status, int *mean_iters, int nthreads) {
    const int N = 36 + nk + nl, q = 39 + nk + nl, nu = pm_nu(pm), m = nk;
    if (do_update && mm != SLB_MM_USCKF_VO) return -1;
    parallel_for(B, nthreads, [&](int i) {
        Usckf f(Vec(mu + (size_t)i * q, mu + (size_t)(i + 1) * q), nk, nl, load_mat(P + (size_t)i * N * N, N, N));
        if (do_predict) f.predict(make_process_model(pm, u + (size_t)i * nu, dt), load_mat(Q, 12, 12));
        if (do_update)
            f.update(Vec(z + (size_t)i * m, z + (size_t)(i + 1) * m), [nk](const Vec &a) { return mm_usckf_vo(a, nk); },
                     load_mat(R, m, m), gate_dof);
        std::copy(f.mu.begin(), f.mu.end(), mu + (size_t)i * q);
        store_mat(f.Pk, P + (size_t)i * N * N);
        if (status) status[i] = f.status;
        if (mean_iters) mean_iters[i] = f.last_mean_iters;
    });
    return 0;
}
int slo_usckf_clone(int mode, int B, int nk, int nl, double *mu, double *P, int nthreads) {
    const int N = 36 + nk + nl, q = 39 + nk + nl;
    parallel_for(B, nthreads, [&](int i) {
        Usckf f(Vec(mu + (size_t)i * q, mu + (size_t)(i + 1) * q), nk, nl, load_mat(P + (size_t)i * N * N, N, N));
        f.cloning(mode);
        std::copy(f.mu.begin(), f.mu.end(), mu + (size_t)i * q);
        store_mat(f.Pk, P + (size_t)i * N * N);
    });
    return 0;
}
// ctor #2 (Usckf.hpp:90-103): single-state in, 39-vector and 36x36 out
int slo_usckf_ctor_single(int B, const double *mu_single, const double *P_single, double *mu, double *P) {
    for (int i = 0; i < B; ++i) {
        Usckf f(Vec(mu_single + (size_t)i * 13, mu_single + (size_t)(i + 1) * 13), load_mat(P_single + (size_t)i * 144, 12, 12));
        std::copy(f.mu.begin(), f.mu.end(), mu + (size_t)i * 39);
        store_mat(f.Pk, P + (size_t)i * 36 * 36);
    }
    return 0;
}
// setMeasurement: sizes change, so in/out arrays are separate.  len = z size; R: len x len shared
int slo_usckf_set_measurement(int mode, int B, int nk, int nl, const double *mu_in, const double *P_in,
                              int len, const double *z, const double *R, double *mu_out, double *P_out) {
    const int N = 36 + nk + nl, q = 39 + nk + nl;
    const int nk2 = mode == SLB_STATEK ? len : nk, nl2 = mode == SLB_STATEK_L ? len : nl;
    const int N2 = 36 + nk2 + nl2, q2 = 39 + nk2 + nl2;
    for (int i = 0; i < B; ++i) {
        Usckf f(Vec(mu_in + (size_t)i * q, mu_in + (size_t)(i + 1) * q), nk, nl, load_mat(P_in + (size_t)i * N * N, N, N));
        f.set_measurement(mode, Vec(z + (size_t)i * len, z + (size_t)(i + 1) * len), load_mat(R, len, len));
        std::copy(f.mu.begin(), f.mu.end(), mu_out + (size_t)i * q2);
        store_mat(f.Pk, P_out + (size_t)i * N2 * N2);
    }
    return 0;
}

// ---- localization::Msckf --------------------------------------------------------------------
// mu: B x (13+7k), P: B x N x N, N = 12+6k; u: B x 13; z: B x m; landmarks: nfeat x 3 (m = 2 nfeat)
int slo_msckf_predict(int pm, int B, int k, double *mu, double *P, const double *u, double dt,
                      const double *Q, int *status, int nthreads) {
    const int N = 12 + 6 * k, q = 13 + 7 * k, nu = pm_nu(pm);
    parallel_for(B, nthreads, [&](int i) {
        Msckf f(k, Vec(mu + (size_t)i * q, mu + (size_t)(i + 1) * q), load_mat(P + (size_t)i * N * N, N, N));
        f.predict(make_process_model(pm, u + (size_t)i * nu, dt), load_mat(Q, 12, 12));
        std::copy(f.mu.begin(), f.mu.end(), mu + (size_t)i * q);
        store_mat(f.Pk, P + (size_t)i * N * N);
        if (status) status[i] = f.status;
    });
    return 0;
}
int slo_msckf_update(int mm, int B, int k, double *mu, double *P, const double *landmarks, int m,
                     const double *z, const double *R, int gate, int *outliers, int *status,
                     int *mean_iters, int nthreads) {
    const int N = 12 + 6 * k, q = 13 + 7 * k, nfeat = m / 2;
    if (mm != SLB_MM_MSCKF_REPROJ) return -1;
    parallel_for(B, nthreads, [&](int i) {
        Msckf f(k, Vec(mu + (size_t)i * q, mu + (size_t)(i + 1) * q), load_mat(P + (size_t)i * N * N, N, N));
        unsigned o = f.update(Vec(z + (size_t)i * m, z + (size_t)(i + 1) * m),
                              [=](const Vec &s) { return mm_msckf_reproj(s, k, landmarks, nfeat); },
                              load_mat(R, m, m), gate != 0);
        std::copy(f.mu.begin(), f.mu.end(), mu + (size_t)i * q);
        store_mat(f.Pk, P + (size_t)i * N * N);
        if (outliers) outliers[i] = (int)o;
        if (status) status[i] = f.status;
        if (mean_iters) mean_iters[i] = f.last_mean_iters;
    });
    return 0;
}
// checkSigmaPoints() Usckf.hpp:769-789 (kind 2) / Msckf.hpp:818-838 (kind 3): sigma points of (mu, Pk), their manifold
// mean muX and covariance Pktest.  The reference asserts max|Pktest - Pk| <= 1e-6 and mu == muX; here the two quantities
// are returned (diff[2i] = max|Pktest - Pk|, diff[2i+1] = |muX [-] mu|_inf) with flags bit 0: covariance off by more
// than 1e-6, bit 1: mean moved by more than 1e-12 (the reference's exact == is rounding noise), bit 2: LLT failed.
int slo_check_sigma_points(int kind, int B, int nk, int nl, int k, const double *mu, const double *P, int *flags,
                           double *diff, int nthreads) {
    const int N = kind == 2 ? 36 + nk + nl : 12 + 6 * k, q = kind == 2 ? 39 + nk + nl : 13 + 7 * k;
    if (kind != 2 && kind != 3) return -1;
    parallel_for(B, nthreads, [&](int i) {
        const Vec m(mu + (size_t)i * q, mu + (size_t)(i + 1) * q);
        const Mat Pk = load_mat(P + (size_t)i * N * N, N, N);
        std::vector<Vec> X;
        Layout lay;
        int info;
        if (kind == 2) {
            Usckf f(m, nk, nl, Pk);
            lay = f.aug();
            info = f.sigma_points_aug(Vec(N, 0.0), X);
        } else {
            lay = Layout::multi(k);
            info = sigma_points_vec(lay, m, Vec(N, 0.0), Pk, X);
        }
        int fl = info >= 0 ? 4 : 0, st = 0;
        double dP = 0.0, dm = 0.0;
        if (info < 0) {
            const Vec muX = mean_manifold(lay, X, &st);
            const Mat Pt = cov_manifold(lay, muX, X);
            for (int r = 0; r < N; ++r)
                for (int c = 0; c <= r; ++c) dP = std::max(dP, std::fabs(Pt(r, c) - Pk(r, c)));
            for (double d : boxminus(lay, muX, m)) dm = std::max(dm, std::fabs(d));
            if (dP > 1e-6) fl |= 1;
            if (dm > 1e-12) fl |= 2;
        }
        flags[i] = fl;
        if (diff) { diff[2 * i] = dP; diff[2 * i + 1] = dm; }
    });
    return 0;
}
// removeOutliers alone (Msckf.hpp:723-754, quirk Q6): returns the kept original row indices
int slo_msckf_remove_outliers(int m, int N, const double *innov, const double *S, int *kept, int *nkept) {
    Msckf f(0, Layout::state12().identity(), Mat(12, 12));
    Vec in(innov, innov + m);
    Mat s = load_mat(S, m, m), c(N, m);
    std::vector<int> k;
    unsigned o = f.remove_outliers(in, c, s, &k);
    std::copy(k.begin(), k.end(), kept);
    *nkept = (int)k.size();
    return (int)o;
}

// ---- localization::DataModel ----------------------------------------------------------------
// x*: n x d, C*: n x d x d; op: 0 fusion, +1 operator+, -1 operator-
int slo_datamodel(int op, int d, long long n, const double *x1, const double *C1, const double *x2,
                  const double *C2, double *xo, double *Co, int nthreads) {
    parallel_for((int)n, nthreads, [&](int i) {
        DataModel a(Vec(x1 + (size_t)i * d, x1 + (size_t)(i + 1) * d), load_mat(C1 + (size_t)i * d * d, d, d));
        DataModel b(Vec(x2 + (size_t)i * d, x2 + (size_t)(i + 1) * d), load_mat(C2 + (size_t)i * d * d, d, d));
        if (op == 0) a.fusion(b);
        else if (op > 0) a = a.plus(b);
        else a = a.minus(b);
        std::copy(a.data.begin(), a.data.end(), xo + (size_t)i * d);
        store_mat(a.Cov, Co + (size_t)i * d * d);
    });
    return 0;
}
// default constructor (DataModel.hpp:32-36)
void slo_datamodel_default(int d, double *x, double *C) {
    DataModel a(d);
    std::copy(a.data.begin(), a.data.end(), x);
    store_mat(a.Cov, C);
}

// ---- SURVEY 8(f) next rows (slo_next.hpp) ------------------------------------------------------
// f2: error-state EKF, batches instance-major: mu n x 48, err n x 45, P n x 45 x 45
int slo_ekf_predict(long long n, double *err, double *P, const double *F, const double *Q, int nthreads) {
    const Mat Qm = load_mat(Q, EKF_NS, EKF_NS);
    parallel_for((int)n, nthreads, [&](int i) {
        Vec e(err + (size_t)i * EKF_NA, err + (size_t)(i + 1) * EKF_NA);
        Mat Pm = load_mat(P + (size_t)i * EKF_NA * EKF_NA, EKF_NA, EKF_NA);
        ekf_predict(e, Pm, load_mat(F + (size_t)i * EKF_NS * EKF_NS, EKF_NS, EKF_NS), Qm);
        std::copy(e.begin(), e.end(), err + (size_t)i * EKF_NA);
        store_mat(Pm, P + (size_t)i * EKF_NA * EKF_NA);
    });
    return 0;
}
// H: m x 45 shared, R: m x m shared, z: n x m; ret: n x m; accepted: n
int slo_ekf_update(long long n, int m, const double *mu, double *P, const double *z, const double *H, const double *R,
                   int gate, double *ret, int *accepted, int nthreads) {
    const Mat Hm = load_mat(H, m, EKF_NA), Rm = load_mat(R, m, m);
    parallel_for((int)n, nthreads, [&](int i) {
        Mat Pm = load_mat(P + (size_t)i * EKF_NA * EKF_NA, EKF_NA, EKF_NA);
        Vec r;
        accepted[i] = ekf_update(mu + (size_t)i * EKF_QA, Pm, Vec(z + (size_t)i * m, z + (size_t)(i + 1) * m), Hm, Rm, gate, r) ? 1 : 0;
        std::copy(r.begin(), r.end(), ret + (size_t)i * m);
        store_mat(Pm, P + (size_t)i * EKF_NA * EKF_NA);
    });
    return 0;
}
// H: m x 15 shared
int slo_ekf_single_update(long long n, int m, double *mu, const double *err, double *P, const double *z, const double *H,
                          const double *R, int gate, int *accepted, int nthreads) {
    const Mat Hm = load_mat(H, m, EKF_NS), Rm = load_mat(R, m, m);
    parallel_for((int)n, nthreads, [&](int i) {
        Mat Pm = load_mat(P + (size_t)i * EKF_NA * EKF_NA, EKF_NA, EKF_NA);
        accepted[i] = ekf_single_update(mu + (size_t)i * EKF_QA, Vec(err + (size_t)i * EKF_NA, err + (size_t)(i + 1) * EKF_NA), Pm,
                                        Vec(z + (size_t)i * m, z + (size_t)(i + 1) * m), Hm, Rm, gate) ? 1 : 0;
        store_mat(Pm, P + (size_t)i * EKF_NA * EKF_NA);
    });
    return 0;
}
int slo_ekf_clone(long long n, double *mu, double *err, double *P) {
    for (long long i = 0; i < n; ++i) {
        Vec e(err + (size_t)i * EKF_NA, err + (size_t)(i + 1) * EKF_NA);
        Mat Pm = load_mat(P + (size_t)i * EKF_NA * EKF_NA, EKF_NA, EKF_NA);
        ekf_clone(mu + (size_t)i * EKF_QA, e, Pm);
        std::copy(e.begin(), e.end(), err + (size_t)i * EKF_NA);
        store_mat(Pm, P + (size_t)i * EKF_NA * EKF_NA);
    }
    return 0;
}
// f1: Msckf::update EKF flavour; same array conventions as slo_msckf_update
int slo_msckf_update_ekf(int mm, int B, int k, double *mu, double *P, const double *landmarks, int m, const double *z,
                         const double *R, int gate, int *outliers, int *status, int nthreads) {
    const int N = 12 + 6 * k, q = 13 + 7 * k, nfeat = m / 2;
    if (mm != SLB_MM_MSCKF_REPROJ) return -1;
    parallel_for(B, nthreads, [&](int i) {
        Msckf f(k, Vec(mu + (size_t)i * q, mu + (size_t)(i + 1) * q), load_mat(P + (size_t)i * N * N, N, N));
        unsigned o = msckf_update_ekf(f, Vec(z + (size_t)i * m, z + (size_t)(i + 1) * m),
                                      [=](const Vec &s, Mat &H) { return mm_msckf_reproj_jac(s, k, landmarks, nfeat, H); },
                                      load_mat(R, m, m), gate != 0);
        std::copy(f.mu.begin(), f.mu.end(), mu + (size_t)i * q);
        store_mat(f.Pk, P + (size_t)i * N * N);
        if (outliers) outliers[i] = (int)o;
        if (status) status[i] = f.status;
    });
    return 0;
}
// measurement + Jacobian of the reprojection model for one state (H: m x N)
void slo_msckf_reproj_jac(int k, const double *mu, const double *landmarks, int nfeat, double *z, double *H) {
    Mat Hm;
    Vec zz = mm_msckf_reproj_jac(Vec(mu, mu + 13 + 7 * k), k, landmarks, nfeat, Hm);
    std::copy(zz.begin(), zz.end(), z);
    store_mat(Hm, H);
}
void slo_householder_qr(int rows, int cols, const double *A, double *QR, double *tau, double *thinQ) {
    Mat a = load_mat(A, rows, cols);
    Vec t;
    householder_qr(a, t);
    store_mat(a, QR);
    std::copy(t.begin(), t.end(), tau);
    store_mat(householder_thin_q(a, t, cols), thinQ);
}
// f3: safeFusion, d = 3
int slo_safe_fusion(long long n, const double *x1, const double *C1, const double *x2, const double *C2, double *xo,
                    double *Co, int nthreads) {
    parallel_for((int)n, nthreads, [&](int i) {
        Vec a(x1 + (size_t)i * 3, x1 + (size_t)(i + 1) * 3);
        Mat Ca = load_mat(C1 + (size_t)i * 9, 3, 3);
        safe_fusion3(a, Ca, Vec(x2 + (size_t)i * 3, x2 + (size_t)(i + 1) * 3), load_mat(C2 + (size_t)i * 9, 3, 3));
        std::copy(a.begin(), a.end(), xo + (size_t)i * 3);
        store_mat(Ca, Co + (size_t)i * 9);
    });
    return 0;
}
void slo_jacobi_svd(int n, const double *A, double *U, double *sv) {
    Mat Um;
    Vec s;
    jacobi_svd(load_mat(A, n, n), Um, s);
    store_mat(Um, U);
    std::copy(s.begin(), s.end(), sv);
}
// f4: poses n x 7 (pos, quat wxyz), covariances n x 6 x 6 over [r t]
int slo_transform_compose(long long n, const double *pose2, const double *cov2, const double *pose1, const double *cov1,
                          double *pose_out, double *cov_out, int nthreads) {
    parallel_for((int)n, nthreads, [&](int i) {
        Mat co;
        transform_compose(pose2 + (size_t)i * 7, load_mat(cov2 + (size_t)i * 36, 6, 6), pose1 + (size_t)i * 7,
                          load_mat(cov1 + (size_t)i * 36, 6, 6), pose_out + (size_t)i * 7, co);
        store_mat(co, cov_out + (size_t)i * 36);
    });
    return 0;
}
// vel0 / vel1: n x 6 (linear, angular) at t and t - dt; velcov: 6 x 6 shared
int slo_dr_update_pose(long long n, double dt, const double *vel0, const double *vel1, const double *velcov,
                       const double *prev_pose, const double *prev_cov, double *post_pose, double *post_cov,
                       double *delta_pose, double *delta_cov, int nthreads) {
    const Mat vc = load_mat(velcov, 6, 6);
    parallel_for((int)n, nthreads, [&](int i) {
        Mat pc, dc;
        dr_update_pose(dt, vel0 + (size_t)i * 6, vel1 + (size_t)i * 6, vc, prev_pose + (size_t)i * 7,
                       load_mat(prev_cov + (size_t)i * 36, 6, 6), post_pose + (size_t)i * 7, pc, delta_pose + (size_t)i * 7, dc);
        store_mat(pc, post_cov + (size_t)i * 36);
        store_mat(dc, delta_cov + (size_t)i * 36);
    });
    return 0;
}
void slo_dr_update_attitude(double dt, const double *w0, const double *w1, double *dq) { dr_update_attitude(dt, w0, w1, dq); }

int slo_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
