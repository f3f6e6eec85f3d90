// oracle/slo_next.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the SURVEY section 8(f) "next" rows:
//   f1  Msckf::update, EKF flavour with QR compression  src/filters/Msckf.hpp:297-349,756-816
//   f2  error-state EKF with Joseph-form update   src/filters/UsckfError.hpp:87-137,322-384,489-571,573-603
//   f3  DataModel<double,3>::safeFusion            src/core/DataModel.hpp:62-130
//   f4  DeadReckon::updateAttitude / updatePose and TransformWithUncertainty::operator*
//                                                  src/core/DeadReckon.hpp:30-79,246-286, src/core/Transform.cpp:35-136,215-254
// PARITY UNPINNED (see slo_core.hpp): the reference asserts none of these outputs.  Third-party
// algorithms restated from their published sources (Eigen 3.3 series; no version pinned by the
// reference, src/CMakeLists.txt:26):
//   * Eigen::HouseholderQR (makeHouseholderInPlace: beta = -sign(c0) |x|, tau = (beta - c0)/beta, tau = 0 for a
//     zero tail; applyHouseholderOnTheLeft; householderQ() applied to a thin identity)  -- Msckf.hpp:799-805
//   * Eigen::JacobiSVD for real square matrices (two-sided Jacobi, sweep order p = 1.., q < p,
//     threshold 2 eps max|diag|, sign fix-up on U, descending sort)  -- DataModel.hpp:83,97
//   * Eigen::Quaternion(Matrix3) (Shoemake's trace method), Quaternion::toRotationMatrix,
//     AngleAxis(Quaternion) (angle = 2 atan2(|v|, |w|), axis sign follows w), Quaternion::normalize
//                                                                        -- Transform.cpp:37-50,225-228
// safeFusion's result depends on JacobiSVD's sign and ordering conventions because the reference forms
// T = U2^T sqrt(D1) U1 (DataModel.hpp:106; the textbook algorithm has U1^T): both are reproduced as written.
#pragma once
#include <cfloat>

#include "slo_filters.hpp"
#include "slo_models.hpp"

namespace slo {

// Eigen::Quaternion::toRotationMatrix (used by f1 and f4)
inline Mat quat_to_rot_fwd(const double q[4]) {
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    Mat R(3, 3);
    R(0, 0) = 1 - (tyy + tzz); R(0, 1) = txy - twz; R(0, 2) = txz + twy;
    R(1, 0) = txy + twz; R(1, 1) = 1 - (txx + tzz); R(1, 2) = tyz - twx;
    R(2, 0) = txz - twy; R(2, 1) = tyz + twx; R(2, 2) = 1 - (txx + tyy);
    return R;
}


// ==========================================================================================
// f1: Msckf::update, EKF flavour (Msckf.hpp:297-349) with removeOutliers on H (:756-792, index quirk Q6 and the
// never-compacted `information` matrix reproduced) and reduceDimension (:794-816).
// ==========================================================================================
// Jacobian of mm_msckf_reproj with respect to the tangent perturbation of the multi-state (p' = p + dp,
// q' = q exp(dtheta)):  pc' = pc - R^T dp + [pc]x dtheta,  z = (x/z, y/z).  The reference's EKF update receives H
// from the caller's functor h(mu, H) (:311); this is the functor of the builder-defined reprojection model.
inline Vec mm_msckf_reproj_jac(const Vec &s, int k, const double *lm, int nfeat, Mat &H) {
    const int N = 12 + 6 * k;
    H = Mat(2 * nfeat, N);
    Vec z(2 * nfeat);
    for (int f = 0; f < nfeat; ++f) {
        const int j = f % k;
        const double *p = &s[13 + 7 * j], *q = &s[13 + 7 * j + 3];
        const double d[3] = {lm[3 * f] - p[0], lm[3 * f + 1] - p[1], lm[3 * f + 2] - p[2]};
        double qc[4], pc[3];
        quat_conj(q, qc);
        quat_rotate(qc, d, pc);
        z[2 * f] = pc[0] / pc[2];
        z[2 * f + 1] = pc[1] / pc[2];
        const Mat R = quat_to_rot_fwd(q);
        double J[3][6];  // d pc / d (dp, dtheta)
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) J[a][b] = -R(b, a);
        J[0][3] = 0.0; J[0][4] = -pc[2]; J[0][5] = pc[1];
        J[1][3] = pc[2]; J[1][4] = 0.0; J[1][5] = -pc[0];
        J[2][3] = -pc[1]; J[2][4] = pc[0]; J[2][5] = 0.0;
        const double iz = 1.0 / pc[2];
        for (int c = 0; c < 6; ++c) {
            H(2 * f, 12 + 6 * j + c) = iz * (J[0][c] - z[2 * f] * J[2][c]);
            H(2 * f + 1, 12 + 6 * j + c) = iz * (J[1][c] - z[2 * f + 1] * J[2][c]);
        }
    }
    return z;
}

// Eigen::HouseholderQR, in place: R in the upper triangle, essential parts below, coefficients in tau.
inline void householder_qr(Mat &A, Vec &tau) {
    const int rows = A.r, cols = A.c, size = std::min(rows, cols);
    tau.assign(size, 0.0);
    for (int k = 0; k < size; ++k) {
        const double c0 = A(k, k);
        double tail = 0.0;
        for (int i = k + 1; i < rows; ++i) tail += A(i, k) * A(i, k);
        double beta;
        if (tail <= DBL_MIN) {
            tau[k] = 0.0;
            beta = c0;
            for (int i = k + 1; i < rows; ++i) A(i, k) = 0.0;
        } else {
            beta = std::sqrt(c0 * c0 + tail);
            if (c0 >= 0.0) beta = -beta;
            for (int i = k + 1; i < rows; ++i) A(i, k) /= (c0 - beta);
            tau[k] = (beta - c0) / beta;
        }
        A(k, k) = beta;
        for (int j = k + 1; j < cols; ++j) {  // applyHouseholderOnTheLeft
            double tmp = 0.0;
            for (int i = k + 1; i < rows; ++i) tmp += A(i, k) * A(i, j);
            tmp += A(k, j);
            A(k, j) -= tau[k] * tmp;
            for (int i = k + 1; i < rows; ++i) A(i, j) -= tau[k] * A(i, k) * tmp;
        }
    }
}
// householderQ() * Identity(rows, ncols): reflectors applied last to first
inline Mat householder_thin_q(const Mat &QR, const Vec &tau, int ncols) {
    const int rows = QR.r;
    Mat Q(rows, ncols);
    for (int i = 0; i < std::min(rows, ncols); ++i) Q(i, i) = 1.0;
    for (int k = (int)tau.size() - 1; k >= 0; --k)
        for (int j = 0; j < ncols; ++j) {
            double tmp = 0.0;
            for (int i = k + 1; i < rows; ++i) tmp += QR(i, k) * Q(i, j);
            tmp += Q(k, j);
            Q(k, j) -= tau[k] * tmp;
            for (int i = k + 1; i < rows; ++i) Q(i, j) -= tau[k] * QR(i, k) * tmp;
        }
    return Q;
}

enum { ST_QR_ROWS = 16 };  // reduceDimension needs rows >= DOF (R.block(0,0,N,N), :808: out of range otherwise)

// returns the outlier count (:349); hj(mu, H) is the measurement functor that also fills the Jacobian
inline unsigned msckf_update_ekf(Msckf &f, const Vec &z, const std::function<Vec(const Vec &, Mat &)> &hj, const Mat &R0,
                                 bool gate) {
    const Layout lm = f.multi();
    const int N = f.dof();
    Mat H;
    const Vec mean_z = hj(f.mu, H);                                                        // :311
    Vec innov(z.size());
    for (size_t i = 0; i < z.size(); ++i) innov[i] = z[i] - mean_z[i];                     // :313
    Mat R = R0;
    unsigned outliers = 0;
    if (gate) {                                                                            // :315 -> :756-792
        const Mat info = inverse_lu(add(matmul(matmul(H, f.Pk), H.transpose()), R));       // :765-766, never compacted
        std::vector<int> kept(innov.size());
        for (size_t i = 0; i < kept.size(); ++i) kept[i] = (int)i;
        unsigned i = 0;
        while (i < kept.size() / 2) {
            const double v0 = innov[kept[2 * i]], v1 = innov[kept[2 * i + 1]];
            const double i00 = info(2 * i, 2 * i), i01 = info(2 * i, 2 * i + 1), i10 = info(2 * i + 1, 2 * i),
                         i11 = info(2 * i + 1, 2 * i + 1);
            const double m2 = v0 * (i00 * v0 + i01 * v1) + v1 * (i10 * v0 + i11 * v1);     // :773
            if (!accept_mahalanobis_distance(m2, 2)) {
                Msckf::remove_at(kept, 2 * i);                                             // :778-781 (Q6)
                Msckf::remove_at(kept, 2 * i + 1);
                ++outliers;
            } else {
                ++i;
            }
        }
        const int m = (int)kept.size();
        Vec in2(m);
        Mat H2(m, N), R2(m, m);
        for (int p = 0; p < m; ++p) {
            in2[p] = innov[kept[p]];
            for (int c = 0; c < N; ++c) H2(p, c) = H(kept[p], c);
            for (int q = 0; q < m; ++q) R2(p, q) = R(kept[p], kept[q]);
        }
        innov = in2; H = H2; R = R2;
    }
    if (!innov.empty()) {                                                                  // :322
        if ((int)innov.size() < N) {
            f.status |= ST_QR_ROWS;
            return outliers;
        }
        Vec tau;                                                                           // :794-816
        Mat QR = H;
        householder_qr(QR, tau);
        const Mat thinQ = householder_thin_q(QR, tau, N);
        Mat Hr(N, N);
        for (int i = 0; i < N; ++i)
            for (int j = i; j < N; ++j) Hr(i, j) = QR(i, j);
        const Mat Qt = thinQ.transpose();
        innov = matvec(Qt, innov);
        R = matmul(matmul(Qt, R), thinQ);
        const Mat Ht = Hr.transpose();
        const Mat S = add(matmul(matmul(Hr, f.Pk), Ht), R);                                // :330
        const Mat K = matmul(matmul(f.Pk, Ht), inverse_lu(S));                             // :331
        f.Pk = sub(f.Pk, matmul(matmul(K, S), K.transpose()));                             // :336
        f.mu = boxplus(lm, f.mu, matvec(K, innov));                                        // :337
    }
    return outliers;
}

// ==========================================================================================
// f2: error-state EKF (UsckfError.hpp).  Builder-defined 15-DOF single state (the reference's type is
// not in its tree): pos vel orient gbias abias, UsckfError.hpp:527-531; q-vector per single state =
// pos3 vel3 quat(w,x,y,z) gbias3 abias3 (16 doubles); augmented = statek | statek_l | statek_i.
// The ERROR_QUATERNION vectorisation of a state is pos vel (qx qy qz) gbias abias (:521-524 builds the
// error quaternion as (1, x, y, z) from exactly those three slots).
// ==========================================================================================
constexpr int EKF_NS = 15, EKF_NA = 45, EKF_QS = 16, EKF_QA = 48;

inline Vec ekf_vectorize(const double *mu48) {
    Vec x(EKF_NA);
    for (int s = 0; s < 3; ++s) {
        const double *q = mu48 + EKF_QS * s;
        double *o = x.data() + EKF_NS * s;
        for (int i = 0; i < 6; ++i) o[i] = q[i];
        o[6] = q[7]; o[7] = q[8]; o[8] = q[9];
        for (int i = 0; i < 6; ++i) o[9 + i] = q[10 + i];
    }
    return x;
}

// ekfPredict(F, Q) UsckfError.hpp:87-137.  err: 45-vector mu_error (vectorised), P: 45x45.
inline void ekf_predict(Vec &err, Mat &P, const Mat &F, const Mat &Q) {
    const int o = 2 * EKF_NS;
    Vec ei(err.begin() + o, err.end());
    ei = matvec(F, ei);  // :93
    std::copy(ei.begin(), ei.end(), err.begin() + o);
    const Mat Ft = F.transpose();
    Mat Pk = add(matmul(matmul(F, P.block(o, o, EKF_NS, EKF_NS)), Ft), Q);  // :96
    P.set_block(o, o, Pk);
    for (int b = 0; b < 2; ++b) {
        const int ob = EKF_NS * b;
        P.set_block(ob, o, matmul(P.block(ob, o, EKF_NS, EKF_NS), Ft));  // :109-116
    }
    for (int b = 0; b < 2; ++b) {
        const int ob = EKF_NS * b;
        P.set_block(o, ob, matmul(F, P.block(o, ob, EKF_NS, EKF_NS)));  // :119-126
    }
}

// (I - K H) P (I - K H)^T + K R K^T, then 0.5 (P + P^T)   UsckfError.hpp:356-359,525-528
inline Mat joseph(const Mat &P, const Mat &K, const Mat &H, const Mat &R) {
    const int n = P.r;
    const Mat IKH = sub(Mat::identity(n), matmul(K, H));
    Mat Pn = add(matmul(matmul(IKH, P), IKH.transpose()), matmul(matmul(K, R), K.transpose()));
    Mat Ps(n, n);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) Ps(i, j) = 0.5 * (Pn(i, j) + Pn(j, i));
    return Ps;
}

// ekfUpdate(z, H, R, mt) UsckfError.hpp:322-384.  Only Pk_error changes (x_hat is a local, :354); returns
// true when the update was accepted; `ret` is the function's return value (zeros / the innovation).
// gate: 0 = accept any, otherwise the 5% table with dof = m - 1 (:350, quirk: size()-1).
inline bool ekf_update(const double *mu48, Mat &P, const Vec &z, const Mat &H, const Mat &R, int gate, Vec &ret) {
    const int m = (int)z.size();
    const Vec x_hat = ekf_vectorize(mu48);
    const Mat Ht = H.transpose();
    const Mat S = add(matmul(matmul(H, P), Ht), R);
    const Mat Si = inverse_fixed(S);
    const Mat K = matmul(matmul(P, Ht), Si);
    const Vec hx = matvec(H, x_hat);
    Vec innov(m);
    for (int i = 0; i < m; ++i) innov[i] = z[i] - hx[i];
    const Vec t = matvec(Si, innov);
    double m2 = 0;
    for (int i = 0; i < m; ++i) m2 += innov[i] * t[i];
    const bool ok = gate == 0 ? true : accept_mahalanobis_distance(m2, m - 1);
    if (ok) {
        P = joseph(P, K, H, R);
        ret.assign(m, 0.0);
    } else {
        ret = innov;
    }
    return ok;
}

// ekfSingleUpdate(z, H, R, mt) UsckfError.hpp:489-571.  H: m x 15.  The correction is applied to
// mu_state.statek_i whether or not the gate accepted (:553-568); mu_error is not written (:503 local).
inline bool ekf_single_update(double *mu48, const Vec &err, Mat &P, const Vec &z, const Mat &H, const Mat &R, int gate) {
    const int m = (int)z.size(), o = 2 * EKF_NS;
    Vec xk(err.begin() + o, err.end());
    Mat Pk = P.block(o, o, EKF_NS, EKF_NS);
    const Mat Ht = H.transpose();
    const Mat S = add(matmul(matmul(H, Pk), Ht), R);
    const Mat Si = inverse_fixed(S);
    const Mat K = matmul(matmul(Pk, Ht), Si);
    const Vec hx = matvec(H, xk);
    Vec innov(m);
    for (int i = 0; i < m; ++i) innov[i] = z[i] - hx[i];
    const Vec t = matvec(Si, innov);
    double m2 = 0;
    for (int i = 0; i < m; ++i) m2 += innov[i] * t[i];
    const bool ok = gate == 0 ? true : accept_mahalanobis_distance(m2, m - 1);
    if (ok) {
        const Vec kd = matvec(K, innov);
        for (int i = 0; i < EKF_NS; ++i) xk[i] += kd[i];
        Pk = joseph(Pk, K, H, R);
    }
    P.set_block(o, o, Pk);
    double *s = mu48 + 2 * EKF_QS;
    for (int i = 0; i < 3; ++i) { s[i] += xk[i]; s[3 + i] += xk[3 + i]; }
    const double qe[4] = {1.0, xk[6], xk[7], xk[8]};
    double q[4];
    quat_mul(s + 6, qe, q);
    const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);  // Eigen normalize(): coeffs / norm
    for (int i = 0; i < 4; ++i) s[6 + i] = q[i] / n;
    for (int i = 0; i < 6; ++i) s[10 + i] += xk[9 + i];
    return ok;
}

// cloning() UsckfError.hpp:573-603: every one of the nine 15x15 blocks becomes Pk_i.
inline void ekf_clone(double *mu48, Vec &err, Mat &P) {
    for (int i = 0; i < EKF_QS; ++i) mu48[EKF_QS + i] = mu48[2 * EKF_QS + i];
    for (int i = 0; i < EKF_QS; ++i) mu48[i] = mu48[EKF_QS + i];
    for (int i = 0; i < EKF_NS; ++i) err[EKF_NS + i] = err[2 * EKF_NS + i];
    for (int i = 0; i < EKF_NS; ++i) err[i] = err[EKF_NS + i];
    const Mat Pk = P.block(2 * EKF_NS, 2 * EKF_NS, EKF_NS, EKF_NS);
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) P.set_block(EKF_NS * a, EKF_NS * b, Pk);
}

// ==========================================================================================
// f3: safeFusion.  Eigen::JacobiSVD<MatrixXd>(A, ComputeThinU) for a square real A.
// ==========================================================================================
struct JRot { double c, s; };  // Eigen::JacobiRotation: J = [c s; -s c]
inline JRot jrot_transpose(JRot j) { return {j.c, -j.s}; }
inline JRot jrot_mul(JRot a, JRot b) { return {a.c * b.c - a.s * b.s, a.c * b.s + a.s * b.c}; }
// internal::apply_rotation_in_the_plane(x, y, j): x' = c x + s y, y' = -s x + c y
inline void rot_rows(Mat &M, int p, int q, JRot j) {
    for (int i = 0; i < M.c; ++i) {
        const double x = M(p, i), y = M(q, i);
        M(p, i) = j.c * x + j.s * y;
        M(q, i) = -j.s * x + j.c * y;
    }
}
inline void rot_cols(Mat &M, int p, int q, JRot j) {  // applyOnTheRight(p, q, j) uses j.transpose()
    const JRot t = jrot_transpose(j);
    for (int i = 0; i < M.r; ++i) {
        const double x = M(i, p), y = M(i, q);
        M(i, p) = t.c * x + t.s * y;
        M(i, q) = -t.s * x + t.c * y;
    }
}
// JacobiRotation::makeJacobi(x, y, z) for the symmetric 2x2 [x y; y z]
inline JRot make_jacobi(double x, double y, double z) {
    const double deno = 2.0 * std::fabs(y);
    if (deno < DBL_MIN) return {1.0, 0.0};
    const double tau = (x - z) / deno;
    const double w = std::sqrt(tau * tau + 1.0);
    const double t = tau > 0.0 ? 1.0 / (tau + w) : 1.0 / (tau - w);
    const double sign_t = t > 0.0 ? 1.0 : -1.0;
    const double n = 1.0 / std::sqrt(t * t + 1.0);
    return {n, -sign_t * (y / std::fabs(y)) * std::fabs(t) * n};
}
// internal::real_2x2_jacobi_svd
inline void real_2x2_jacobi_svd(const Mat &W, int p, int q, JRot &jl, JRot &jr) {
    Mat m(2, 2);
    m(0, 0) = W(p, p); m(0, 1) = W(p, q); m(1, 0) = W(q, p); m(1, 1) = W(q, q);
    JRot rot1;
    const double t = m(0, 0) + m(1, 1), d = m(1, 0) - m(0, 1);
    if (std::fabs(d) < DBL_MIN) {
        rot1 = {1.0, 0.0};
    } else {
        const double u = t / d;
        const double tmp = std::sqrt(1.0 + u * u);
        rot1 = {u / tmp, 1.0 / tmp};
    }
    rot_rows(m, 0, 1, rot1);
    jr = make_jacobi(m(0, 0), m(0, 1), m(1, 1));
    jl = jrot_mul(rot1, jrot_transpose(jr));
}
inline void jacobi_svd(const Mat &A, Mat &U, Vec &sv) {
    const int n = A.r;
    const double precision = 2.0 * DBL_EPSILON, consider_zero = DBL_MIN;
    double scale = 0.0;
    for (double v : A.a) scale = std::max(scale, std::fabs(v));
    if (scale == 0.0) scale = 1.0;
    Mat W(n, n);
    for (size_t i = 0; i < W.a.size(); ++i) W.a[i] = A.a[i] / scale;
    U = Mat::identity(n);
    double max_diag = 0.0;
    for (int i = 0; i < n; ++i) max_diag = std::max(max_diag, std::fabs(W(i, i)));
    bool finished = false;
    while (!finished) {
        finished = true;
        for (int p = 1; p < n; ++p)
            for (int q = 0; q < p; ++q) {
                const double thr = std::max(consider_zero, precision * max_diag);
                if (std::fabs(W(p, q)) > thr || std::fabs(W(q, p)) > thr) {
                    finished = false;
                    JRot jl, jr;
                    real_2x2_jacobi_svd(W, p, q, jl, jr);
                    rot_rows(W, p, q, jl);
                    rot_cols(U, p, q, jrot_transpose(jl));
                    rot_cols(W, p, q, jr);
                    max_diag = std::max(max_diag, std::max(std::fabs(W(p, p)), std::fabs(W(q, q))));
                }
            }
    }
    sv.assign(n, 0.0);
    for (int i = 0; i < n; ++i) {
        const double a = std::fabs(W(i, i));
        sv[i] = a;
        if (a != 0.0) {
            const double f = W(i, i) / a;
            for (int r = 0; r < n; ++r) U(r, i) *= f;
        }
    }
    for (int i = 0; i < n; ++i) sv[i] *= scale;
    for (int i = 0; i < n; ++i) {
        int pos = i;
        for (int k = i + 1; k < n; ++k)
            if (sv[k] > sv[pos]) pos = k;
        if (sv[pos] == 0.0) break;
        if (pos != i) {
            std::swap(sv[i], sv[pos]);
            for (int r = 0; r < n; ++r) std::swap(U(r, i), U(r, pos));
        }
    }
}

// DataModel<double,3>::safeFusion(data2) DataModel.hpp:62-130 (valid for _DIM = 3 only, :104)
inline void safe_fusion3(Vec &x1, Mat &C1, const Vec &x2, const Mat &C2) {
    const Mat I1 = inverse_fixed(C1);
    Mat I2 = inverse_fixed(C2);
    Mat U1, U2;
    Vec s1, s2;
    jacobi_svd(I1, U1, s1);
    Mat sqrtD1(3, 3);
    for (int i = 0; i < 3; ++i) sqrtD1(i, i) = std::sqrt(s1[i]);
    const Mat isqrtD1 = inverse_fixed(sqrtD1);
    I2 = matmul(matmul(matmul(matmul(isqrtD1, U1.transpose()), I2), U1), isqrtD1);  // :93
    jacobi_svd(I2, U2, s2);
    const Mat T = matmul(matmul(U2.transpose(), sqrtD1), U1);  // :106, as written
    const Vec d1 = matvec(T, x1), d2 = matvec(T, x2);
    Vec result(3);
    Mat D3(3, 3);
    for (int i = 0; i < 3; ++i) {
        if (s2[i] < 1.0) { result[i] = d1[i]; D3(i, i) = 1.0; }
        else { result[i] = d2[i]; D3(i, i) = s2[i]; }
    }
    const Mat Ti = inverse_fixed(T);
    x1 = matvec(Ti, result);
    C1 = matmul(matmul(Ti, inverse_fixed(D3)), Ti.transpose());
}

// ==========================================================================================
// f4: dead reckoning with uncertainty.
// ==========================================================================================
// Eigen::Quaternion::toRotationMatrix
inline Mat quat_to_rot(const double q[4]) {
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    Mat R(3, 3);
    R(0, 0) = 1 - (tyy + tzz); R(0, 1) = txy - twz; R(0, 2) = txz + twy;
    R(1, 0) = txy + twz; R(1, 1) = 1 - (txx + tzz); R(1, 2) = tyz - twx;
    R(2, 0) = txz - twy; R(2, 1) = tyz + twx; R(2, 2) = 1 - (txx + tyy);
    return R;
}
// Eigen::Quaternion(Matrix3) -- quaternionbase_assign_impl<Other,3,3>
inline void rot_to_quat(const Mat &m, double q[4]) {
    double t = m(0, 0) + m(1, 1) + m(2, 2);
    if (t > 0.0) {
        t = std::sqrt(t + 1.0);
        q[0] = 0.5 * t;
        t = 0.5 / t;
        q[1] = (m(2, 1) - m(1, 2)) * t;
        q[2] = (m(0, 2) - m(2, 0)) * t;
        q[3] = (m(1, 0) - m(0, 1)) * t;
    } else {
        int i = 0;
        if (m(1, 1) > m(0, 0)) i = 1;
        if (m(2, 2) > m(i, i)) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
        q[1 + i] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (m(k, j) - m(j, k)) * t;
        q[1 + j] = (m(j, i) + m(i, j)) * t;
        q[1 + k] = (m(k, i) + m(i, k)) * t;
    }
}
// q_to_r (Transform.cpp:46-50) through Eigen::AngleAxis(Quaternion)
inline void q_to_r(const double q[4], double r[3]) {
    double n = std::sqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    if (n < DBL_EPSILON) n = std::sqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);  // stableNorm(): same value here
    if (n != 0.0) {
        const double angle = 2.0 * std::atan2(n, std::fabs(q[0]));
        if (q[0] < 0.0) n = -n;
        for (int i = 0; i < 3; ++i) r[i] = q[1 + i] / n * angle;
    } else {
        r[0] = 0.0; r[1] = 0.0; r[2] = 0.0;  // angle 0 about (1,0,0)
    }
}
inline Mat skew(const double r[3]) {
    Mat S(3, 3);
    S(0, 1) = -r[2]; S(0, 2) = r[1]; S(1, 0) = r[2]; S(1, 2) = -r[0]; S(2, 0) = -r[1]; S(2, 1) = r[0];
    return S;
}
inline Mat outer3(const double a[3], const double b[3]) {
    Mat M(3, 3);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M(i, j) = a[i] * b[j];
    return M;
}
inline Mat scaled(const Mat &A, double s) {
    Mat B = A;
    for (double &v : B.a) v *= s;
    return B;
}
// Transform.cpp:65-77
inline Mat dq_by_dr(const double q[4]) {
    double r[3];
    q_to_r(q, r);
    const double theta = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    const double kappa = 0.5 - theta * theta / 48.0;
    const double lambda = 1.0 / 24.0 * (1.0 - theta * theta / 40.0);
    Mat res(4, 3);
    for (int j = 0; j < 3; ++j) res(0, j) = -q[1 + j] / 2.0;
    const Mat rr = outer3(r, r);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) res(1 + i, j) = kappa * (i == j ? 1.0 : 0.0) - lambda * rr(i, j);
    return res;
}
// Transform.cpp:79-90
inline Mat dr_by_dq(const double q[4]) {
    const double mu = std::sqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const double sg = q[0] > 0 ? 1.0 : -1.0;
    const double tau = 2.0 * sg * (1.0 + mu * mu / 6.0);
    const double nu = -2.0 * sg * (2.0 / 3.0 + mu * mu / 5.0);
    Mat res(3, 4);
    const Mat vv = outer3(q + 1, q + 1);
    for (int i = 0; i < 3; ++i) {
        res(i, 0) = -2 * q[1 + i];
        for (int j = 0; j < 3; ++j) res(i, 1 + j) = tau * (i == j ? 1.0 : 0.0) + nu * vv(i, j);
    }
    return res;
}
// Transform.cpp:92-106: sgn = +1 for dq2q1_by_dq1(q2), -1 for dq2q1_by_dq2(q1)
inline Mat dq2q1_by(const double q[4], double sgn) {
    Mat res(4, 4);
    const Mat S = skew(q + 1);
    for (int j = 0; j < 3; ++j) { res(0, 1 + j) = -q[1 + j]; res(1 + j, 0) = q[1 + j]; }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) res(1 + i, 1 + j) = sgn * S(i, j);
    Mat out(4, 4);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) out(i, j) = (i == j ? 1.0 : 0.0) * q[0] + res(i, j);
    return out;
}
// Transform.cpp:108-122
inline Mat dr2r1_by_r1(const double q[4], const double q1[4], const double q2[4]) {
    return matmul(matmul(dr_by_dq(q), dq2q1_by(q2, 1.0)), dq_by_dr(q1));
}
inline Mat dr2r1_by_r2(const double q[4], const double q1[4], const double q2[4]) {
    return matmul(matmul(dr_by_dq(q), dq2q1_by(q1, -1.0)), dq_by_dr(q2));
}
// Transform.cpp:124-138
inline Mat drx_by_dr(const double q[4], const double x[3]) {
    double r[3];
    q_to_r(q, r);
    const double theta = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    const double alpha = 1.0 - theta * theta / 6.0;
    const double beta = 0.5 - theta * theta / 24.0;
    const double gamma = 1.0 / 3.0 - theta * theta / 30.0;
    const double delta = -1.0 / 12.0 + theta * theta / 180.0;
    const Mat rr = outer3(r, r), Sx = skew(x), Sr = skew(r), I = Mat::identity(3);
    const Mat A = add(sub(scaled(rr, gamma), scaled(Sr, beta)), scaled(I, alpha));
    const Mat B = add(scaled(rr, delta), scaled(I, 2.0 * beta));
    return sub(matmul(scaled(Sx, -1.0), A), matmul(matmul(Sr, Sx), B));
}

// pose = pos(3) quat(w,x,y,z); covariance 6x6 over [r t] (rotation first), Transform.hpp:48-60.
// result = t2 * t1 with both uncertain, TransformWithUncertainty::operator* Transform.cpp:215-254.
inline void transform_compose(const double *pose2, const Mat &cov2, const double *pose1, const Mat &cov1,
                              double *pose_out, Mat &cov_out) {
    const Mat R1 = quat_to_rot(pose1 + 3), R2 = quat_to_rot(pose2 + 3);
    double q1[4], q2[4], q[4];
    rot_to_quat(R1, q1);
    rot_to_quat(R2, q2);
    quat_mul(q2, q1, q);
    Mat J1(6, 6), J2(6, 6);
    J1.set_block(0, 0, dr2r1_by_r1(q, q1, q2));
    J1.set_block(3, 3, R2);
    J2.set_block(0, 0, dr2r1_by_r2(q, q1, q2));
    J2.set_block(3, 0, drx_by_dr(q2, pose1));
    J2.set_block(3, 3, Mat::identity(3));
    cov_out = add(matmul(matmul(J1, cov1), J1.transpose()), matmul(matmul(J2, cov2), J2.transpose()));
    // t2.getTransform() * t1.getTransform(): linear = R2 R1, translation = R2 p1 + p2
    const Mat R = matmul(R2, R1);
    for (int i = 0; i < 3; ++i) pose_out[i] = R2(i, 0) * pose1[0] + R2(i, 1) * pose1[1] + R2(i, 2) * pose1[2] + pose2[i];
    rot_to_quat(R, pose_out + 3);
}

// DeadReckon::updateAttitude DeadReckon.hpp:246-286 (w0 = angularVelocities[0], w1 = [1])
inline void dr_update_attitude(double dt, const double w0[3], const double w1[3], double dq[4]) {
    auto omega = [](const double w[3]) {
        Mat O(4, 4);
        O(0, 1) = -w[0]; O(0, 2) = -w[1]; O(0, 3) = -w[2];
        O(1, 0) = w[0]; O(1, 2) = w[2]; O(1, 3) = -w[1];
        O(2, 0) = w[1]; O(2, 1) = -w[2]; O(2, 3) = w[0];
        O(3, 0) = w[2]; O(3, 1) = w[1]; O(3, 2) = -w[0];
        return O;
    };
    const Mat O4 = omega(w0), Oo = omega(w1), I = Mat::identity(4);
    const double n2 = w0[0] * w0[0] + w0[1] * w0[1] + w0[2] * w0[2];
    const double dt2 = std::pow(dt, 2), dt3 = std::pow(dt, 3);
    Mat M = add(I, scaled(scaled(O4, 0.75), dt));
    M = sub(M, scaled(scaled(Oo, 0.25), dt));
    M = sub(M, scaled(I, (1.0 / 6.0) * n2 * dt2));
    M = sub(M, scaled(matmul(scaled(O4, 1.0 / 24.0), Oo), dt2));
    M = sub(M, scaled(scaled(O4, (1.0 / 48.0) * n2), dt3));
    const double quat[4] = {M(0, 0), M(1, 0), M(2, 0), M(3, 0)};  // M * (1,0,0,0)
    const double n = std::sqrt(quat[0] * quat[0] + quat[1] * quat[1] + quat[2] * quat[2] + quat[3] * quat[3]);
    for (int i = 0; i < 4; ++i) dq[i] = quat[i] / n;
}

// DeadReckon::updatePose(delta_t, cartesianVelocities[2], cartesianVelCov, prevPose, postPose)
// DeadReckon.hpp:30-79.  vel0/vel1: (linear 3, angular 3); velcov 6x6; returns deltaPose too.
inline void dr_update_pose(double dt, const double vel0[6], const double vel1[6], const Mat &velcov, const double *prev_pose,
                           const Mat &prev_cov, double *post_pose, Mat &post_cov, double *delta_pose, Mat &delta_cov) {
    double dq[4];
    dr_update_attitude(dt, vel0 + 3, vel1 + 3, dq);
    // deltaTrans = deltaq (Affine3d from a quaternion: linear = toRotationMatrix) ; the composition converts back
    for (int i = 0; i < 3; ++i) delta_pose[i] = (dt / 2.0) * (vel0[i] + vel1[i]);
    for (int i = 0; i < 4; ++i) delta_pose[3 + i] = dq[i];
    Mat dcov(6, 6);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            dcov(i, j) = velcov(3 + i, 3 + j) * dt * dt;
            dcov(3 + i, 3 + j) = velcov(i, j) * dt * dt;
        }
    Mat L;
    llt_lower(dcov, L);  // :50-52: cov = L L^T
    delta_cov = matmul(L, L.transpose());
    transform_compose(prev_pose, prev_cov, delta_pose, delta_cov, post_pose, post_cov);  // postPose = prevPose * deltaPose
}

}  // namespace slo
