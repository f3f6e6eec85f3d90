// oracle/slo_core.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement ("oracle") of the numerical primitives underneath the reference's
// sigma-point filters: MTK manifold algebra (vect<3>, SO3 exp/log/boxplus/boxminus) and the
// Eigen dense routines the filters call (LLT, PartialPivLU inverse, cofactor inverse).
//
// PARITY UNPINNED for filter outputs: the reference (jhidalgocarrio/slam-localization)
// cannot be compiled here (Eigen, Boost, MTK, ukfom, Rock base-types are absent) and its tests
// assert no filter output.  What *is* pinned by the reference's own tests
// (test/MsckfUnitTest.cpp:61-71,104-113) is the manifold algebra in this file; see
// tests/test_oracle_manifold.py.
//
// Third-party code restated here (source not under /root/reference, no version pinned by the
// reference: manifest.xml:15-16, src/CMakeLists.txt:26):
//   * MTK (Rock package slam/mtk): SO3::exp/log/boxplus/boxminus, vect::boxplus/boxminus,
//     cos_sinc_sqrt -- call sites State.hpp:82,128,179,186-200,231,282,327.
//   * Eigen 3: LLT (Usckf.hpp:537,577; Msckf.hpp:412,447), .inverse() (Usckf.hpp:154,286;
//     Msckf.hpp:138,257,736; DataModel.hpp:54-55).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// build, link or call anything in this directory.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace slo {

// ------------------------------------------------------------------------------------------
// Dense row-major matrix of doubles (stand-in for Eigen::Matrix<double,Dynamic,Dynamic>).
// ------------------------------------------------------------------------------------------
struct Mat {
    int r = 0, c = 0;
    std::vector<double> a;
    Mat() {}
    Mat(int rows, int cols) : r(rows), c(cols), a((size_t)rows * cols, 0.0) {}
    double &operator()(int i, int j) { return a[(size_t)i * c + j]; }
    double operator()(int i, int j) const { return a[(size_t)i * c + j]; }
    static Mat identity(int n) {
        Mat m(n, n);
        for (int i = 0; i < n; ++i) m(i, i) = 1.0;
        return m;
    }
    Mat block(int i0, int j0, int rows, int cols) const {
        Mat m(rows, cols);
        for (int i = 0; i < rows; ++i)
            for (int j = 0; j < cols; ++j) m(i, j) = (*this)(i0 + i, j0 + j);
        return m;
    }
    void set_block(int i0, int j0, const Mat &b) {
        for (int i = 0; i < b.r; ++i)
            for (int j = 0; j < b.c; ++j) (*this)(i0 + i, j0 + j) = b(i, j);
    }
    Mat transpose() const {
        Mat m(c, r);
        for (int i = 0; i < r; ++i)
            for (int j = 0; j < c; ++j) m(j, i) = (*this)(i, j);
        return m;
    }
};
typedef std::vector<double> Vec;

inline Mat matmul(const Mat &A, const Mat &B) {
    Mat C(A.r, B.c);
    for (int i = 0; i < A.r; ++i)
        for (int k = 0; k < A.c; ++k) {
            const double aik = A(i, k);
            for (int j = 0; j < B.c; ++j) C(i, j) += aik * B(k, j);
        }
    return C;
}
inline Vec matvec(const Mat &A, const Vec &x) {
    Vec y(A.r, 0.0);
    for (int i = 0; i < A.r; ++i) {
        double s = 0;
        for (int j = 0; j < A.c; ++j) s += A(i, j) * x[j];
        y[i] = s;
    }
    return y;
}
inline Mat add(const Mat &A, const Mat &B) {
    Mat C(A.r, A.c);
    for (size_t i = 0; i < C.a.size(); ++i) C.a[i] = A.a[i] + B.a[i];
    return C;
}
inline Mat sub(const Mat &A, const Mat &B) {
    Mat C(A.r, A.c);
    for (size_t i = 0; i < C.a.size(); ++i) C.a[i] = A.a[i] - B.a[i];
    return C;
}
inline double norm2(const Vec &v) {
    double s = 0;
    for (double x : v) s += x * x;
    return std::sqrt(s);
}

// ------------------------------------------------------------------------------------------
// Eigen::LLT restated.  Reads only the lower triangle (Q8).  Unblocked left-looking column
// algorithm for n < 32, right-looking blocked (block = clamp((n/8)/16*16, 8, 128)) otherwise,
// which is the schedule Eigen 3's llt_inplace<Lower> uses.  Returns -1 on success or the index
// of the first non-positive pivot; like Eigen, the partially factored matrix is left in place
// and the caller (Usckf.hpp:537-538, Msckf.hpp:412-413) never looks at the status.
// The result is a full matrix whose strict upper triangle is zero (matrixL()).
// ------------------------------------------------------------------------------------------
inline int llt_unblocked(Mat &A, int off, int n) {
    for (int k = 0; k < n; ++k) {
        double x = A(off + k, off + k);
        for (int p = 0; p < k; ++p) x -= A(off + k, off + p) * A(off + k, off + p);
        if (!(x > 0.0)) return k;
        x = std::sqrt(x);
        A(off + k, off + k) = x;
        for (int i = k + 1; i < n; ++i) {
            double s = A(off + i, off + k);
            for (int p = 0; p < k; ++p) s -= A(off + i, off + p) * A(off + k, off + p);
            A(off + i, off + k) = s / x;
        }
    }
    return -1;
}
inline int llt_lower(const Mat &P, Mat &L) {
    const int n = P.r;
    L = P;
    int info = -1;
    if (n < 32) {
        info = llt_unblocked(L, 0, n);
    } else {
        int bs = n / 8;
        bs = (bs / 16) * 16;
        if (bs < 8) bs = 8;
        if (bs > 128) bs = 128;
        for (int k = 0; k < n && info < 0; k += bs) {
            const int b = (bs < n - k) ? bs : (n - k);
            const int rs = n - k - b;
            int ret = llt_unblocked(L, k, b);
            if (ret >= 0) { info = k + ret; break; }
            // A21 <- A21 * L11^-T
            for (int i = 0; i < rs; ++i)
                for (int j = 0; j < b; ++j) {
                    double s = L(k + b + i, k + j);
                    for (int p = 0; p < j; ++p) s -= L(k + b + i, k + p) * L(k + j, k + p);
                    L(k + b + i, k + j) = s / L(k + j, k + j);
                }
            // A22 <- A22 - A21 A21^T (lower part)
            for (int i = 0; i < rs; ++i)
                for (int j = 0; j <= i; ++j) {
                    double s = 0;
                    for (int p = 0; p < b; ++p) s += L(k + b + i, k + p) * L(k + b + j, k + p);
                    L(k + b + i, k + b + j) -= s;
                }
        }
    }
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j) L(i, j) = 0.0;
    return info;
}

// ------------------------------------------------------------------------------------------
// Eigen's general .inverse(): PartialPivLU, then solve against the identity (Q9).
// ------------------------------------------------------------------------------------------
inline Mat inverse_lu(const Mat &A) {
    const int n = A.r;
    Mat LU = A;
    std::vector<int> perm(n);
    for (int i = 0; i < n; ++i) perm[i] = i;
    for (int k = 0; k < n; ++k) {
        int piv = k;
        double best = std::fabs(LU(k, k));
        for (int i = k + 1; i < n; ++i)
            if (std::fabs(LU(i, k)) > best) { best = std::fabs(LU(i, k)); piv = i; }
        if (piv != k) {
            for (int j = 0; j < n; ++j) std::swap(LU(k, j), LU(piv, j));
            std::swap(perm[k], perm[piv]);
        }
        const double d = LU(k, k);
        for (int i = k + 1; i < n; ++i) LU(i, k) /= d;
        for (int i = k + 1; i < n; ++i) {
            const double lik = LU(i, k);
            for (int j = k + 1; j < n; ++j) LU(i, j) -= lik * LU(k, j);
        }
    }
    Mat X(n, n);
    for (int col = 0; col < n; ++col) {
        Vec y(n);
        for (int i = 0; i < n; ++i) {  // forward: L y = P e_col
            double s = (perm[i] == col) ? 1.0 : 0.0;
            for (int p = 0; p < i; ++p) s -= LU(i, p) * y[p];
            y[i] = s;
        }
        for (int i = n - 1; i >= 0; --i) {  // backward: U x = y
            double s = y[i];
            for (int p = i + 1; p < n; ++p) s -= LU(i, p) * X(p, col);
            X(i, col) = s / LU(i, i);
        }
    }
    return X;
}

// The same PartialPivLU inverse with the multiply-subtract steps of the elimination and of the two triangular solves
// fused (std::fma), which is what Eigen's kernels do when the reference is built for an FMA-capable CPU (pmadd).
// Used by DataModel::fusion for D > 3 so that the CUDA kernel -- FP64-issue-bound on its three 6x6 inverses -- can run
// the identical sequence with DFMA and stay bit-exact at half the floating-point instructions.
inline Mat inverse_lu_fma(const Mat &A) {
    const int n = A.r;
    Mat LU = A;
    std::vector<int> perm(n);
    for (int i = 0; i < n; ++i) perm[i] = i;
    for (int k = 0; k < n; ++k) {
        int piv = k;
        double best = std::fabs(LU(k, k));
        for (int i = k + 1; i < n; ++i)
            if (std::fabs(LU(i, k)) > best) { best = std::fabs(LU(i, k)); piv = i; }
        if (piv != k) {
            for (int j = 0; j < n; ++j) std::swap(LU(k, j), LU(piv, j));
            std::swap(perm[k], perm[piv]);
        }
        const double d = LU(k, k);
        for (int i = k + 1; i < n; ++i) LU(i, k) /= d;
        for (int i = k + 1; i < n; ++i) {
            const double lik = LU(i, k);
            for (int j = k + 1; j < n; ++j) LU(i, j) = std::fma(-lik, LU(k, j), LU(i, j));
        }
    }
    Mat X(n, n);
    for (int col = 0; col < n; ++col) {
        Vec y(n);
        for (int i = 0; i < n; ++i) {
            double s = (perm[i] == col) ? 1.0 : 0.0;
            for (int p = 0; p < i; ++p) s = std::fma(-LU(i, p), y[p], s);
            y[i] = s;
        }
        for (int i = n - 1; i >= 0; --i) {
            double s = y[i];
            for (int p = i + 1; p < n; ++p) s = std::fma(-LU(i, p), X(p, col), s);
            X(i, col) = s / LU(i, i);
        }
    }
    return X;
}

// Eigen's fixed-size inverse for 3x3: cofactors / determinant (used by DataModel<double,3>).
inline Mat inverse_3x3_cofactor(const Mat &A) {
    Mat C(3, 3);
    auto cof = [&](int i, int j) {
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
        return A(i1, j1) * A(i2, j2) - A(i1, j2) * A(i2, j1);
    };
    const double c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
    const double det = c00 * A(0, 0) + c10 * A(1, 0) + c20 * A(2, 0);
    const double invdet = 1.0 / det;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C(j, i) = cof(i, j) * invdet;
    return C;
}

// What Eigen's Matrix<double,D,D>::inverse() does for a *fixed* dimension D.
inline Mat inverse_fixed(const Mat &A) {
    if (A.r == 1) { Mat m(1, 1); m(0, 0) = 1.0 / A(0, 0); return m; }
    if (A.r == 2) {
        Mat m(2, 2);
        const double invdet = 1.0 / (A(0, 0) * A(1, 1) - A(1, 0) * A(0, 1));
        m(0, 0) = A(1, 1) * invdet; m(1, 0) = -A(1, 0) * invdet;
        m(0, 1) = -A(0, 1) * invdet; m(1, 1) = A(0, 0) * invdet;
        return m;
    }
    if (A.r == 3) return inverse_3x3_cofactor(A);
    return inverse_lu_fma(A);  // D == 4 uses a cofactor kernel in Eigen; not on the hot path here
}

// ------------------------------------------------------------------------------------------
// MTK SO3 / vect algebra.  Quaternions are stored (w, x, y, z).
// ------------------------------------------------------------------------------------------
// MTK cos_sinc_sqrt(x): returns cos(sqrt(x)) and sin(sqrt(x))/sqrt(x), Taylor below eps^(1/4).
inline void cos_sinc_sqrt(double x, double &c, double &s) {
    static const double taylor_0 = std::numeric_limits<double>::epsilon();
    static const double taylor_2 = std::sqrt(taylor_0);
    static const double taylor_n = std::sqrt(taylor_2);
    if (x >= taylor_n) {
        const double sx = std::sqrt(x);
        c = std::cos(sx);
        s = std::sin(sx) / sx;
        return;
    }
    static const double inv[] = {1 / 3., 1 / 4., 1 / 5., 1 / 6., 1 / 7., 1 / 8., 1 / 9.};
    double cosi = 1., sinc = 1.;
    double term = -1 / 2. * x;
    for (int i = 0; i < 3; ++i) {
        cosi += term;
        term *= inv[2 * i];
        sinc += term;
        term *= -inv[2 * i + 1] * x;
    }
    c = cosi;
    s = sinc;
}

// SO3::exp(v, scale): unit quaternion for a rotation of scale*|v| rad about v/|v|.
// Convention pinned by test/MsckfUnitTest.cpp:104-110 (State::set uses exp(v,1), State.hpp:179,
// and must equal the boxplus increment).
inline void so3_exp(const double v[3], double scale, double q[4]) {
    const double h = scale * 0.5;
    const double n2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    double c, s;
    cos_sinc_sqrt(h * h * n2, c, s);
    const double mult = s * h;
    q[0] = c;
    q[1] = mult * v[0];
    q[2] = mult * v[1];
    q[3] = mult * v[2];
}

// SO3::log(q): 2*atan(|qv|/qw)/|qv| * qv  (atan, not atan2: +-q identified; |qv| clamped 1e-11).
inline void so3_log(const double q[4], double v[3]) {
    double nv = std::sqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    if (nv < 1e-11) nv = 1e-11;
    const double s = 2.0 / nv * std::atan(nv / q[0]);
    v[0] = s * q[1];
    v[1] = s * q[2];
    v[2] = s * q[3];
}

inline void quat_mul(const double a[4], const double b[4], double o[4]) {
    const double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    const double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    const double y = a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3];
    const double z = a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1];
    o[0] = w; o[1] = x; o[2] = y; o[3] = z;
}
inline void quat_conj(const double a[4], double o[4]) {
    o[0] = a[0]; o[1] = -a[1]; o[2] = -a[2]; o[3] = -a[3];
}
// Eigen's Quaternion * Vector3 (QuaternionBase::_transformVector): v + w*t + qv x t, t = 2 qv x v.
inline void quat_rotate(const double q[4], const double v[3], double o[3]) {
    const double tx = 2.0 * (q[2] * v[2] - q[3] * v[1]);
    const double ty = 2.0 * (q[3] * v[0] - q[1] * v[2]);
    const double tz = 2.0 * (q[1] * v[1] - q[2] * v[0]);
    o[0] = v[0] + q[0] * tx + (q[2] * tz - q[3] * ty);
    o[1] = v[1] + q[0] * ty + (q[3] * tx - q[1] * tz);
    o[2] = v[2] + q[0] * tz + (q[1] * ty - q[2] * tx);
}
// SO3::boxplus(v, scale): q <- q * exp(v, scale)
inline void so3_boxplus(double q[4], const double v[3], double scale = 1.0) {
    double d[4], o[4];
    so3_exp(v, scale, d);
    quat_mul(q, d, o);
    std::memcpy(q, o, sizeof(o));
}
// SO3::boxminus: this [-] other = log(other^-1 * this)
inline void so3_boxminus(const double q[4], const double other[4], double v[3]) {
    double oc[4], d[4];
    quat_conj(other, oc);
    quat_mul(oc, q, d);
    so3_log(d, v);
}

// ------------------------------------------------------------------------------------------
// Compound manifold layout: a sequence of 3-DOF blocks (vect<3> or SO3) followed by `nfeat`
// plain scalars.  Covers State (V S V V, State.hpp:141-149), SensorState / ReducedState
// (V S, State.hpp:246-252,44-52), mtk_state (V S V, test/UKFoMUnitTest.cpp:31-35),
// MultiState (State + k SensorState, State.hpp:341-376) and AugmentedState (3 State + two
// feature vectors, State.hpp:536-545).  A point is stored as a flat "q-vector": 3 doubles per
// vect block, 4 (w,x,y,z) per SO3 block, then the scalars.
// ------------------------------------------------------------------------------------------
struct Layout {
    std::vector<uint8_t> so3;  // per block: 1 = SO3, 0 = vect<3>
    int nfeat = 0;
    int dof() const { return 3 * (int)so3.size() + nfeat; }
    int qdim() const {
        int q = nfeat;
        for (uint8_t s : so3) q += s ? 4 : 3;
        return q;
    }
    static Layout blocks(std::initializer_list<int> b, int nfeat = 0) {
        Layout l;
        for (int x : b) l.so3.push_back((uint8_t)x);
        l.nfeat = nfeat;
        return l;
    }
    static Layout pose6() { return blocks({0, 1}); }
    static Layout mtk9() { return blocks({0, 1, 0}); }
    static Layout state12() { return blocks({0, 1, 0, 0}); }
    static Layout multi(int k) {
        Layout l = state12();
        for (int i = 0; i < k; ++i) { l.so3.push_back(0); l.so3.push_back(1); }
        return l;
    }
    static Layout augmented(int nk, int nl) {
        Layout l;
        for (int i = 0; i < 3; ++i) { l.so3.push_back(0); l.so3.push_back(1); l.so3.push_back(0); l.so3.push_back(0); }
        l.nfeat = nk + nl;
        return l;
    }
    Vec identity() const {
        Vec x(qdim(), 0.0);
        int o = 0;
        for (uint8_t s : so3) { if (s) { x[o] = 1.0; o += 4; } else o += 3; }
        return x;
    }
};

// x [+] d  (State::boxplus State.hpp:186-192, SensorState :286-290, MultiState :418-434)
inline Vec boxplus(const Layout &l, const Vec &x, const Vec &d, double scale = 1.0) {
    Vec y = x;
    int o = 0, k = 0;
    for (uint8_t s : l.so3) {
        if (s) { so3_boxplus(&y[o], &d[k], scale); o += 4; }
        else { for (int i = 0; i < 3; ++i) y[o + i] += scale * d[k + i]; o += 3; }
        k += 3;
    }
    for (int i = 0; i < l.nfeat; ++i) y[o + i] += scale * d[k + i];
    return y;
}
// a [-] b  (State::boxminus State.hpp:194-200, MultiState :460-481)
inline Vec boxminus(const Layout &l, const Vec &a, const Vec &b) {
    Vec d(l.dof());
    int o = 0, k = 0;
    for (uint8_t s : l.so3) {
        if (s) { so3_boxminus(&a[o], &b[o], &d[k]); o += 4; }
        else { for (int i = 0; i < 3; ++i) d[k + i] = a[o + i] - b[o + i]; o += 3; }
        k += 3;
    }
    for (int i = 0; i < l.nfeat; ++i) d[k + i] = a[o + i] - b[o + i];
    return d;
}
// State::set(v, ANGLE_AXIS) (State.hpp:166-184): vect blocks copied, SO3 blocks = exp(v, 1).
inline Vec set_from_vector(const Layout &l, const Vec &v) {
    Vec x(l.qdim());
    int o = 0, k = 0;
    for (uint8_t s : l.so3) {
        if (s) { so3_exp(&v[k], 1.0, &x[o]); o += 4; }
        else { for (int i = 0; i < 3; ++i) x[o + i] = v[k + i]; o += 3; }
        k += 3;
    }
    for (int i = 0; i < l.nfeat; ++i) x[o + i] = v[k + i];
    return x;
}
// State::getVectorizedState(ANGLE_AXIS) (State.hpp:215-239): SO3 blocks = log(q).
inline Vec get_vectorized(const Layout &l, const Vec &x) {
    Vec v(l.dof());
    int o = 0, k = 0;
    for (uint8_t s : l.so3) {
        if (s) { so3_log(&x[o], &v[k]); o += 4; }
        else { for (int i = 0; i < 3; ++i) v[k + i] = x[o + i]; o += 3; }
        k += 3;
    }
    for (int i = 0; i < l.nfeat; ++i) v[k + i] = x[o + i];
    return v;
}

// AugmentedState "state (+) state" and "state (-) state -> state" (State.hpp:595-634 through
// MtkMultiStateWrap::operator+/- MtkWrap.hpp:277-310): the delta travels as a state, so it is
// passed through log before the boxplus and through exp after the boxminus (quirk Q10).
inline Vec aug_plus_state(const Layout &l, const Vec &x, const Vec &delta_state) {
    return boxplus(l, x, get_vectorized(l, delta_state));
}
inline Vec aug_minus_state(const Layout &l, const Vec &a, const Vec &b) {
    return set_from_vector(l, boxminus(l, a, b));
}

}  // namespace slo
