// slb_math.cuh -- device-side manifold algebra and small dense FP64 helpers (sm_100a).
//
// Same maths as the reference's MTK SO3/vect blocks (State.hpp:186-200 via MTK::SO3::boxplus /
// boxminus / exp / log) and the Eigen routines its filters call; written for registers:
// every loop is fully unrolled over compile-time layouts so arrays never leave the register file.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace slbd {

#define SLB_DEV __device__ __forceinline__
#define SLB_HD __host__ __device__ __forceinline__

// ---- compile-time manifold layout: NB blocks of 3 DOF, bit b of MASK set = block b is SO3 ------
template <int NB_, unsigned MASK_>
struct Layout {
    static constexpr int NB = NB_;
    static constexpr unsigned MASK = MASK_;
    static constexpr int N = 3 * NB_;
    static constexpr int NP = N * (N + 1) / 2;  // packed lower triangle
    SLB_HD static constexpr bool so3(int b) { return (MASK_ >> b) & 1u; }
    SLB_HD static constexpr int qoff(int b) {
        int o = 0;
        for (int i = 0; i < b; ++i) o += ((MASK_ >> i) & 1u) ? 4 : 3;
        return o;
    }
    static constexpr int QD = qoff(NB_);
};
typedef Layout<2, 0x2> LayPose6;    // vect3 pos, SO3 orient            (SensorState)
typedef Layout<3, 0x2> LayMtk9;     // pos, orient, vel                 (mtk_state)
typedef Layout<4, 0x2> LayState12;  // pos, orient, velo, angvelo       (State)

SLB_HD constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }  // i >= j

// ---- SO3 -----------------------------------------------------------------------------------------
// MTK cos_sinc_sqrt: cos(sqrt(x)), sin(sqrt(x))/sqrt(x); 3-term Taylor pair below eps^(1/4).
SLB_DEV void cos_sinc_sqrt(double x, double &c, double &s) {
    const double taylor_n = 1.220703125e-4;  // sqrt(sqrt(DBL_EPSILON)) = 2^-13
    if (x >= taylor_n) {
        const double sx = sqrt(x);
        double sn, cs;
        sincos(sx, &sn, &cs);
        c = cs;
        s = sn / sx;
    } else {
        double cosi = 1.0, sinc = 1.0;
        double term = -0.5 * x;
        cosi += term; term *= (1.0 / 3.0); sinc += term; term *= -(1.0 / 4.0) * x;
        cosi += term; term *= (1.0 / 5.0); sinc += term; term *= -(1.0 / 6.0) * x;
        cosi += term; term *= (1.0 / 7.0); sinc += term;
        c = cosi;
        s = sinc;
    }
}
// exp(v, scale): (w,x,y,z) of a rotation by scale*|v| about v/|v|
SLB_DEV void so3_exp(const double v[3], double scale, double q[4]) {
    const double h = 0.5 * scale;
    const double n2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    double c, s;
    cos_sinc_sqrt(h * h * n2, c, s);
    const double m = s * h;
    q[0] = c; q[1] = m * v[0]; q[2] = m * v[1]; q[3] = m * v[2];
}
// log(q) = 2 atan(|qv|/qw)/|qv| qv   (atan: +-q identified; |qv| clamped at 1e-11)
SLB_DEV void so3_log(const double q[4], double v[3]) {
    double nv = sqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    nv = nv < 1e-11 ? 1e-11 : nv;
    const double s = 2.0 / nv * atan(nv / q[0]);
    v[0] = s * q[1]; v[1] = s * q[2]; v[2] = s * q[3];
}
SLB_DEV void quat_mul(const double a[4], const double b[4], double o[4]) {
    const double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    const double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    const double y = a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3];
    const double z = a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1];
    o[0] = w; o[1] = x; o[2] = y; o[3] = z;
}
// conj(a) * b
SLB_DEV void quat_cmul(const double a[4], const double b[4], double o[4]) {
    const double w = a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3];
    const double x = a[0] * b[1] - a[1] * b[0] - a[2] * b[3] + a[3] * b[2];
    const double y = a[0] * b[2] - a[2] * b[0] - a[3] * b[1] + a[1] * b[3];
    const double z = a[0] * b[3] - a[3] * b[0] - a[1] * b[2] + a[2] * b[1];
    o[0] = w; o[1] = x; o[2] = y; o[3] = z;
}
// q * v (Eigen _transformVector form)
SLB_DEV void quat_rotate(const double q[4], const double v[3], double o[3]) {
    const double tx = 2.0 * (q[2] * v[2] - q[3] * v[1]);
    const double ty = 2.0 * (q[3] * v[0] - q[1] * v[2]);
    const double tz = 2.0 * (q[1] * v[1] - q[2] * v[0]);
    o[0] = v[0] + q[0] * tx + (q[2] * tz - q[3] * ty);
    o[1] = v[1] + q[0] * ty + (q[3] * tx - q[1] * tz);
    o[2] = v[2] + q[0] * tz + (q[1] * ty - q[2] * tx);
}
// conj(q) * v
SLB_DEV void quat_rotate_inv(const double q[4], const double v[3], double o[3]) {
    const double qc[4] = {q[0], -q[1], -q[2], -q[3]};
    quat_rotate(qc, v, o);
}

// ---- compound manifold ops on register arrays -------------------------------------------------------
// y = x [+] (sign * d)
template <class L>
SLB_DEV void boxplus(const double *x, const double *d, double sign, double *y) {
#pragma unroll
    for (int b = 0; b < L::NB; ++b) {
        const int o = L::qoff(b);
        if (L::so3(b)) {
            const double v[3] = {sign * d[3 * b], sign * d[3 * b + 1], sign * d[3 * b + 2]};
            double e[4];
            so3_exp(v, 1.0, e);
            quat_mul(x + o, e, y + o);
        } else {
#pragma unroll
            for (int i = 0; i < 3; ++i) y[o + i] = x[o + i] + sign * d[3 * b + i];
        }
    }
}
// d = a [-] b
template <class L>
SLB_DEV void boxminus(const double *a, const double *b_, double *d) {
#pragma unroll
    for (int b = 0; b < L::NB; ++b) {
        const int o = L::qoff(b);
        if (L::so3(b)) {
            double r[4];
            quat_cmul(b_ + o, a + o, r);
            so3_log(r, d + 3 * b);
        } else {
#pragma unroll
            for (int i = 0; i < 3; ++i) d[3 * b + i] = a[o + i] - b_[o + i];
        }
    }
}

// ---- in-register packed Cholesky (lower, row-major packed).  Returns false on pivot <= 0. ----------
template <int N>
SLB_DEV bool chol_packed(double *A) {
    bool ok = true;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        double x = A[tri(k, k)];
#pragma unroll
        for (int p = 0; p < k; ++p) x -= A[tri(k, p)] * A[tri(k, p)];
        ok = ok && (x > 0.0);
        x = sqrt(x);
        A[tri(k, k)] = x;
        const double inv = 1.0 / x;
#pragma unroll
        for (int i = k + 1; i < N; ++i) {
            double s = A[tri(i, k)];
#pragma unroll
            for (int p = 0; p < k; ++p) s -= A[tri(i, p)] * A[tri(k, p)];
            A[tri(i, k)] = s * inv;
        }
    }
    return ok;
}

// symmetric 3x3 inverse from packed lower S (s00 s10 s11 s20 s21 s22) -> packed lower
SLB_DEV void sym3_inverse(const double *S, double *Si) {
    const double a = S[0], b = S[1], c = S[2], d = S[3], e = S[4], f = S[5];
    const double c00 = c * f - e * e, c10 = d * e - b * f, c20 = b * e - c * d;
    const double det = a * c00 + b * c10 + d * c20;
    const double id = 1.0 / det;
    Si[0] = c00 * id;
    Si[1] = c10 * id;
    Si[2] = (a * f - d * d) * id;
    Si[3] = c20 * id;
    Si[4] = (b * d - a * e) * id;
    Si[5] = (a * c - b * b) * id;
}

SLB_DEV bool chi2_accept(double m2, int dof) {
    // 5% table, Usckf.hpp:794-855; dof 0 = accept_any_mahalanobis_distance
    if (dof == 0) return true;
    const double th[10] = {0, 3.84, 5.99, 7.81, 9.49, 11.07, 12.59, 14.07, 15.51, 16.92};
    if (dof < 1 || dof > 9) return false;
    return m2 < th[dof];
}

}  // namespace slbd
