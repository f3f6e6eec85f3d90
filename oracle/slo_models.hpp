// oracle/slo_models.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the process / measurement models the reference ships in its tests (they
// are the only models on the hot path: src/filters/ProcessModels.hpp is empty), keyed by the
// same ids as the device catalogue in include/slb.h.  q-vector layouts:
//   STATE12: pos[0:3) quat[3:7) velo[7:10) angvelo[10:13)
//   MTK9   : pos[0:3) quat[3:7) vel[7:10)
//   POSE6  : pos[0:3) quat[3:7)
#pragma once
#include "../include/slb.h"
#include "slo_filters.hpp"

namespace slo {

// Eigen's Quaternion::toRotationMatrix applied to a vector (Affine3d * Vector3d path used by
// test/UsckfUnitTest.cpp:70-79).
inline void rotmat_apply(const double q[4], const double v[3], double o[3]) {
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    const double r00 = 1 - (tyy + tzz), r01 = txy - twz, r02 = txz + twy;
    const double r10 = txy + twz, r11 = 1 - (txx + tzz), r12 = tyz - twx;
    const double r20 = txz - twy, r21 = tyz + twx, r22 = 1 - (txx + tyy);
    o[0] = r00 * v[0] + r01 * v[1] + r02 * v[2];
    o[1] = r10 * v[0] + r11 * v[1] + r12 * v[2];
    o[2] = r20 * v[0] + r21 * v[1] + r22 * v[2];
}

// test/UsckfUnitTest.cpp:34-49
inline Vec pm_usckf_test(const Vec &s, const double *u, double dt) {
    Vec o(13);
    const double ax[3] = {u[3] * dt, u[4] * dt, u[5] * dt};
    double rot[4];
    so3_exp(ax, 1.0, rot);
    quat_mul(&s[3], rot, &o[3]);
    for (int i = 0; i < 3; ++i) {
        o[10 + i] = u[3 + i];
        o[7 + i] = u[i];
        o[i] = s[i] + s[7 + i] * dt;
    }
    return o;
}
// test/MsckfUnitTest.cpp:33-47   u = dp(3) dq(4) vel(3) angvel(3)
inline Vec pm_msckf_deltapose(const Vec &s, const double *u) {
    Vec o(13);
    quat_mul(&s[3], &u[3], &o[3]);
    double t[3];
    quat_rotate(&o[3], &u[0], t);
    for (int i = 0; i < 3; ++i) {
        o[i] = s[i] + t[i];
        o[7 + i] = u[7 + i];
        o[10 + i] = u[10 + i];
    }
    return o;
}
// test/UKFoMUnitTest.cpp:45-70   u = acc(3) gyro(3)
inline Vec pm_ukfom_imu(const Vec &s, const double *u, double dt, bool refbug) {
    Vec o(10);
    const double ax[3] = {u[3] * dt, u[4] * dt, u[5] * dt};
    double q[4] = {1, 0, 0, 0};
    if (!refbug) std::memcpy(q, &s[3], sizeof(q));
    so3_boxplus(q, ax, 1.0);  // :53
    std::memcpy(&o[3], q, sizeof(q));
    double ra[3];
    quat_rotate(&s[3], &u[0], ra);
    const double g[3] = {0, 0, 9.81};
    for (int i = 0; i < 3; ++i) {
        o[7 + i] = s[7 + i] + (ra[i] + g[i]) * dt;  // :61
        o[i] = s[i] + s[7 + i] * dt;                // :64
    }
    return o;
}
// builder-defined pose odometry   u = v_body(3) w(3)
inline Vec pm_pose6_odom(const Vec &s, const double *u, double dt) {
    Vec o(7);
    double rv[3];
    quat_rotate(&s[3], &u[0], rv);
    for (int i = 0; i < 3; ++i) o[i] = s[i] + rv[i] * dt;
    const double ax[3] = {u[3] * dt, u[4] * dt, u[5] * dt};
    std::memcpy(&o[3], &s[3], 4 * sizeof(double));
    so3_boxplus(&o[3], ax, 1.0);
    return o;
}
inline Model make_process_model(int pm, const double *u, double dt) {
    switch (pm) {
        case SLB_PM_UKFOM_IMU: return [=](const Vec &s) { return pm_ukfom_imu(s, u, dt, false); };
        case SLB_PM_UKFOM_IMU_REFBUG: return [=](const Vec &s) { return pm_ukfom_imu(s, u, dt, true); };
        case SLB_PM_POSE6_ODOM: return [=](const Vec &s) { return pm_pose6_odom(s, u, dt); };
        case SLB_PM_USCKF_TEST: return [=](const Vec &s) { return pm_usckf_test(s, u, dt); };
        case SLB_PM_MSCKF_DELTAPOSE: return [=](const Vec &s) { return pm_msckf_deltapose(s, u); };
    }
    return Model();
}

// test/UKFoMUnitTest.cpp:82-85
inline Vec mm_gps_pos(const Vec &s) { return Vec(s.begin(), s.begin() + 3); }

// test/UsckfUnitTest.cpp:62-86 on the augmented q-vector (statek at 0, statek_i at 26,
// featuresk at 39).
inline Vec mm_usckf_vo(const Vec &a, int nk) {
    const Layout ls = Layout::state12();
    const Vec sk(a.begin(), a.begin() + 13), si(a.begin() + 26, a.begin() + 39);
    const Vec delta = set_from_vector(ls, boxminus(ls, sk, si));  // delta_state = statek - statek_i (:70)
    Vec z(nk);
    for (int i = 0; i + 2 < nk; i += 3) {
        double c[3] = {a[39 + i], a[39 + i + 1], a[39 + i + 2]}, r[3];
        rotmat_apply(&delta[3], c, r);
        z[i] = r[0] + delta[0];
        z[i + 1] = r[1] + delta[1];
        z[i + 2] = r[2] + delta[2];
    }
    return z;
}

// Builder-defined MSCKF visual measurement: landmark f (world frame,
// shared by the batch) is seen from clone j = f % k with pose (p_j, q_j):
//   pc = R(q_j)^T (lm_f - p_j);  z_f = (pc.x/pc.z, pc.y/pc.z)
inline Vec mm_msckf_reproj(const Vec &s, int k, const double *lm, int nfeat) {
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
    }
    return z;
}

}  // namespace slo
