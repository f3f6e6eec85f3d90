// slb_models.cuh -- device model catalogue (ids in include/slb.h).
//
// The reference takes arbitrary host functors (Usckf.hpp:113-114,260-263); its only models on the
// hot path are the ones in test/*.cpp (src/filters/ProcessModels.hpp is empty).  Each is restated
// here as a small struct: prepare() does the per-instance, sigma-point-independent part once
// (e.g. exp(w dt)), apply() maps one sigma point held in registers.
#pragma once
#include "../../include/slb.h"
#include "slb_math.cuh"

namespace slbd {

template <int PM>
struct ProcessModel;

// test/UKFoMUnitTest.cpp:45-70.  MTK9 q-vector: pos[0:3) quat[3:7) vel[7:10).  u = acc gyro.
template <bool REFBUG>
struct PmUkfomImu {
    static constexpr int NU = 6;
    double rot[4], acc[3], dt;
    SLB_DEV void prepare(const double *u, double dt_) {
        dt = dt_;
        const double ax[3] = {u[3] * dt_, u[4] * dt_, u[5] * dt_};
        so3_exp(ax, 1.0, rot);  // boxplus increment of :53, identical for every sigma point
        acc[0] = u[0]; acc[1] = u[1]; acc[2] = u[2];
    }
    SLB_DEV void apply(const double *s, double *o) const {
        if (REFBUG) { o[3] = rot[0]; o[4] = rot[1]; o[5] = rot[2]; o[6] = rot[3]; }  // identity * rot
        else quat_mul(s + 3, rot, o + 3);
        double ra[3];
        quat_rotate(s + 3, acc, ra);
        o[7] = s[7] + (ra[0] + 0.0) * dt;
        o[8] = s[8] + (ra[1] + 0.0) * dt;
        o[9] = s[9] + (ra[2] + 9.81) * dt;
        o[0] = s[0] + s[7] * dt;
        o[1] = s[1] + s[8] * dt;
        o[2] = s[2] + s[9] * dt;
    }
};
template <> struct ProcessModel<SLB_PM_UKFOM_IMU> : PmUkfomImu<false> {};
template <> struct ProcessModel<SLB_PM_UKFOM_IMU_REFBUG> : PmUkfomImu<true> {};

// builder-defined pose odometry.  POSE6 q-vector: pos[0:3) quat[3:7).  u = v_body w.
template <>
struct ProcessModel<SLB_PM_POSE6_ODOM> {
    static constexpr int NU = 6;
    double rot[4], vdt[3];
    SLB_DEV void prepare(const double *u, double dt) {
        const double ax[3] = {u[3] * dt, u[4] * dt, u[5] * dt};
        so3_exp(ax, 1.0, rot);
        vdt[0] = u[0]; vdt[1] = u[1]; vdt[2] = u[2];
        dt_ = dt;
    }
    double dt_;
    SLB_DEV void apply(const double *s, double *o) const {
        double rv[3];
        quat_rotate(s + 3, vdt, rv);
        o[0] = s[0] + rv[0] * dt_;
        o[1] = s[1] + rv[1] * dt_;
        o[2] = s[2] + rv[2] * dt_;
        quat_mul(s + 3, rot, o + 3);
    }
};

// test/UsckfUnitTest.cpp:34-49.  STATE12 q-vector: pos quat velo angvelo.  u = velocity angvel.
template <>
struct ProcessModel<SLB_PM_USCKF_TEST> {
    static constexpr int NU = 6;
    double rot[4], vel[3], w[3], dt;
    SLB_DEV void prepare(const double *u, double dt_) {
        dt = dt_;
        const double ax[3] = {u[3] * dt_, u[4] * dt_, u[5] * dt_};
        so3_exp(ax, 1.0, rot);
#pragma unroll
        for (int i = 0; i < 3; ++i) { vel[i] = u[i]; w[i] = u[3 + i]; }
    }
    SLB_DEV void apply(const double *s, double *o) const {
        quat_mul(s + 3, rot, o + 3);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            o[10 + i] = w[i];
            o[7 + i] = vel[i];
            o[i] = s[i] + s[7 + i] * dt;
        }
    }
};

// test/MsckfUnitTest.cpp:33-47.  u = dp(3) dq(w,x,y,z) velocity(3) angvel(3).
template <>
struct ProcessModel<SLB_PM_MSCKF_DELTAPOSE> {
    static constexpr int NU = 13;
    double dp[3], dq[4], vel[3], w[3];
    SLB_DEV void prepare(const double *u, double) {
#pragma unroll
        for (int i = 0; i < 3; ++i) { dp[i] = u[i]; vel[i] = u[7 + i]; w[i] = u[10 + i]; }
#pragma unroll
        for (int i = 0; i < 4; ++i) dq[i] = u[3 + i];
    }
    SLB_DEV void apply(const double *s, double *o) const {
        quat_mul(s + 3, dq, o + 3);
        double t[3];
        quat_rotate(o + 3, dp, t);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            o[i] = s[i] + t[i];
            o[7 + i] = vel[i];
            o[10 + i] = w[i];
        }
    }
};

// test/UKFoMUnitTest.cpp:82-85: z = pos
struct MmGpsPos {
    static constexpr int M = 3;
    SLB_DEV static void apply(const double *s, double *z) { z[0] = s[0]; z[1] = s[1]; z[2] = s[2]; }
};

// Eigen Quaternion::toRotationMatrix() * v (Affine3d path of test/UsckfUnitTest.cpp:70-79)
SLB_DEV void rotmat_apply(const double q[4], const double v[3], double o[3]) {
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    o[0] = (1 - (tyy + tzz)) * v[0] + (txy - twz) * v[1] + (txz + twy) * v[2];
    o[1] = (txy + twz) * v[0] + (1 - (txx + tzz)) * v[1] + (tyz - twx) * v[2];
    o[2] = (txz - twy) * v[0] + (tyz + twx) * v[1] + (1 - (txx + tyy)) * v[2];
}

}  // namespace slbd
