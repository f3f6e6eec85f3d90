// facade_test.cpp -- the reference's own test flows (test/DataModelUnitTest.cpp:30-67,
// test/UKFoMUnitTest.cpp:93-117, test/UsckfUnitTest.cpp:175-284, test/MsckfUnitTest.cpp:151-213)
// re-written against the host C++ facade.  Prints `key v0 v1 ...` lines that tests/test_facade.py
// compares with the CPU oracle.  Exits 3 with "NO_DEVICE" when there is no GPU (no CPU fallback).
#include <cmath>
#include <cstdio>

#include "../slam-localization_b200/facade/localization_b200.hpp"

using namespace slb200;

static void print(const char *key, const Vec &v, size_t n) {
    std::printf("%s", key);
    for (size_t i = 0; i < n && i < v.size(); ++i) std::printf(" %.17g", v[i]);
    std::printf("\n");
}

int main() {
    try {
        const double D2R = M_PI / 180.0;
        {   // DATAMODEL
            DataModel<3> data1, data2;
            data1.data = {0.0124889, 0.00171945, -0.0138983};
            data2.data = {0.0168381, 0.000632167, -0.0235605};
            DataModel<3> data3 = data1 + data2;
            print("dm_plus_x", data3.data, 3);
            print("dm_plus_C", data3.Cov, 9);
            data3 = data1 - data2;
            print("dm_minus_x", data3.data, 3);
            print("dm_minus_C", data3.Cov, 9);
            data3 = data1;
            data3.fusion(data2);
            print("dm_fusion_x", data3.data, 3);
            print("dm_fusion_C", data3.Cov, 9);
            data3 = data1;                                   // test/DataModelUnitTest.cpp:72-76
            data3.safeFusion(data2);
            print("dm_safe_x", data3.data, 3);
            print("dm_safe_C", data3.Cov, 9);
        }
        {   // UKFOM
            Vec mu0 = {0, 0, 0, 1, 0, 0, 0, 0, 0, 0}, P0(81, 0.0), Q(81, 0.0), R(9, 0.0);
            for (int i = 0; i < 9; ++i) P0[i * 9 + i] = 0.001;
            const double dt = 0.01;
            for (int i = 3; i < 6; ++i) Q[i * 9 + i] = 0.0001 * dt;
            for (int i = 6; i < 9; ++i) Q[i * 9 + i] = 0.0002 * dt;
            for (int i = 0; i < 3; ++i) R[i * 3 + i] = 0.00000001;
            Ukf filter(1, SLB_LAYOUT_MTK9, mu0, P0);
            Vec u = {0.0, 0.0, 0.0, 10.0 * D2R, 0.0, 0.0}, gps = {1.0, 0.0, 0.0};
            filter.predict(SLB_PM_UKFOM_IMU_REFBUG, u, dt, Q);
            filter.update(gps, SLB_MM_GPS_POS, R);
            print("ukfom_mu", filter.mu(), 10);
            print("ukfom_sigma", filter.sigma(), 81);
        }
        {   // USCKF_DYNAMIC
            Vec single = {0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0}, P0(144, 0.0);
            for (int i = 0; i < 12; ++i) P0[i * 12 + i] = 0.0025;
            Usckf filter(1, 3, 9, single, P0, true);
            Vec vo(3, 3.34), voC(9, 0.0), icp(9, 1.34), icpC(81, 0.0), vo2(3, 3.35), vo2C(9, 0.0);
            for (int i = 0; i < 3; ++i) { voC[i * 3 + i] = 0.008; vo2C[i * 3 + i] = 0.05; }
            for (int i = 0; i < 9; ++i) icpC[i * 9 + i] = 0.008;
            filter.setMeasurement(STATEK, vo, voC);
            filter.setMeasurement(STATEK_L, icp, icpC);
            filter.setMeasurement(STATEK, vo2, vo2C);
            const double dt = 0.01;
            Vec Q(144, 0.0);
            for (int i = 0; i < 12; ++i) Q[i * 12 + i] = 0.1 * dt;
            Vec u = {100.0, 0.0, 0.0, 100.0 * D2R, 100.0 * D2R, 100.0 * D2R};
            for (int i = 0; i < 2; ++i) {
                filter.predict(SLB_PM_USCKF_TEST, u, dt, Q);
                Vec Pi = filter.PkSingleState(STATEK_I);
                double tr = 0;
                for (int d = 0; d < 12; ++d) tr += Pi[d * 12 + d];
                print("usckf_trace", Vec{tr}, 1);
            }
            print("usckf_mu", filter.muState(), 51);
            Vec z = {2.33, 3.35, 3.35}, R(9, 0.0);
            for (int i = 0; i < 3; ++i) R[i * 3 + i] = 0.01;
            std::printf("usckf_check %d\n", filter.checkSigmaPoints()[0]);   // indefinite ctor-#2 covariance: LLT fails (bit 2)
            filter.update(z, SLB_MM_USCKF_VO, R);
            std::printf("usckf_status %d\n", filter.status()[0]);
        }
        {   // MSCKF: ctor + two predicts with the delta-pose process model
            const int k = 4, N = 12 + 6 * k, QD = 13 + 7 * k;
            Vec mu(QD, 0.0), P((size_t)N * N, 0.0), Q(144, 0.0);
            mu[3] = 1.0;
            for (int c = 0; c < k; ++c) mu[13 + 7 * c + 3] = 1.0;
            for (int i = 0; i < N; ++i) P[(size_t)i * N + i] = 0.025;
            for (int i = 0; i < 12; ++i) Q[i * 12 + i] = 0.01;
            Msckf filter(1, k, mu, P);
            // delta orientation Rz(1 deg) Ry(1 deg) Rx(1 deg) as (w,x,y,z)
            const double h = 0.5 * D2R, c = std::cos(h), s = std::sin(h);
            Vec u = {0.1, 0.1, 0.1, c * c * c + s * s * s, s * c * c - c * s * s, c * s * c + s * c * s, c * c * s - s * s * c,
                     0.1, 0.1, 0.1, 0.1, 0.1, 0.1};
            for (int i = 0; i < 2; ++i) filter.predict(SLB_PM_MSCKF_DELTAPOSE, u, 0.0, Q);
            print("msckf_mu", filter.muSingleState(), 13);
            print("msckf_P", filter.getPkSingleState(), 144);
            // checkSigmaPoints (Msckf.hpp:818-838), the setters muSingleState(state) / setPkSingleState (:351,:363) and the
            // (single-GPU, communicator-less) ensemble-statistics gather
            std::printf("msckf_check %d\n", filter.checkSigmaPoints()[0]);
            Vec st13 = filter.muSingleState(), P12 = filter.getPkSingleState();
            st13[0] += 1.0;
            for (int i = 0; i < 12; ++i) P12[i * 12 + i] *= 2.0;
            filter.muSingleState(st13);
            filter.setPkSingleState(P12);
            print("msckf_mu_set", filter.muSingleState(), 13);
            print("msckf_P_set", filter.getPkSingleState(), 144);
            print("msckf_stats", filter.gatherStats(), 4);
        }
        {   // SURVEY 8f rows f2 / f4 through the facade: one error-state EKF cycle, one dead-reckoning step
            Vec state(48, 0.0), error(45, 0.0), P0(45 * 45, 0.0), F(225, 0.0), Q(225, 0.0), H(3 * 45, 0.0), R(9, 0.0);
            for (int s3 = 0; s3 < 3; ++s3) state[16 * s3 + 6] = 1.0;
            state[32] = 1.0;  // statek_i.pos.x
            for (int i = 0; i < 45; ++i) P0[i * 45 + i] = 0.01;
            for (int i = 0; i < 15; ++i) { F[i * 15 + i] = 1.0; Q[i * 15 + i] = 1e-4; }
            for (int i = 0; i < 3; ++i) { F[i * 15 + 3 + i] = 0.01; H[i * 45 + 15 + i] = -1.0; H[i * 45 + 30 + i] = 1.0; R[i * 3 + i] = 0.0025; }
            UsckfError filter(state, error, P0);
            filter.ekfPredict(F, Q);
            Vec ret = filter.ekfUpdate({1.02, 0.01, -0.01}, H, R, true);
            print("ekf_ret", ret, 3);
            print("ekf_P", filter.PkAugmentedState(), 45 * 45);
            PoseWithUncertainty prev, post;
            prev.pose = {1, 2, 3, 1, 0, 0, 0};
            prev.cov.assign(36, 0.0);
            Vec velcov(36, 0.0);
            for (int i = 0; i < 6; ++i) { prev.cov[i * 6 + i] = 1e-3; velcov[i * 6 + i] = i < 3 ? 1e-2 : 1e-3; }
            PoseWithUncertainty delta = DeadReckon::updatePose(0.01, {1, 0, 0, 0, 0, 0.1}, {1, 0, 0, 0, 0, 0.1}, velcov, prev, post);
            print("dr_post", post.pose, 7);
            print("dr_post_cov", post.cov, 36);
            print("dr_delta", delta.pose, 7);
        }
        std::printf("OK\n");
        return 0;
    } catch (const Error &e) {
        if (e.code == SLB_ERR_NO_DEVICE) {
            std::printf("NO_DEVICE %s\n", e.what());
            return 3;
        }
        std::printf("ERROR %d %s\n", e.code, e.what());
        return 1;
    }
}
