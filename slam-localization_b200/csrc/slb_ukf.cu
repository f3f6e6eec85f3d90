// slb_ukf.cu -- ukfom::ukf<state> predict / update / fused step, one filter instance per thread.
//
// Replaces (per instance) ukfom::ukf::predict / update / apply_delta as exercised at
// test/UKFoMUnitTest.cpp:104-117, whose skeleton is the single-state path of Msckf.hpp
// (:435-496 sigma points + manifold mean, :554-570 covariance, :612-633 cross-covariance,
// :668-675 applyDelta).
//
// Mapping (see DESIGN.md "UKF kernel"):
//   * state is SoA in HBM: mu[c][stride], P[e][stride] with P packed lower-triangular, so the 32
//     lanes of a warp read 256 contiguous bytes per field;
//   * mean and covariance live in registers (all loops over compile-time layouts are unrolled);
//   * the 2n+1 propagated sigma points of one instance live in a private column of shared memory
//     (element e of thread t at sm[e*TPB + t]: conflict-free), with the Cholesky factor parked in
//     the not-yet-written tail of the same column;
//   * predict+update fused in one launch moves each instance through HBM exactly once.
#include "slb_internal.h"
#include "slb_models.cuh"

namespace slbd {

template <class L, int TPB>
struct Col {
    static constexpr int N = L::N, QD = L::QD, NP = L::NP, NS = 2 * L::N + 1;
    static constexpr int NSQ = NS * QD;  // doubles per thread column
    double *sm;
    int tid;
    SLB_DEV double &at(int e) const { return sm[e * TPB + tid]; }
    // sigma point s, component c
    SLB_DEV double &Y(int s, int c) const { return at(s * QD + c); }
    // Cholesky factor, column-major packed from the END of the column (col N-1 last)
    SLB_DEV double &Lt(int r, int j) const { return at(NSQ - (N - j) * (N - j + 1) / 2 + (r - j)); }
};

// y = x [+] sign*d, knowing d[r] == 0 for r < j0 (column j0 of a lower-triangular factor)
template <class L>
SLB_DEV void boxplus_from(const double *x, const double *d, int j0, double sign, double *y) {
#pragma unroll
    for (int b = 0; b < L::NB; ++b) {
        const int o = L::qoff(b);
        if (3 * b + 2 < j0) {
#pragma unroll
            for (int i = 0; i < (L::so3(b) ? 4 : 3); ++i) y[o + i] = x[o + i];
        } else if (L::so3(b)) {
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

// Msckf.hpp:471-496 / Usckf.hpp:601-627: reference = X0; do { d = mean(Xi [-] ref); ref [+]= d }
// while (|d| > 1e-6 && ++i < 10000).  Returns false if the cap was hit.
template <class L, int TPB>
SLB_DEV bool manifold_mean(const Col<L, TPB> &col, double *ref) {
    constexpr int N = L::N, QD = L::QD, NS = 2 * N + 1;
#pragma unroll
    for (int c = 0; c < QD; ++c) ref[c] = col.Y(0, c);
    int it = 0;
    double nrm2;
    do {
        double md[N];
#pragma unroll
        for (int r = 0; r < N; ++r) md[r] = 0.0;
#pragma unroll 2
        for (int s = 0; s < NS; ++s) {
            double y[QD], d[N];
#pragma unroll
            for (int c = 0; c < QD; ++c) y[c] = col.Y(s, c);
            boxminus<L>(y, ref, d);
#pragma unroll
            for (int r = 0; r < N; ++r) md[r] += d[r];
        }
        nrm2 = 0.0;
#pragma unroll
        for (int r = 0; r < N; ++r) {
            md[r] = md[r] / (double)NS;
            nrm2 += md[r] * md[r];
        }
        double nr[QD];
        boxplus<L>(ref, md, 1.0, nr);
#pragma unroll
        for (int c = 0; c < QD; ++c) ref[c] = nr[c];
    } while (sqrt(nrm2) > 1e-6 && ++it < 10000);
    return it < 10000;
}

// Msckf.hpp:554-570: 0.5 * sum (Yi [-] mean)(Yi [-] mean)^T, packed lower
template <class L, int TPB>
SLB_DEV void manifold_cov(const Col<L, TPB> &col, const double *mean, double *C) {
    constexpr int N = L::N, QD = L::QD, NP = L::NP, NS = 2 * N + 1;
#pragma unroll
    for (int e = 0; e < NP; ++e) C[e] = 0.0;
#pragma unroll 1
    for (int s = 0; s < NS; ++s) {
        double y[QD], d[N];
#pragma unroll
        for (int c = 0; c < QD; ++c) y[c] = col.Y(s, c);
        boxminus<L>(y, mean, d);
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) C[tri(i, j)] += d[i] * d[j];
    }
#pragma unroll
    for (int e = 0; e < NP; ++e) C[e] *= 0.5;
}

// Factor P (registers, destroyed) and park L in the column tail.
template <class L, int TPB>
SLB_DEV bool factor_to_tail(const Col<L, TPB> &col, double *P) {
    constexpr int N = L::N;
    const bool ok = chol_packed<N>(P);
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int r = j; r < N; ++r) col.Lt(r, j) = P[tri(r, j)];
    return ok;
}

template <class L, int TPB>
SLB_DEV void load_col(const Col<L, TPB> &col, int j, double *d) {
#pragma unroll
    for (int r = 0; r < L::N; ++r) d[r] = (r >= j) ? col.Lt(r < j ? j : r, j) : 0.0;
}

template <class L, int PM, class MM, int TPB, bool PRED, bool UPD>
__global__ void __launch_bounds__(TPB) ukf_kernel(slb::FilterArgs a) {
    constexpr int N = L::N, QD = L::QD, NP = L::NP, NS = 2 * N + 1, M = MM::M, MP = M * (M + 1) / 2;
    static_assert(NS * M + 2 * NP <= NS * QD, "column too small for the update scratch");
    extern __shared__ double sm[];
    const int tid = threadIdx.x;
    const int i = blockIdx.x * TPB + tid;
    if (i >= a.B) return;
    Col<L, TPB> col{sm, tid};

    double mu[QD], P[NP];
#pragma unroll
    for (int c = 0; c < QD; ++c) mu[c] = a.mu[(size_t)c * a.stride + i];
#pragma unroll
    for (int e = 0; e < NP; ++e) P[e] = a.P[(size_t)e * a.stride + i];
    int st = 0;
    bool alive = true;

    if (PRED) {
        // ---- predict(g, Q): sigma points -> g -> manifold mean -> cov + Q ------------------------
        if (!factor_to_tail(col, P)) { st |= SLB_ST_CHOL_FAIL; alive = false; }
        if (alive) {
            ProcessModel<PM> g;
            {
                double u[ProcessModel<PM>::NU];
#pragma unroll
                for (int c = 0; c < ProcessModel<PM>::NU; ++c) u[c] = a.u[(size_t)i * ProcessModel<PM>::NU + c];
                g.prepare(u, a.dt);
            }
            {
                double y[QD];
                g.apply(mu, y);
#pragma unroll
                for (int c = 0; c < QD; ++c) col.Y(0, c) = y[c];
            }
#pragma unroll 1
            for (int j = 0; j < N; ++j) {
                double d[N];
                load_col(col, j, d);
                double xp[QD], xm[QD], yp[QD], ym[QD];
                boxplus_from<L>(mu, d, j, 1.0, xp);
                boxplus_from<L>(mu, d, j, -1.0, xm);
                g.apply(xp, yp);
                g.apply(xm, ym);
#pragma unroll
                for (int c = 0; c < QD; ++c) { col.Y(1 + 2 * j, c) = yp[c]; col.Y(2 + 2 * j, c) = ym[c]; }
            }
            if (!manifold_mean(col, mu)) st |= SLB_ST_MEAN_NOCONV;
            manifold_cov(col, mu, P);
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int c = 0; c <= r; ++c) P[tri(r, c)] += __ldg(a.Q + r * N + c);
        }
    }

    if (UPD && alive) {
        // ---- update(z, h, R, mt) ----------------------------------------------------------------
        constexpr int ZOFF = 0, PSOFF = NS * M;
#pragma unroll
        for (int e = 0; e < NP; ++e) col.at(PSOFF + e) = P[e];  // stash P: needed for P -= K S K^T
        if (!factor_to_tail(col, P)) { st |= SLB_ST_CHOL_FAIL; alive = false; }
        if (alive) {
            double zsum[M];
            {
                double z0[M];
                MM::apply(mu, z0);
#pragma unroll
                for (int c = 0; c < M; ++c) { col.at(ZOFF + c) = z0[c]; zsum[c] = z0[c]; }
            }
#pragma unroll 1
            for (int j = 0; j < N; ++j) {
                double d[N];
                load_col(col, j, d);
                double xp[QD], xm[QD], zp[M], zm[M];
                boxplus_from<L>(mu, d, j, 1.0, xp);
                boxplus_from<L>(mu, d, j, -1.0, xm);
                MM::apply(xp, zp);
                MM::apply(xm, zm);
#pragma unroll
                for (int c = 0; c < M; ++c) {
                    col.at(ZOFF + (1 + 2 * j) * M + c) = zp[c];
                    col.at(ZOFF + (2 + 2 * j) * M + c) = zm[c];
                    zsum[c] += zp[c];
                    zsum[c] += zm[c];
                }
            }
            double zbar[M];
#pragma unroll
            for (int c = 0; c < M; ++c) zbar[c] = zsum[c] / (double)NS;
            // S = 0.5 sum (Zi - zbar)(Zi - zbar)^T + R   (packed lower)
            double S[MP];
#pragma unroll
            for (int e = 0; e < MP; ++e) S[e] = 0.0;
#pragma unroll 1
            for (int s = 0; s < NS; ++s) {
                double dz[M];
#pragma unroll
                for (int c = 0; c < M; ++c) dz[c] = col.at(ZOFF + s * M + c) - zbar[c];
#pragma unroll
                for (int r = 0; r < M; ++r)
#pragma unroll
                    for (int c = 0; c <= r; ++c) S[tri(r, c)] += dz[r] * dz[c];
            }
#pragma unroll
            for (int r = 0; r < M; ++r)
#pragma unroll
                for (int c = 0; c <= r; ++c) S[tri(r, c)] = 0.5 * S[tri(r, c)] + __ldg(a.R + r * M + c);
            // Pxz = 0.5 sum (Xi [-] mu)(Zi - zbar)^T.  Xi [-] mu is +-L e_j by construction (the
            // reference recovers it through log(exp(.)), identical below |.| < pi), and the zbar terms
            // of a +- pair cancel: Pxz = 0.5 sum_j L e_j (Z+_j - Z-_j)^T.
            double Pxz[N * M];
#pragma unroll
            for (int e = 0; e < N * M; ++e) Pxz[e] = 0.0;
#pragma unroll 1
            for (int j = 0; j < N; ++j) {
                double d[N], dz[M];
                load_col(col, j, d);
#pragma unroll
                for (int c = 0; c < M; ++c)
                    dz[c] = (col.at(ZOFF + (1 + 2 * j) * M + c) - zbar[c]) - (col.at(ZOFF + (2 + 2 * j) * M + c) - zbar[c]);
#pragma unroll
                for (int r = 0; r < N; ++r)
#pragma unroll
                    for (int c = 0; c < M; ++c) Pxz[r * M + c] += d[r] * dz[c];
            }
#pragma unroll
            for (int e = 0; e < N * M; ++e) Pxz[e] *= 0.5;
            double Si[MP];
            static_assert(M == 3, "only 3-dof measurement models are wired to the UKF kernel");
            sym3_inverse(S, Si);
            auto SiAt = [&](int r, int c) { return r >= c ? Si[tri(r, c)] : Si[tri(c, r)]; };
            auto SAt = [&](int r, int c) { return r >= c ? S[tri(r, c)] : S[tri(c, r)]; };
            double K[N * M];
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int c = 0; c < M; ++c) {
                    double s = 0.0;
#pragma unroll
                    for (int p = 0; p < M; ++p) s += Pxz[r * M + p] * SiAt(p, c);
                    K[r * M + c] = s;
                }
            double innov[M], m2 = 0.0;
#pragma unroll
            for (int c = 0; c < M; ++c) innov[c] = a.z[(size_t)i * M + c] - zbar[c];
#pragma unroll
            for (int r = 0; r < M; ++r) {
                double s = 0.0;
#pragma unroll
                for (int c = 0; c < M; ++c) s += SiAt(r, c) * innov[c];
                m2 += innov[r] * s;
            }
#pragma unroll
            for (int e = 0; e < NP; ++e) P[e] = col.at(PSOFF + e);
            if (chi2_accept(m2, a.gate)) {
                // sigma -= K S K^T   (lower triangle; the reference's LLT reads only that, Q8)
                double KS[N * M];
#pragma unroll
                for (int r = 0; r < N; ++r)
#pragma unroll
                    for (int c = 0; c < M; ++c) {
                        double s = 0.0;
#pragma unroll
                        for (int p = 0; p < M; ++p) s += K[r * M + p] * SAt(p, c);
                        KS[r * M + c] = s;
                    }
#pragma unroll
                for (int r = 0; r < N; ++r)
#pragma unroll
                    for (int c = 0; c <= r; ++c) {
                        double s = 0.0;
#pragma unroll
                        for (int p = 0; p < M; ++p) s += KS[r * M + p] * K[c * M + p];
                        P[tri(r, c)] -= s;
                    }
                // apply_delta(K * innovation): re-draw sigma points around mu [+] delta
                double delta[N];
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    double s = 0.0;
#pragma unroll
                    for (int c = 0; c < M; ++c) s += K[r * M + c] * innov[c];
                    delta[r] = s;
                }
                if (!factor_to_tail(col, P)) {
                    st |= SLB_ST_CHOL_FAIL;
                    alive = false;
                } else {
                    {
                        double y[QD];
                        boxplus<L>(mu, delta, 1.0, y);
#pragma unroll
                        for (int c = 0; c < QD; ++c) col.Y(0, c) = y[c];
                    }
#pragma unroll 1
                    for (int j = 0; j < N; ++j) {
                        double d[N], dp[N], dm[N];
                        load_col(col, j, d);
#pragma unroll
                        for (int r = 0; r < N; ++r) { dp[r] = delta[r] + d[r]; dm[r] = delta[r] - d[r]; }
                        double yp[QD], ym[QD];
                        boxplus<L>(mu, dp, 1.0, yp);
                        boxplus<L>(mu, dm, 1.0, ym);
#pragma unroll
                        for (int c = 0; c < QD; ++c) { col.Y(1 + 2 * j, c) = yp[c]; col.Y(2 + 2 * j, c) = ym[c]; }
                    }
                    if (!manifold_mean(col, mu)) st |= SLB_ST_MEAN_NOCONV;
                    manifold_cov(col, mu, P);
                }
            } else {
                st |= SLB_ST_GATE_REJECT;
            }
        }
    }

    if (alive) {
        bool finite = true;
#pragma unroll
        for (int c = 0; c < QD; ++c) finite = finite && isfinite(mu[c]);
        if (!finite) st |= SLB_ST_NONFINITE;
#pragma unroll
        for (int c = 0; c < QD; ++c) a.mu[(size_t)c * a.stride + i] = mu[c];
#pragma unroll
        for (int e = 0; e < NP; ++e) a.P[(size_t)e * a.stride + i] = P[e];
    }
    if (st) a.status[i] |= st;
}

}  // namespace slbd

namespace slb {

template <class L, int PM, class MM, int TPB>
static int launch_ukf_t(bool predict, bool update, const FilterArgs &a, cudaStream_t s) {
    constexpr size_t smem = (size_t)TPB * (2 * L::N + 1) * L::QD * sizeof(double);
    static_assert(smem <= 227 * 1024, "sigma-point columns exceed shared memory");
    const int grid = (a.B + TPB - 1) / TPB;
    auto go = [&](auto kern) -> int {
        SLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, TPB, smem, s>>>(a);
        count_launch();
        SLB_CUDA(cudaGetLastError());
        return SLB_OK;
    };
    if (predict && update) return go(slbd::ukf_kernel<L, PM, MM, TPB, true, true>);
    if (predict) return go(slbd::ukf_kernel<L, PM, MM, TPB, true, false>);
    if (update) return go(slbd::ukf_kernel<L, PM, MM, TPB, false, true>);
    return set_error(SLB_ERR_INVALID, "ukf: nothing to do");
}

int launch_ukf(int layout, int pm, int mm, bool predict, bool update, const FilterArgs &a, cudaStream_t s) {
    using namespace slbd;
    if (update && mm != SLB_MM_GPS_POS) return set_error(SLB_ERR_INVALID, "ukf: unsupported measurement model");
    if (!predict) pm = layout == SLB_LAYOUT_POSE6 ? SLB_PM_POSE6_ODOM : SLB_PM_UKFOM_IMU;
    if (layout == SLB_LAYOUT_MTK9 && pm == SLB_PM_UKFOM_IMU)
        return launch_ukf_t<LayMtk9, SLB_PM_UKFOM_IMU, MmGpsPos, 128>(predict, update, a, s);
    if (layout == SLB_LAYOUT_MTK9 && pm == SLB_PM_UKFOM_IMU_REFBUG)
        return launch_ukf_t<LayMtk9, SLB_PM_UKFOM_IMU_REFBUG, MmGpsPos, 128>(predict, update, a, s);
    if (layout == SLB_LAYOUT_POSE6 && pm == SLB_PM_POSE6_ODOM)
        return launch_ukf_t<LayPose6, SLB_PM_POSE6_ODOM, MmGpsPos, 256>(predict, update, a, s);
    return set_error(SLB_ERR_INVALID, "ukf: unsupported layout / process model combination");
}

}  // namespace slb
