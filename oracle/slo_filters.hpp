// oracle/slo_filters.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the reference's sigma-point filters and covariance fusion, following
//   src/filters/Usckf.hpp  (Usckf<AugmentedState,State>)
//   src/filters/Msckf.hpp  (Msckf<MultiState,State>, UKF-flavoured update)
//   ukfom::ukf<state>      (third party, not in tree; same skeleton as the Msckf single-state
//                           helpers Msckf.hpp:435-496,554-570,612-633,668-675; call pattern at
//                           test/UKFoMUnitTest.cpp:104-117)
//   src/core/DataModel.hpp (DataModel<double,D>::fusion, operator+/-)
// PARITY UNPINNED: see the header of slo_core.hpp.  Each function cites the lines it follows.
#pragma once
#include <functional>

#include "slo_core.hpp"

namespace slo {

enum StatusBits {
    ST_CHOL_FAIL = 1,     // LLT hit a non-positive pivot (reference ignores it, Q8)
    ST_MEAN_NOCONV = 2,   // manifold mean hit max_it (reference asserts, Q7)
    ST_GATE_REJECT = 4,   // significance test rejected the update
    ST_NONFINITE = 8
};

typedef std::function<Vec(const Vec &)> Model;

// ---- sigma-point helpers shared by all three filters ---------------------------------------

// Usckf.hpp:572-598 / Msckf.hpp:407-431,442-468: X0 = mu[+]delta, X(2j+1) = mu[+](delta+Lj),
// X(2j+2) = mu[+](delta-Lj); unscaled (Q1).
inline int sigma_points_vec(const Layout &l, const Vec &mu, const Vec &delta, const Mat &P,
                            std::vector<Vec> &X) {
    const int n = l.dof();
    Mat L;
    const int info = llt_lower(P, L);
    X.assign(2 * n + 1, Vec());
    X[0] = boxplus(l, mu, delta);
    Vec d(n);
    for (int j = 0; j < n; ++j) {
        for (int i = 0; i < n; ++i) d[i] = delta[i] + L(i, j);
        X[1 + 2 * j] = boxplus(l, mu, d);
        for (int i = 0; i < n; ++i) d[i] = delta[i] - L(i, j);
        X[2 + 2 * j] = boxplus(l, mu, d);
    }
    return info;
}

// Usckf.hpp:601-627 / Msckf.hpp:471-525: iterative mean on the manifold (Q7).
inline Vec mean_manifold(const Layout &l, const std::vector<Vec> &X, int *status, int *iters = nullptr) {
    Vec ref = X[0];
    const int n = l.dof();
    Vec md(n);
    const size_t max_it = 10000;
    size_t it = 0;
    do {
        std::fill(md.begin(), md.end(), 0.0);
        for (const Vec &x : X) {
            Vec d = boxminus(l, x, ref);
            for (int i = 0; i < n; ++i) md[i] += d[i];
        }
        for (int i = 0; i < n; ++i) md[i] /= (double)X.size();
        ref = boxplus(l, ref, md);
    } while (norm2(md) > 1e-6 && ++it < max_it);
    if (it >= max_it && status) *status |= ST_MEAN_NOCONV;
    if (iters) *iters = (int)it + 1;
    return ref;
}

// Usckf.hpp:630-640 / Msckf.hpp:528-538: arithmetic mean of measurement sigma points.
inline Vec mean_vector(const std::vector<Vec> &Z) {
    Vec m(Z[0].size(), 0.0);
    for (const Vec &z : Z)
        for (size_t i = 0; i < m.size(); ++i) m[i] += z[i];
    for (double &x : m) x /= (double)Z.size();
    return m;
}

// Usckf.hpp:654-670 / Msckf.hpp:554-589: 0.5 * sum_i (Vi [-] mean)(Vi [-] mean)^T.
inline Mat cov_manifold(const Layout &l, const Vec &mean, const std::vector<Vec> &V) {
    const int n = l.dof();
    Mat c(n, n);
    for (const Vec &v : V) {
        Vec d = boxminus(l, v, mean);
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) c(i, j) += d[i] * d[j];
    }
    for (double &x : c.a) x *= 0.5;
    return c;
}
// Usckf.hpp:672-689 / Msckf.hpp:593-610
inline Mat cov_vector(const Vec &mean, const std::vector<Vec> &V) {
    const int m = (int)mean.size();
    Mat c(m, m);
    for (const Vec &v : V)
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < m; ++j) c(i, j) += (v[i] - mean[i]) * (v[j] - mean[j]);
    for (double &x : c.a) x *= 0.5;
    return c;
}
// Usckf.hpp:691-712 / Msckf.hpp:612-657: 0.5 * sum_i (Xi [-] meanX)(Zi - meanZ)^T, with Z either a
// plain vector (zl == nullptr) or itself a manifold point (predict's X-before / X-after pair).
inline Mat crosscov(const Layout &lx, const Vec &meanX, const std::vector<Vec> &X,
                    const Layout *lz, const Vec &meanZ, const std::vector<Vec> &Z) {
    const int n = lx.dof();
    const int m = lz ? lz->dof() : (int)meanZ.size();
    Mat c(n, m);
    for (size_t s = 0; s < X.size(); ++s) {
        Vec dx = boxminus(lx, X[s], meanX);
        Vec dz(m);
        if (lz) dz = boxminus(*lz, Z[s], meanZ);
        else for (int j = 0; j < m; ++j) dz[j] = Z[s][j] - meanZ[j];
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < m; ++j) c(i, j) += dx[i] * dz[j];
    }
    for (double &x : c.a) x *= 0.5;
    return c;
}

// Usckf.hpp:794-855 / Msckf.hpp:844-905: chi-square 5% gate, dof 1..9, anything else rejects.
inline bool accept_mahalanobis_distance(double m2, int dof) {
    static const double th[10] = {0, 3.84, 5.99, 7.81, 9.49, 11.07, 12.59, 14.07, 15.51, 16.92};
    if (dof < 1 || dof > 9) return false;
    return m2 < th[dof];
}

// ============================================================================================
// ukfom::ukf<state>  [restated from memory of MTK's ukfom/ukf.hpp; same skeleton as Msckf]
// ============================================================================================
struct Ukf {
    Layout lay;
    Vec mu;
    Mat sigma;
    int status = 0;
    int last_mean_iters = 0;

    Ukf(const Layout &l, const Vec &mu0, const Mat &P0) : lay(l), mu(mu0), sigma(P0) {}

    // predict(g, R): sigma points -> g -> manifold mean -> cov + R   
    void predict(const Model &g, const Mat &Q) {
        std::vector<Vec> X;
        Vec zero(lay.dof(), 0.0);
        if (sigma_points_vec(lay, mu, zero, sigma, X) >= 0) status |= ST_CHOL_FAIL;
        for (Vec &x : X) x = g(x);
        mu = mean_manifold(lay, X, &status, &last_mean_iters);
        sigma = add(cov_manifold(lay, mu, X), Q);
    }
    // update(z, h, R, mt): gate_dof = 0 -> accept_any_mahalanobis_distance
    bool update(const Vec &z, const Model &h, const Mat &R, int gate_dof = 0) {
        std::vector<Vec> X;
        Vec zero(lay.dof(), 0.0);
        if (sigma_points_vec(lay, mu, zero, sigma, X) >= 0) status |= ST_CHOL_FAIL;
        std::vector<Vec> Z(X.size());
        for (size_t i = 0; i < X.size(); ++i) Z[i] = h(X[i]);
        const Vec meanZ = mean_vector(Z);
        const Mat S = add(cov_vector(meanZ, Z), R);
        const Mat covXZ = crosscov(lay, mu, X, nullptr, meanZ, Z);
        const Mat Sinv = inverse_lu(S);
        const Mat K = matmul(covXZ, Sinv);
        Vec innov(z.size());
        for (size_t i = 0; i < z.size(); ++i) innov[i] = z[i] - meanZ[i];
        const Vec Si = matvec(Sinv, innov);
        double m2 = 0;
        for (size_t i = 0; i < z.size(); ++i) m2 += innov[i] * Si[i];
        const bool ok = gate_dof == 0 ? true : accept_mahalanobis_distance(m2, gate_dof);
        if (ok) {
            sigma = sub(sigma, matmul(matmul(K, S), K.transpose()));
            apply_delta(matvec(K, innov));
        } else {
            status |= ST_GATE_REJECT;
        }
        return ok;
    }
    // apply_delta (Msckf.hpp:668-675 shape): re-draw sigma points around mu[+]delta, re-estimate.
    void apply_delta(const Vec &delta) {
        std::vector<Vec> X;
        if (sigma_points_vec(lay, mu, delta, sigma, X) >= 0) status |= ST_CHOL_FAIL;
        mu = mean_manifold(lay, X, &status, &last_mean_iters);
        sigma = cov_manifold(lay, mu, X);
    }
};

// ============================================================================================
// localization::Usckf<AugmentedState<Dynamic>, State>   (src/filters/Usckf.hpp)
// q-vector of the augmented state: statek(13) statek_l(13) statek_i(13) featuresk featuresk_l.
// ============================================================================================
enum CloningMode { STATEK = 1, STATEK_L = 2, STATEK_I = 3 };

struct Usckf {
    static const int NS = 12;       // State::DOF
    static const int QS = 13;       // q-size of State
    static const int NA = 36;       // AugmentedState::DOF (static part)
    int nk = 0, nl = 0;             // featuresk.size(), featuresk_l.size()
    Vec mu;                         // 39 + nk + nl
    Mat Pk;                         // (36+nk+nl)^2
    int status = 0;
    int last_mean_iters = 0;

    Layout single() const { return Layout::state12(); }
    Layout aug() const { return Layout::augmented(nk, nl); }
    int dof() const { return NA + nk + nl; }

    // ctor #1 (Usckf.hpp:83)
    Usckf(const Vec &state, int nk_, int nl_, const Mat &P0) : nk(nk_), nl(nl_), mu(state), Pk(P0) {}
    // ctor #2 (Usckf.hpp:90-103): statek_i = single; P_ii = P0; cloning(I); cloning(L)  (Q13)
    Usckf(const Vec &single_state, const Mat &P0_single) {
        mu = Layout::augmented(0, 0).identity();
        for (int i = 0; i < QS; ++i) mu[2 * QS + i] = single_state[i];
        Pk = Mat(NA, NA);
        Pk.set_block(24, 24, P0_single);
        cloning(STATEK_I);
        cloning(STATEK_L);
    }

    Vec sub_state(int which) const {  // which: 0 statek, 1 statek_l, 2 statek_i
        return Vec(mu.begin() + which * QS, mu.begin() + (which + 1) * QS);
    }

    // Usckf.hpp:113-244
    void predict(const Model &f, const Mat &Q) {
        const Layout ls = single();
        const Vec statek_i = sub_state(2);
        Mat Pk_i = Pk.block(24, 24, NS, NS);
        std::vector<Vec> X;
        Vec zero(NS, 0.0);
        if (sigma_points_vec(ls, statek_i, zero, Pk_i, X) >= 0) status |= ST_CHOL_FAIL;  // :130
        const std::vector<Vec> XCopy = X;                                              // :133
        for (Vec &x : X) x = f(x);                                                     // :141
        const Vec mean = mean_manifold(ls, X, &status, &last_mean_iters);              // :148
        for (int i = 0; i < QS; ++i) mu[2 * QS + i] = mean[i];
        const Mat Pxy = crosscov(ls, statek_i, XCopy, &ls, mean, X);                   // :152
        const Mat Fk = matmul(Pxy.transpose(), inverse_lu(Pk_i));                      // :154
        Pk_i = add(cov_manifold(ls, mean, X), Q);                                      // :178
        Pk.set_block(24, 24, Pk_i);                                                    // :181
        const Mat FkT = Fk.transpose();
        Pk.set_block(0, 24, matmul(Pk.block(0, 24, NS, NS), FkT));                     // :191-193
        Pk.set_block(12, 24, matmul(Pk.block(12, 24, NS, NS), FkT));                   // :196-198
        Pk.set_block(24, 0, matmul(Fk, Pk.block(24, 0, NS, NS)));                      // :201-203
        Pk.set_block(24, 12, matmul(Fk, Pk.block(24, 12, NS, NS)));                    // :206-208
        if (nk > 0) {                                                                  // :222-227
            const Mat Pzk = matmul(Fk, Pk.block(24, NA, NS, nk));
            Pk.set_block(24, NA, Pzk);
            Pk.set_block(NA, 24, Pzk.transpose());
        }
        if (nl > 0) {                                                                  // :230-235
            const Mat Pzkl = matmul(Fk, Pk.block(24, NA + nk, NS, nl));
            Pk.set_block(24, NA + nk, Pzkl);
            Pk.set_block(NA + nk, 24, Pzkl.transpose());
        }
    }

    // Usckf.hpp:532-561 sigma points of the full augmented state; deltas travel as states (Q10)
    int sigma_points_aug(const Vec &delta, std::vector<Vec> &X) const {
        const Layout la = aug();
        const int N = dof();
        Mat L;
        const int info = llt_lower(Pk, L);
        const Vec delta_state = set_from_vector(la, delta);                   // :546-547
        X.assign(2 * N + 1, Vec());
        X[0] = aug_plus_state(la, mu, delta_state);                           // :549
        Vec col(N);
        for (int j = 0; j < N; ++j) {
            for (int i = 0; i < N; ++i) col[i] = L(i, j);
            const Vec l_state = set_from_vector(la, col);                     // :553-554
            X[1 + 2 * j] = aug_plus_state(la, mu, aug_plus_state(la, delta_state, l_state));   // :555
            X[2 + 2 * j] = aug_plus_state(la, mu, aug_minus_state(la, delta_state, l_state));  // :556
        }
        return info;
    }

    // Usckf.hpp:260-308.  gate_dof = 0 mirrors accept_any_mahalanobis_distance (:249,257).
    bool update(const Vec &z, const Model &h, const Mat &R, int gate_dof = 0) {
        const Layout la = aug();
        const int N = dof();
        std::vector<Vec> X;
        Vec zero(N, 0.0);
        if (sigma_points_aug(zero, X) >= 0) status |= ST_CHOL_FAIL;           // :275
        std::vector<Vec> Z(X.size());
        for (size_t i = 0; i < X.size(); ++i) Z[i] = h(X[i]);                 // :278
        const Vec meanZ = mean_vector(Z);                                      // :280
        const Mat S = add(cov_vector(meanZ, Z), R);                            // :282
        // :283 -> :714-737: (Xi - meanX) is a state; it is re-vectorised with log before use
        const int m = (int)meanZ.size();
        Mat covXZ(N, m);
        for (size_t s = 0; s < X.size(); ++s) {
            const Vec dx = get_vectorized(la, aug_minus_state(la, X[s], mu));
            for (int i = 0; i < N; ++i)
                for (int j = 0; j < m; ++j) covXZ(i, j) += dx[i] * (Z[s][j] - meanZ[j]);
        }
        for (double &x : covXZ.a) x *= 0.5;
        const Mat Sinv = inverse_lu(S);                                        // :286
        const Mat K = matmul(covXZ, Sinv);                                     // :288
        Vec innov(m);
        for (int i = 0; i < m; ++i) innov[i] = z[i] - meanZ[i];                // :290
        const Vec Si = matvec(Sinv, innov);
        double m2 = 0;
        for (int i = 0; i < m; ++i) m2 += innov[i] * Si[i];                    // :292
        const bool ok = gate_dof == 0 ? true : accept_mahalanobis_distance(m2, gate_dof);
        if (ok) {
            Pk = sub(Pk, matmul(matmul(K, S), K.transpose()));                 // :296 (Q2)
            const Vec innovation_state = set_from_vector(la, matvec(K, innov));  // :299-300
            mu = aug_plus_state(la, mu, innovation_state);                     // :301 (Q3)
        } else {
            status |= ST_GATE_REJECT;
        }
        return ok;
    }

    // Usckf.hpp:391-433
    void cloning(int mode) {
        if (mode == STATEK_I) {
            for (int i = 0; i < QS; ++i) mu[QS + i] = mu[2 * QS + i];          // :401
            const Mat Pi = Pk.block(24, 24, NS, NS);
            Pk.set_block(12, 12, Pi); Pk.set_block(12, 24, Pi); Pk.set_block(24, 12, Pi);  // :404-407
            const Mat Z(NS, NS);
            Pk.set_block(0, 24, Z); Pk.set_block(24, 0, Z);                    // :410-411
            Pk.set_block(0, 12, Z); Pk.set_block(12, 0, Z);                    // :412-413
        } else if (mode == STATEK_L) {
            for (int i = 0; i < QS; ++i) mu[i] = mu[QS + i];                   // :419
            const Mat Pl = Pk.block(12, 12, NS, NS);
            Pk.set_block(0, 0, Pl); Pk.set_block(0, 12, Pl); Pk.set_block(12, 0, Pl);      // :422-425
        }
    }

    // Usckf.hpp:322-389
    void set_measurement(int mode, const Vec &z, const Mat &R) {
        const Mat Pstates = Pk.block(0, 0, NA, NA);                            // :329
        Vec fk(mu.begin() + 3 * QS, mu.begin() + 3 * QS + nk);
        Vec fl(mu.begin() + 3 * QS + nk, mu.end());
        if (mode == STATEK) {
            // NB :342 reads the surviving block at an offset computed with the *new* featuresk size
            const int nk_new = (int)z.size();
            Mat Pz_l;
            if (nl > 0) Pz_l = safe_block(NA + nk_new, NA + nk_new, nl, nl);
            fk = z; nk = nk_new;
            Pk = Mat(NA + nk + nl, NA + nk + nl);                              // :346-348
            Pk.set_block(NA, NA, R);                                           // :353
            if (nl > 0) Pk.set_block(NA + nk, NA + nk, Pz_l);                  // :354-355
        } else if (mode == STATEK_L) {
            Mat Pz_k;
            if (nk > 0) Pz_k = Pk.block(NA, NA, nk, nk);                       // :366-370
            fl = z; nl = (int)z.size();
            Pk = Mat(NA + nk + nl, NA + nk + nl);                              // :374-376
            Pk.set_block(NA + nk, NA + nk, R);                                 // :381
            if (nk > 0) Pk.set_block(NA, NA, Pz_k);                            // :382-383
        }
        Pk.set_block(0, 0, Pstates);                                           // :388
        Vec m(mu.begin(), mu.begin() + 3 * QS);
        m.insert(m.end(), fk.begin(), fk.end());
        m.insert(m.end(), fl.begin(), fl.end());
        mu = m;
    }

    // Reading outside the old matrix is undefined behaviour in the reference (it happens when the
    // size of featuresk changes, :342); the restatement returns zeros there and flags nothing.
    Mat safe_block(int i0, int j0, int rows, int cols) const {
        Mat m(rows, cols);
        for (int i = 0; i < rows; ++i)
            for (int j = 0; j < cols; ++j)
                if (i0 + i < Pk.r && j0 + j < Pk.c) m(i, j) = Pk(i0 + i, j0 + j);
        return m;
    }
};

// ============================================================================================
// localization::Msckf<MultiState<State,SensorState>, State>   (src/filters/Msckf.hpp)
// ============================================================================================
struct Msckf {
    static const int NS = 12, QS = 13;
    int k = 0;   // sensorsk.size()
    Vec mu;      // 13 + 7k
    Mat Pk;      // (12+6k)^2
    int status = 0;
    int last_mean_iters = 0;

    Msckf(int k_, const Vec &mu0, const Mat &P0) : k(k_), mu(mu0), Pk(P0) {}
    Layout multi() const { return Layout::multi(k); }
    int dof() const { return NS + 6 * k; }

    // Msckf.hpp:97-189: top-left block only; cross blocks are left stale (Q5); Fk unused (Q15).
    void predict(const Model &f, const Mat &Q) {
        const Layout ls = Layout::state12();
        const Vec statek(mu.begin(), mu.begin() + QS);
        const Mat Pk_i = Pk.block(0, 0, NS, NS);
        std::vector<Vec> X;
        Vec zero(NS, 0.0);
        if (sigma_points_vec(ls, statek, zero, Pk_i, X) >= 0) status |= ST_CHOL_FAIL;  // :114
        for (Vec &x : X) x = f(x);                                                     // :125
        const Vec mean = mean_manifold(ls, X, &status, &last_mean_iters);              // :132
        for (int i = 0; i < QS; ++i) mu[i] = mean[i];
        Pk.set_block(0, 0, add(cov_manifold(ls, mean, X), Q));                         // :162-165
    }

    // Msckf.hpp:723-754 with the row-index quirk Q6 reproduced: rows/cols {2i, 2i+2} of the
    // *current* arrays are deleted (second index not re-based after the first deletion), except
    // for the last block where the second deletion falls off the end and drops the last row.
    static void remove_at(std::vector<int> &idx, unsigned pos) {
        const unsigned num = (unsigned)idx.size() - 1;
        if (pos < num) idx.erase(idx.begin() + pos);
        else idx.resize(num);
    }
    unsigned remove_outliers(Vec &innov, Mat &covXZ, Mat &S, std::vector<int> *kept_out = nullptr) {
        const unsigned dofb = 2;
        std::vector<int> kept(innov.size());
        for (size_t i = 0; i < kept.size(); ++i) kept[i] = (int)i;
        unsigned outliers = 0, i = 0;
        while (i < kept.size() / dofb) {
            const int a = kept[dofb * i], b = kept[dofb * i + 1];
            Mat blk(2, 2);
            blk(0, 0) = S(a, a); blk(0, 1) = S(a, b); blk(1, 0) = S(b, a); blk(1, 1) = S(b, b);
            const Mat bi = inverse_lu(blk);
            const double v0 = innov[a], v1 = innov[b];
            const double m2 = v0 * (bi(0, 0) * v0 + bi(0, 1) * v1) + v1 * (bi(1, 0) * v0 + bi(1, 1) * v1);
            if (!accept_mahalanobis_distance(m2, (int)dofb)) {
                remove_at(kept, dofb * i);
                remove_at(kept, dofb * i + 1);
                ++outliers;
            } else {
                ++i;
            }
        }
        const int m = (int)kept.size();
        Vec in2(m);
        Mat S2(m, m), C2(covXZ.r, m);
        for (int p = 0; p < m; ++p) {
            in2[p] = innov[kept[p]];
            for (int q = 0; q < m; ++q) S2(p, q) = S(kept[p], kept[q]);
            for (int r = 0; r < covXZ.r; ++r) C2(r, p) = covXZ(r, kept[p]);
        }
        innov = in2; S = S2; covXZ = C2;
        if (kept_out) *kept_out = kept;
        return outliers;
    }

    // Msckf.hpp:220-277 (+ applyDelta :659-666)
    unsigned update(const Vec &z, const Model &h, const Mat &R, bool gate = true) {
        const Layout lm = multi();
        const int N = dof();
        std::vector<Vec> X;
        Vec zero(N, 0.0);
        if (sigma_points_vec(lm, mu, zero, Pk, X) >= 0) status |= ST_CHOL_FAIL;   // :229
        std::vector<Vec> Z(X.size());
        for (size_t i = 0; i < X.size(); ++i) Z[i] = h(X[i]);                     // :232
        const Vec mean_z = mean_vector(Z);                                         // :234
        Vec innov(z.size());
        for (size_t i = 0; i < z.size(); ++i) innov[i] = z[i] - mean_z[i];         // :236
        Mat S = add(cov_vector(mean_z, Z), R);                                     // :238
        Mat covXZ = crosscov(lm, mu, X, nullptr, mean_z, Z);                       // :239
        unsigned outliers = 0;
        if (gate) outliers = remove_outliers(innov, covXZ, S);                     // :241
        if (!innov.empty()) {                                                      // :250
            const Mat K = matmul(covXZ, inverse_lu(S));                            // :257
            Pk = sub(Pk, matmul(matmul(K, S), K.transpose()));                     // :262
            apply_delta(matvec(K, innov));                                         // :263
        }
        return outliers;                                                           // :276
    }
    void apply_delta(const Vec &delta) {                                           // :659-666
        const Layout lm = multi();
        std::vector<Vec> X;
        if (sigma_points_vec(lm, mu, delta, Pk, X) >= 0) status |= ST_CHOL_FAIL;
        mu = mean_manifold(lm, X, &status, &last_mean_iters);
        Pk = cov_manifold(lm, mu, X);
    }
};

// ============================================================================================
// localization::DataModel<double, D>   (src/core/DataModel.hpp)
// ============================================================================================
struct DataModel {
    Vec data;
    Mat Cov;
    explicit DataModel(int d) : data(d, 0.0), Cov(d, d) {                      // :32-36
        for (int i = 0; i < d; ++i) Cov(i, i) = 1.0e-10;                       // ZERO_UNCERTAINTY
    }
    DataModel(const Vec &x, const Mat &C) : data(x), Cov(C) {}                 // :38-41
    // :48-60  P=(C1^-1+C2^-1)^-1 ; x = P (C1^-1 x1 + C2^-1 x2); inverses as Eigen's fixed-size ones
    void fusion(const DataModel &o) {
        const Mat I1 = inverse_fixed(Cov), I2 = inverse_fixed(o.Cov);
        const Mat P = inverse_fixed(add(I1, I2));
        const Vec a = matvec(I1, data), b = matvec(I2, o.data);
        Vec s(a.size());
        for (size_t i = 0; i < s.size(); ++i) s[i] = a[i] + b[i];
        data = matvec(P, s);
        Cov = P;
    }
    DataModel plus(const DataModel &o) const {                                 // :132-141
        DataModel r = *this;
        for (size_t i = 0; i < data.size(); ++i) r.data[i] = data[i] + o.data[i];
        r.Cov = add(Cov, o.Cov);
        return r;
    }
    DataModel minus(const DataModel &o) const {                                // :143-152 (Cov ADDS)
        DataModel r = *this;
        for (size_t i = 0; i < data.size(); ++i) r.data[i] = data[i] - o.data[i];
        r.Cov = add(Cov, o.Cov);
        return r;
    }
};

}  // namespace slo
