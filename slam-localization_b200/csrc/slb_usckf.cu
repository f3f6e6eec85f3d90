// slb_usckf.cu -- localization::Usckf<AugmentedState,State>: predict, update, cloning,
// setMeasurement (src/filters/Usckf.hpp), one filter instance per WARP.
//
// Instance record in HBM: mu[qstride] (statek 13 | statek_l 13 | statek_i 13 | featuresk | featuresk_l)
// and the lower triangle of Pk packed row-major, P[T(i)+j], T(i) = i(i+1)/2, padded to 128 B.
// A record is contiguous, so a warp streams it with one TMA bulk copy (cp.async.bulk -> UBLKCP).
//
// predict (Usckf.hpp:113-244) touches only statek_i: rows 24..35 of the packed triangle (one
//   contiguous 366-double span) and the 12-wide segments of the feature rows.  Lane s evaluates sigma
//   point s (25 of 32 lanes); Fk = Pxy^T Pii^-1 is obtained as W^T L^-1 with W = 0.5 (dY+ - dY-) by a
//   triangular solve, because X_i [-] mu_old is +-L e_j by construction (same algebra as :152-154,
//   without the explicit inverse).
// update (Usckf.hpp:260-308): the 48x48 Cholesky runs right-looking on 2D-cyclic REGISTER tiles in its
//   square-root-free form (see CholStep); finished columns are published through shared memory.
//   Sigma points are evaluated lane-per-point for the columns that can move h (j < 36+nk); the rest
//   equal Z0.
//   P -= K S K^T is applied straight to the HBM record (re-read through L2), coalesced.
#include <cstdlib>

#include "slb_predict12.cuh"
#include "slb_usckf_step.cuh"

namespace slbd {

// =====================================================================================================
// update with the VO measurement model of test/UsckfUnitTest.cpp:62-86 (m = NK)
// =====================================================================================================
template <int NK, int NL>
struct UpdCfg : CycCfg<36 + NK + NL> {
    static constexpr int N = 36 + NK + NL;
    static constexpr int NP = N * (N + 1) / 2;
    static constexpr int JM = 36 + NK;          // columns j >= JM cannot move h
    static constexpr int NSIG = 2 * JM + 1;     // sigma points that are actually evaluated
    static constexpr int RT = (N + 3) / 4;      // register tile: rows i = a + 4r, r < RT
    static constexpr int CT = (N + 7) / 8;      //                cols j = b + 8c, c < CT
    static constexpr int LS = NP + 16;          // factor, column-major packed (+ pad for the tile overhang)
    static constexpr int ZW = NSIG * NK + JM * NK;
    static constexpr int KK = 2 * N * NK;
    static constexpr int SCR = ZW > KK ? ZW : KK;  // Z | W, later overlaid by K | KS
    static constexpr int SM = (LS + SCR + N + 1) / 2 * 2 + 4;  // + measurement slot (NK <= 3)
};

// one instance, one warp; `smem` is the warp's private slice
template <int NK, int NL>
SLB_DEV void usckf_update_one(const slb::FilterArgs &a, int inst, double *smem_w, int lane) {
    typedef UpdCfg<NK, NL> C;
    constexpr int N = C::N, JM = C::JM, NSIG = C::NSIG, RT = C::RT, CT = C::CT;
    static_assert(NK == 3, "the 3x3 closed-form S^-1 is the only one wired so far");
    static_assert(N > 32 && N <= 64, "two rows per lane in the L W product");
    const int w = 0;
    double *smem = smem_w;
    const int a_ = lane & 3, b_ = lane >> 2;
    double *Ls = smem + (size_t)w * C::SM, *Zs = Ls + C::LS, *Ws = Zs + NSIG * NK, *Ks = Zs, *KSs = Zs + N * NK,
           *dl = Zs + C::SCR;
    double *Pg = a.P + (size_t)inst * a.pstride;
    double *mug = a.mu + (size_t)inst * a.qstride;
    // the measurement is fetched now (LDGSTS into the slot after the record scratch): with zero-copy *_step_host it
    // lives in mapped host memory and its PCIe latency must not sit in the middle of the update
    double *zs = smem + C::SM - 4;
    if (lane < NK) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n cp.async.commit_group;\n" ::"r"((unsigned)__cvta_generic_to_shared(zs + lane)),
                     "l"(a.z + (size_t)inst * NK + lane)
                     : "memory");
    }

    // ---- the lower triangle of Pk straight from the HBM record into the 2D-cyclic register tiles ----------
    // (for a fixed tile the 32 lanes read 4 rows x 8 consecutive doubles: full 32-byte sectors)
    double T[RT][CT];
    int rowoff[RT];  // tri(a + 4r, 0) + b
#pragma unroll
    for (int r = 0; r < RT; ++r) {
        const int i = a_ + 4 * r;
        rowoff[r] = i * (i + 1) / 2 + b_;
#pragma unroll
        for (int c = 0; c < CT; ++c)
            if (C::exists(r, c)) {
                const int j = b_ + 8 * c;
                T[r][c] = (i < N && j <= i) ? Pg[rowoff[r] + 8 * c] : 0.0;
            }
    }
    // ---- Eigen::LLT of Pk (:537) ------------------------------------------------------------------------
    bool ok = true;
    {
        const double x0 = __shfl_sync(FULL, T[0][0], 0);
        CholStep<C, 0>::run(T, Ls, a_, b_, ok, x0, -rcp_fast(x0));
    }
    if (!ok) {
        if (lane == 0) a.status[inst] |= SLB_ST_CHOL_FAIL;
        return;
    }
    __syncwarp();
    // The tiles were consumed by the factorisation; Pk -= K S K^T needs Pk again.  Re-issue the same loads now
    // (L2 hits) so their latency hides behind the sigma-point / gain phases instead of stalling the epilogue.
#pragma unroll
    for (int r = 0; r < RT; ++r) {
        const int i = a_ + 4 * r;
#pragma unroll
        for (int c = 0; c < CT; ++c)
            if (C::exists(r, c)) {
                const int j = b_ + 8 * c;
                T[r][c] = (i < N && j <= i) ? __ldcg(Pg + rowoff[r] + 8 * c) : 0.0;
            }
    }

    // mean blocks that h needs (statek pos/orient, statek_i pos/orient, featuresk)
    double pk[3], qk[4], pi[3], qi[4], ft[NK];
#pragma unroll
    for (int c = 0; c < 3; ++c) { pk[c] = mug[c]; pi[c] = mug[26 + c]; }
#pragma unroll
    for (int c = 0; c < 4; ++c) { qk[c] = mug[3 + c]; qi[c] = mug[29 + c]; }
#pragma unroll
    for (int c = 0; c < NK; ++c) ft[c] = mug[39 + c];

    // this lane's share of the mean for the final mu [+] K nu (lane b < 12 owns block b of the three States, lanes
    // 12.. own the feature scalars): fetched here, after the factorisation (registers), long before the epilogue needs it
    const int msidx = lane >> 2, mbw = lane & 3;
    const int mqo = lane < 12 ? 13 * msidx + (mbw == 0 ? 0 : mbw == 1 ? 3 : mbw == 2 ? 7 : 10) : 39 + lane - 12;
    double mym[4] = {0.0, 0.0, 0.0, 0.0};
    if (lane < 12) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (c < 3 || mbw == 1) mym[c] = mug[mqo + c];
    } else if (lane - 12 < NK + NL) {
        mym[0] = mug[mqo];
    }

    // ---- sigma points through h (:275-278), lane per point ---------------------------------------------
    constexpr int NPASS = (NSIG + 31) / 32;
    double zr[NPASS][NK];
#pragma unroll
    for (int t = 0; t < NPASS; ++t) {
        const int s = lane + 32 * t;
        const bool act = s < NSIG;
        const int j = act && s >= 1 ? (s - 1) >> 1 : 0;
        const int cj = j * N - j * (j - 1) / 2 - j;  // column j of the factor starts at Ls[cj + j]
        // L(:,j) = U(:,j) / sqrt(d_j), d_j = U(j,j): the column scale rides on the sigma point's sign
        double sq_, rs_;
        sqrt_rsqrt(Ls[cj + j], sq_, rs_);
        const double sgn = (s & 1) ? rs_ : -rs_;
        if (act && (s & 1)) dl[j] = rs_;
        auto Lc = [&](int r) -> double { return (act && s >= 1 && r >= j) ? sgn * Ls[cj + r] : 0.0; };
        // column j perturbs rows >= j only: from pass 1 on (j >= 15) statek is untouched, in pass 2 (j >= 31) statek_i too
        const int jmin = t == 0 ? 0 : 16 * t - 1;  // compile-time after unrolling
        double xpk[3], xqk[4], xpi[3], xqi[4], xf[NK];
        if (jmin <= 5) {
#pragma unroll
            for (int c = 0; c < 3; ++c) xpk[c] = pk[c] + Lc(c);
            const double v[3] = {Lc(3), Lc(4), Lc(5)};
            double e[4];
            so3_exp(v, 1.0, e);
            quat_mul(qk, e, xqk);
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) xpk[c] = pk[c];
#pragma unroll
            for (int c = 0; c < 4; ++c) xqk[c] = qk[c];
        }
        if (jmin <= 29) {
#pragma unroll
            for (int c = 0; c < 3; ++c) xpi[c] = pi[c] + Lc(24 + c);
            const double v[3] = {Lc(27), Lc(28), Lc(29)};
            double e[4];
            so3_exp(v, 1.0, e);
            quat_mul(qi, e, xqi);
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) xpi[c] = pi[c];
#pragma unroll
            for (int c = 0; c < 4; ++c) xqi[c] = qi[c];
        }
#pragma unroll
        for (int c = 0; c < NK; ++c) xf[c] = ft[c] + Lc(36 + c);
        // h: delta = statek [-] statek_i as a transform, applied to every 3-D feature of featuresk
        double dq[4];
        quat_cmul(xqi, xqk, dq);
#pragma unroll
        for (int c = 0; c < NK; c += 3) {
            double rz[3];
            rotmat_apply(dq, xf + c, rz);
            zr[t][c] = rz[0] + (xpk[0] - xpi[0]);
            zr[t][c + 1] = rz[1] + (xpk[1] - xpi[1]);
            zr[t][c + 2] = rz[2] + (xpk[2] - xpi[2]);
        }
        if (act) {
#pragma unroll
            for (int c = 0; c < NK; ++c) Zs[s * NK + c] = zr[t][c];
        } else {
#pragma unroll
            for (int c = 0; c < NK; ++c) zr[t][c] = 0.0;
        }
    }
    // ---- mean / innovation covariance (:280-282); the 2(N-JM) untouched points all equal Z0 ------------
    constexpr double NREST = 2.0 * (N - JM);
    double z0[NK], zbar[NK];
#pragma unroll
    for (int c = 0; c < NK; ++c) {
        z0[c] = bcast(zr[0][c], 0);
        double s = 0.0;
#pragma unroll
        for (int t = 0; t < NPASS; ++t) s += zr[t][c];
        zbar[c] = (warp_sum(s) + NREST * z0[c]) * (1.0 / (double)(2 * N + 1));
    }
    double S[NK * (NK + 1) / 2];
#pragma unroll
    for (int r = 0; r < NK; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) {
            double s = 0.0;
#pragma unroll
            for (int t = 0; t < NPASS; ++t)
                if (lane + 32 * t < NSIG) s += (zr[t][r] - zbar[r]) * (zr[t][c] - zbar[c]);
            s = warp_sum(s) + NREST * (z0[r] - zbar[r]) * (z0[c] - zbar[c]);
            S[tri(r, c)] = 0.5 * s + __ldg(a.R + r * NK + c);
        }
    __syncwarp();
    // W[j] = 0.5 (Z+_j - Z-_j): the only part of covXZ's right factor that survives the +- pairing
    // (times 1/sqrt(d_j), so that covXZ = L W = U W')
    for (int e = lane; e < JM * NK; e += 32) {
        const int jj = e / NK, c = e - jj * NK;
        Ws[e] = (0.5 * dl[jj]) * ((Zs[(1 + 2 * jj) * NK + c] - zbar[c]) - (Zs[(2 + 2 * jj) * NK + c] - zbar[c]));
    }
    __syncwarp();
    // ---- covXZ = L W (:283, :714-737): lane owns rows `lane` and `lane + 32`; column jj of the factor is
    //      contiguous in i, so the reads are conflict-free -------------------------------------------------
    const bool hasB = lane + 32 < N;
    double pxA[NK], pxB[NK];
#pragma unroll
    for (int c = 0; c < NK; ++c) pxA[c] = pxB[c] = 0.0;
#pragma unroll
    for (int jj = 0; jj < JM; ++jj) {
        constexpr int dummy = 0;
        (void)dummy;
        const int cj = C::cb(jj) - jj;
        const double la = (jj < 32 && lane >= jj) ? Ls[cj + lane] : 0.0;
        const double lb = (hasB && lane + 32 >= jj) ? Ls[cj + lane + 32] : 0.0;
#pragma unroll
        for (int c = 0; c < NK; ++c) {
            const double wv = Ws[jj * NK + c];
            if (jj < 32) pxA[c] = fma(la, wv, pxA[c]);
            pxB[c] = fma(lb, wv, pxB[c]);
        }
    }
    // ---- K = covXZ S^-1 (:286-288), innovation, Mahalanobis gate (:290-294) ----------------------------
    double Si[6];
    sym3_inverse(S, Si);
    auto SiAt = [&](int r, int c) { return r >= c ? Si[tri(r, c)] : Si[tri(c, r)]; };
    auto SAt = [&](int r, int c) { return r >= c ? S[tri(r, c)] : S[tri(c, r)]; };
    double nu[NK], m2 = 0.0;
#pragma unroll
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncwarp();
#pragma unroll
    for (int c = 0; c < NK; ++c) nu[c] = zs[c] - zbar[c];
#pragma unroll
    for (int r = 0; r < NK; ++r) {
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < NK; ++c) s += SiAt(r, c) * nu[c];
        m2 += nu[r] * s;
    }
    const bool accept = chi2_accept(m2, a.gate);
    __syncwarp();  // Z | W are dead from here: K | KS overlay them
    auto finish_row = [&](const double *px, int i) {
        double K[NK], dsum = 0.0;
#pragma unroll
        for (int c = 0; c < NK; ++c) {
            double s = 0.0;
#pragma unroll
            for (int p = 0; p < NK; ++p) s += px[p] * SiAt(p, c);
            K[c] = s;
            dsum += s * nu[c];
        }
#pragma unroll
        for (int c = 0; c < NK; ++c) {
            double s = 0.0;
#pragma unroll
            for (int p = 0; p < NK; ++p) s += K[p] * SAt(p, c);
            Ks[i * NK + c] = K[c];
            KSs[i * NK + c] = s;
        }
        dl[i] = dsum;
    };
    finish_row(pxA, lane);
    if (hasB) finish_row(pxB, lane + 32);
    __syncwarp();
    if (!accept) {
        if (lane == 0) a.status[inst] |= SLB_ST_GATE_REJECT;
        return;
    }
    // ---- mu = mu [+] K nu (:299-301): lane b < 12 owns block b, lanes 12.. own the feature scalars -----
    bool finite = true;
    if (lane < 12) {
        const double v[3] = {dl[3 * lane], dl[3 * lane + 1], dl[3 * lane + 2]};
        if (mbw == 1) {
            double e[4], o[4];
            so3_exp(v, 1.0, e);
            quat_mul(mym, e, o);
#pragma unroll
            for (int c = 0; c < 4; ++c) { mug[mqo + c] = o[c]; finite = finite && isfinite(o[c]); }
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const double o = mym[c] + v[c];
                mug[mqo + c] = o;
                finite = finite && isfinite(o);
            }
        }
    } else if (lane - 12 < NK + NL) {
        const double o = mym[0] + dl[36 + lane - 12];
        mug[mqo] = o;
        finite = finite && isfinite(o);
    }
    // ---- Pk -= K S K^T (:296) on the HBM record (lower triangle), same 2D-cyclic tiles as the load ------
    {
        double kc[CT][NK];
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            const int j = b_ + 8 * c;
#pragma unroll
            for (int p = 0; p < NK; ++p) kc[c][p] = j < N ? Ks[j * NK + p] : 0.0;
        }
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            const int i = a_ + 4 * r;
            double ks[NK];
#pragma unroll
            for (int p = 0; p < NK; ++p) ks[p] = i < N ? KSs[i * NK + p] : 0.0;
#pragma unroll
            for (int c = 0; c < CT; ++c)
                if (C::exists(r, c)) {
                    const int j = b_ + 8 * c;
                    if (i < N && j <= i)
                        Pg[rowoff[r] + 8 * c] = T[r][c] - (ks[0] * kc[c][0] + ks[1] * kc[c][1] + ks[2] * kc[c][2]);
                }
        }
    }
    if (!__all_sync(FULL, finite) && lane == 0) a.status[inst] |= SLB_ST_NONFINITE;
}

// (A persistent variant -- warps walking instances with cp.async.bulk.prefetch.L2 of the next record -- measured
// slower on B200, 6.36 vs 6.08 ms per 524 288 updates: the walk costs registers the factorisation has none to spare.)
template <int NK, int NL, int WPB, int MINB>
__global__ void __launch_bounds__(WPB * 32, MINB) usckf_update_kernel(slb::FilterArgs a) {
    typedef UpdCfg<NK, NL> C;
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int inst = blockIdx.x * WPB + w;
    if (inst >= a.B) return;
    usckf_update_one<NK, NL>(a, inst, smem + (size_t)w * C::SM, lane);
    // optional instance-major copy of the posterior mean (mapped host memory in the zero-copy *_step_host): whatever the
    // update decided (accepted, gated, factorisation failed) the record now holds the posterior
    if (a.mu_out) {
        __syncwarp();
        const double *mug = a.mu + (size_t)inst * a.qstride;
        constexpr int QD = 39 + NK + NL;
        for (int e = lane; e < QD; e += 32) a.mu_out[(size_t)inst * QD + e] = mug[e];
    }
}

// =====================================================================================================
// cloning (Usckf.hpp:391-433) and setMeasurement (:322-389): pure data movement, thread per element
// =====================================================================================================
__global__ void usckf_clone_kernel(slb::FilterArgs a, int mode) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int per = 13 + 12 * 12 * 2;  // 13 mean scalars + two 12x12 blocks' worth of work items
    if (t >= (int64_t)a.B * per) return;
    const int inst = (int)(t / per), e = (int)(t - (int64_t)inst * per);
    double *P = a.P + (size_t)inst * a.pstride, *mu = a.mu + (size_t)inst * a.qstride;
    auto sym = [&](int r, int c) { return r >= c ? P[tri(r, c)] : P[tri(c, r)]; };
    if (e < 13) {
        if (mode == SLB_STATEK_I) mu[13 + e] = mu[26 + e];   // statek_l = statek_i (:401)
        else mu[e] = mu[13 + e];                              // statek = statek_l (:419)
        return;
    }
    const int q = e - 13, blk = q / 144, rc = q - blk * 144, r = rc / 12, c = rc - r * 12;
    if (mode == SLB_STATEK_I) {
        if (blk == 0) {
            // Pk+l = Pk+i (:405) and Pk+i|k+l = Pk+i (:406-407)
            const double v = sym(24 + r, 24 + c);
            if (c <= r) P[tri(12 + r, 12 + c)] = v;
            P[tri(24 + r, 12 + c)] = v;
        } else {
            // cross blocks with statek zeroed (:410-413)
            P[tri(24 + r, c)] = 0.0;
            P[tri(12 + r, c)] = 0.0;
        }
    } else if (blk == 0) {
        // Pk = Pk+l, Pk|k+l = Pk+l (:422-425)
        const double v = sym(12 + r, 12 + c);
        if (c <= r) P[tri(r, c)] = v;
        P[tri(12 + r, c)] = v;
    }
}
// The cloning kernel reads blocks other threads overwrite only in STATEK_I blk 0 (reads P_ii, writes
// P_ll / P_il) and STATEK_L blk 0 (reads P_ll, writes P_kk / P_lk): sources and destinations are disjoint.

__global__ void usckf_set_measurement_kernel(slb::FilterArgs a, int mode) {
    const int N = 36 + a.nk + a.nl, NP = N * (N + 1) / 2;
    const int nfe = NP - 666;  // packed entries of rows 36..N-1
    const int per = nfe + (mode == SLB_STATEK ? a.nk : a.nl);
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)a.B * per) return;
    const int inst = (int)(t / per), e = (int)(t - (int64_t)inst * per);
    double *P = a.P + (size_t)inst * a.pstride, *mu = a.mu + (size_t)inst * a.qstride;
    const int len = mode == SLB_STATEK ? a.nk : a.nl;
    const int f0 = mode == SLB_STATEK ? 0 : a.nk;  // first feature index being replaced
    if (e >= nfe) {
        const int c = e - nfe;
        mu[39 + f0 + c] = a.z[(size_t)inst * len + c];  // featuresk / featuresk_l = z (:335,:362)
        return;
    }
    int r = 36;
    while (tri(r + 1, 0) - 666 <= e) ++r;
    const int c = e - (tri(r, 0) - 666);
    const int fr = r - 36, fc = c - 36;
    double v = 0.0;  // Pk.setZero() (:348,:376): every state<->feature and k<->k+l cross term
    if (c >= 36) {
        const bool rin = fr >= f0 && fr < f0 + len, cin = fc >= f0 && fc < f0 + len;
        if (rin && cin) v = a.R[(fr - f0) * len + (fc - f0)];      // new block = R (:353,:381)
        else if (!rin && !cin) v = P[tri(r, c)];                    // the block that stays (:355,:383)
    }
    P[tri(r, c)] = v;
}

}  // namespace slbd

namespace slb {

template <int PM>
static int launch_predict_t(const FilterArgs &a, cudaStream_t s) {
    constexpr int WPB = 8;
    constexpr size_t smem = (size_t)WPB * slbd::PRED_SM * sizeof(double);
    auto kern = slbd::predict12_kernel<PM, WPB, 24, 26, true>;
    SLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(a.B + WPB - 1) / WPB, WPB * 32, smem, s>>>(a);
    count_launch();
    SLB_CUDA(cudaGetLastError());
    return SLB_OK;
}

template <int NK, int NL, int MINB>
static int launch_update_m(const FilterArgs &a, cudaStream_t s) {
    constexpr int WPB = 4;
    constexpr size_t smem = (size_t)WPB * slbd::UpdCfg<NK, NL>::SM * sizeof(double);
    auto kern = slbd::usckf_update_kernel<NK, NL, WPB, MINB>;
    SLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(a.B + WPB - 1) / WPB, WPB * 32, smem, s>>>(a);
    count_launch();
    SLB_CUDA(cudaGetLastError());
    return SLB_OK;
}

// experiment knob: SLB_USCKF_MINB=4 selects the 128-register build (4 CTAs of 4 warps per SM)
template <int NK, int NL>
static int launch_update_t(const FilterArgs &a, cudaStream_t s) {
    static const int minb = [] { const char *e = getenv("SLB_USCKF_MINB"); return e ? atoi(e) : 3; }();
    return minb == 4 ? launch_update_m<NK, NL, 4>(a, s) : launch_update_m<NK, NL, 3>(a, s);
}

// record-resident kernel (slb_usckf_step.cuh): PRED && UPD = the fused step, UPD alone = update
template <int NK, int NL, bool PRED, bool UPD>
static int launch_step_t(const FilterArgs &a, cudaStream_t s) {
    typedef slbd::StepCfg<NK, NL> C;
    constexpr int WPB = 4;
    constexpr size_t smem = (size_t)WPB * C::SM * sizeof(double);
    constexpr int fit = (int)((228 * 1024) / (smem + 1024));
    constexpr int MINB = fit > 3 ? 3 : fit < 1 ? 1 : fit;
    static_assert(smem <= 227 * 1024, "record + scratch of one CTA must fit in shared memory");
    auto kern = slbd::usckf_step_kernel<SLB_PM_USCKF_TEST, NK, NL, PRED, UPD, WPB, MINB>;
    SLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // L2 prefetch distance in instances: SLB_USCKF_PREFETCH waves of (SMs x resident warps); default 2 waves
    static const int ahead = [] {
        const char *e = getenv("SLB_USCKF_PREFETCH");
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        return (e ? atoi(e) : 2) * sms * WPB * MINB;
    }();
    FilterArgs b = a;
    b.prefetch = ahead;
    // experiment knob: SLB_USCKF_CTAS=1|2 pads the dynamic shared memory so that only that many CTAs fit per SM
    static const size_t pad = [] {
        const char *e = getenv("SLB_USCKF_CTAS");
        const int n = e ? atoi(e) : 0;
        return n == 1 ? (size_t)(200 * 1024) - smem : n == 2 ? (size_t)(110 * 1024) - smem : (size_t)0;
    }();
    if (pad) SLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem + pad)));
    kern<<<(a.B + WPB - 1) / WPB, WPB * 32, smem + pad, s>>>(b);
    count_launch();
    SLB_CUDA(cudaGetLastError());
    return SLB_OK;
}

template <int NK, int NL>
static int launch_step_nknl(bool predict, const FilterArgs &a, cudaStream_t s) {
    return predict ? launch_step_t<NK, NL, true, true>(a, s) : launch_step_t<NK, NL, false, true>(a, s);
}

int launch_usckf(int pm, int mm, bool predict, bool update, const FilterArgs &a, cudaStream_t s) {
    if (predict && pm != SLB_PM_USCKF_TEST) return set_error(SLB_ERR_INVALID, "usckf: unsupported process model");
    if (update && mm != SLB_MM_USCKF_VO) return set_error(SLB_ERR_INVALID, "usckf: unsupported measurement model");
    // experiment knob: SLB_USCKF_LEGACY=1 selects the round-1 two-launch path (predict12_kernel + usckf_update_kernel)
    static const bool legacy = [] { const char *e = getenv("SLB_USCKF_LEGACY"); return e && atoi(e) != 0; }();
    if (predict && (!update || legacy)) {
        int rc = launch_predict_t<SLB_PM_USCKF_TEST>(a, s);
        if (rc != SLB_OK) return rc;
    }
    if (update && legacy) {
        if (a.nk == 3 && a.nl == 9) return launch_update_t<3, 9>(a, s);
        if (a.nk == 3 && a.nl == 0) return launch_update_t<3, 0>(a, s);
        return set_error(SLB_ERR_INVALID, "usckf update (legacy): built for (nk,nl) = (3,9) and (3,0)");
    }
    if (update) {
        if (a.nk == 3 && a.nl == 9) return launch_step_nknl<3, 9>(predict, a, s);
        if (a.nk == 3 && a.nl == 0) return launch_step_nknl<3, 0>(predict, a, s);
        return set_error(SLB_ERR_INVALID, "usckf update: built for (nk,nl) = (3,9) and (3,0)");
    }
    return SLB_OK;
}

int launch_usckf_clone(int mode, const FilterArgs &a, cudaStream_t s) {
    if (mode != SLB_STATEK_I && mode != SLB_STATEK_L) return SLB_OK;  // default: break (Usckf.hpp:428)
    const int64_t work = (int64_t)a.B * (13 + 288);
    slbd::usckf_clone_kernel<<<(unsigned)((work + 255) / 256), 256, 0, s>>>(a, mode);
    count_launch();
    SLB_CUDA(cudaGetLastError());
    return SLB_OK;
}

int launch_usckf_set_measurement(int mode, const FilterArgs &a, cudaStream_t s) {
    if (mode != SLB_STATEK && mode != SLB_STATEK_L) return SLB_OK;
    const int N = 36 + a.nk + a.nl;
    const int len = mode == SLB_STATEK ? a.nk : a.nl;
    if (len == 0) return SLB_OK;
    const int64_t work = (int64_t)a.B * (N * (N + 1) / 2 - 666 + len);
    slbd::usckf_set_measurement_kernel<<<(unsigned)((work + 255) / 256), 256, 0, s>>>(a, mode);
    count_launch();
    SLB_CUDA(cudaGetLastError());
    return SLB_OK;
}

}  // namespace slb
