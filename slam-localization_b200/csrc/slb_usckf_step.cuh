// slb_usckf_step.cuh -- localization::Usckf predict (Usckf.hpp:107-244) and update (:246-308) of one filter
// instance per WARP with the instance record RESIDENT IN SHARED MEMORY: one TMA bulk load brings the packed lower
// triangle of Pk (and the mean) in, predict and update work on it in place, one TMA bulk store writes it back, so a
// fused predict+update step moves exactly the algorithmic bytes (record in + record out) through HBM.
//
// update: the N x N factorisation is a BLOCKED right-looking one in the square-root-free form Pk = U D^-1 U^T
//   (L(:,k) = U(:,k) / sqrt(d_k), see CholStep in slb_predict12.cuh for the scalar version): panels of 4 columns, the
//   trailing matrix lives in FP64 DMMA accumulator tiles (m8n8k4: lane (a = lane & 3, b = lane >> 2) holds entries
//   (8I + b, 8J + 2a) and (8I + b, 8J + 2a + 1) of tile (I, J)), so one panel's rank-4 update is ONE DMMA per tile fed
//   by one shared-memory word per lane and tile row -- the scalar version needed 4 x (rows + cols) words and 4 FMAs
//   per entry.  Per panel: the 16 lanes holding its columns park them in shared memory, every lane redoes the 4 x 4
//   diagonal block's elimination (4 reciprocals, ~20 FMAs -- the same issue slots as one lane doing it), lane i
//   finishes row i of the panel (6 FMAs), and the finished rows are both the A and (scaled by -1/d_k) the B fragments.
//   Only the columns that can move h are factored (j < 36 + nk, rounded up to a panel): rows below come out of the same
//   panels (the TRSM is implicit), the trailing block of featuresk_l is never touched.
//   The factor is NOT kept: sigma point j = mu [+] +-L(:,j) needs column j only, so after every 4 panels the 32 sigma
//   points of those 16 columns go through h (one per lane), their contribution to the mean / innovation covariance is
//   accumulated as deviations from Z0 = h(mu) (columns that cannot move h contribute nothing), and
//   covXZ += U(:, 16 cols) W'(16 cols, :) with W'_j = (Z+_j - Z-_j) / (2 sqrt d_j)  (:714-737 without the
//   exp/log round trip, quirk Q10).
//   K = covXZ S^-1 (:286-288), K S K^T = covXZ K^T; Pk -= covXZ K^T is again one DMMA per tile, on the record in
//   shared memory.
// predict: slb_predict12.cuh's scheme on rows 24..35 of the resident record.
#pragma once
#include "slb_predict12.cuh"

// Experiment knob (compile time): -DSLB_USCKF_CTA_SYNC=1 re-aligns the warps of a CTA a few times per instance so that
// they walk the (104 KB, fully unrolled) instruction stream together and share instruction-cache lines.
#ifndef SLB_USCKF_CTA_SYNC
#define SLB_USCKF_CTA_SYNC 0
#endif
#if SLB_USCKF_CTA_SYNC
#define SLB_CTA_SYNC() __syncthreads()
#else
#define SLB_CTA_SYNC() ((void)0)
#endif

namespace slbd {

template <int NK_, int NL_>
struct StepCfg {
    static constexpr int NK = NK_, NL = NL_, NF = NK_ + NL_;
    static constexpr int N = 36 + NK_ + NL_;
    static constexpr int NP = N * (N + 1) / 2;
    static constexpr int PSTR = (NP + 15) / 16 * 16;   // == slb_batch_s::pstride
    static constexpr int QD = 39 + NK_ + NL_;
    static constexpr int QS = (QD + 1) / 2 * 2;        // == slb_batch_s::qstride
    static constexpr int NT = (N + 7) / 8;             // 8 x 8 accumulator tiles per side
    static constexpr int NPAD = 8 * NT;
    static constexpr int JM = 36 + NK_;                // columns j >= JM cannot move h
    static constexpr int NPAN = (JM + 3) / 4;          // panels of 4 columns that are factored
    static constexpr int NCOL = 4 * NPAN;
    static constexpr int JLAST = (NCOL - 1) >> 3;      // last tile column that is ever updated
    static constexpr int NGRP = (NPAN + 3) / 4;        // sigma-point passes (16 columns = 32 points each)
    // A finished panel (NPAD rows x 4 columns) is stored as two planes of column PAIRS, element (i, c) at
    // (c >> 1) * PL + 2 i + (c & 1): row accesses (one double2 per lane and plane) and DMMA fragment loads are then
    // contiguous across the warp.  PL = 8 mod 16 puts the two planes half a bank cycle apart (the 16-lane panel
    // extraction writes both planes in one instruction), and with PS4 = 2 mod 16 the 16 columns of a group fall into 16
    // different bank pairs for the sigma points' column reads: (c & 1) + 8 (c >> 1) + 2 q covers 0..15.
    static constexpr int PL = 2 * NPAD + 8;
    static constexpr int PS4 = 2 * PL + 2;             // stride between the 4 panels of a group
    static constexpr int KP = (NK_ + 3) / 4 * 4;       // padded row length of K / covXZ (DMMA k-dimension)
    static constexpr int LF = 4 * PS4;                 // finished panel rows U of the current group
    static constexpr int LRAW = 2 * PL;                // the current panel as extracted from the accumulators (same planes)
    static constexpr int KK = 2 * NPAD * KP + NPAD;    // K | covXZ rows | K nu, overlay LF | LRAW once the factorisation is over
    static constexpr int SCR0 = (LF + LRAW > KK ? LF + LRAW : KK);
    static constexpr int WGS = 16 * NK_;               // W' of the current group, [c][16]: column pairs are one double2
    static constexpr int UPD_SCR = SCR0 + WGS + NPAD;       // + d_k
    static constexpr int PRED_SCR = PRED_LS + 28 * PRED_DS + 144 + 12 * PRED_DS + 16 + 16;   // ... + rsd + reduction slots
    static constexpr int SCR = (UPD_SCR > PRED_SCR ? UPD_SCR : PRED_SCR);
    static constexpr int ZS = (NK_ + 1) / 2 * 2;
    static constexpr int SM = PSTR + QS + SCR + ZS + 2;    // doubles per warp (even); last slot = mbarrier
    static_assert(NK_ % 3 == 0 && NK_ >= 3, "the VO model moves 3-D features");
    static_assert(N <= 64, "two rows per lane");
    static_assert(SM % 2 == 0 && PSTR % 2 == 0 && QS % 2 == 0, "16-byte alignment of the bulk copies");
};

SLB_DEV void bulk_s2g(void *gdst, const void *ssrc, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n cp.async.bulk.commit_group;\n" ::"l"(gdst),
                 "r"(saddr(ssrc)), "r"(bytes)
                 : "memory");
}
SLB_DEV void bulk_prefetch_l2(const void *gsrc, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(gsrc), "r"(bytes) : "memory");
}
SLB_DEV void bulk_store_wait() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
SLB_DEV void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// symmetric inverse of the M x M innovation covariance (packed lower in, packed lower out) through its LDL^T
// factorisation, redundantly on every lane; false if a pivot is not positive
template <int M>
SLB_DEV bool sym_inverse(const double *S, double *Si) {
    double L[M * (M + 1) / 2], dd[M], rd[M];   // unit lower L below the diagonal, D and 1 / D
    bool ok = true;
#pragma unroll
    for (int k = 0; k < M; ++k) {
        double d = S[tri(k, k)];
#pragma unroll
        for (int p = 0; p < k; ++p) d = fma(-L[tri(k, p)] * L[tri(k, p)], dd[p], d);
        ok = ok && (d > 0.0);
        dd[k] = d;
        rd[k] = rcp_fast(d);
#pragma unroll
        for (int i = k + 1; i < M; ++i) {
            double s = S[tri(i, k)];
#pragma unroll
            for (int p = 0; p < k; ++p) s = fma(-L[tri(i, p)] * L[tri(k, p)], dd[p], s);
            L[tri(i, k)] = s * rd[k];
        }
    }
    // X = L^-1 (unit lower), Si = X^T D^-1 X
    double X[M * (M + 1) / 2];
#pragma unroll
    for (int j = 0; j < M; ++j) {
        X[tri(j, j)] = 1.0;
#pragma unroll
        for (int i = j + 1; i < M; ++i) {
            double s = 0.0;
#pragma unroll
            for (int p = j; p < i; ++p) s = fma(-L[tri(i, p)], X[tri(p, j)], s);
            X[tri(i, j)] = s;
        }
    }
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = i; k < M; ++k) s = fma(X[tri(k, i)] * rd[k], X[tri(k, j)], s);
            Si[tri(i, j)] = s;
        }
    return ok;
}

// h of test/UsckfUnitTest.cpp:62-86: delta = statek [-] statek_i as a transform, applied to every 3-D feature
template <int NK>
SLB_DEV void vo_model(const double *pk, const double *qk, const double *pi, const double *qi, const double *f, double *z) {
    double dq[4];
    quat_cmul(qi, qk, dq);
#pragma unroll
    for (int c = 0; c < NK; c += 3) {
        double rz[3];
        rotmat_apply(dq, f + c, rz);
        z[c] = rz[0] + (pk[0] - pi[0]);
        z[c + 1] = rz[1] + (pk[1] - pi[1]);
        z[c + 2] = rz[2] + (pk[2] - pi[2]);
    }
}

// Sums of 12 values over the 32 lanes, result on every lane.  A butterfly per value costs 5 shuffles each (120 SHFL.32 for
// 12 doubles, and shuffles travel through the same data pipe as shared memory, the kernel's limiter); here the lanes halve
// the VALUE set while they halve the lane set (8 + 4 + 2 + 1 + 1 exchanges), the 16 partial owners park their sums in
// shared memory and everybody reads them back with broadcast loads: 32 SHFL.32 + 1 STS + 6 LDS.128.  Fixed order: bitwise
// reproducible.
SLB_DEV void warp_sum12(const double *v, double *out, double *slots, int lane) {
    double a[8], b[4], c[2], d1;
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {   // values 12..15 are zero padding
        const double lo = v[i], hi = i < 4 ? v[8 + i] : 0.0;
        const double r = __shfl_xor_sync(FULL, h16 ? lo : hi, 16);
        a[i] = (h16 ? hi : lo) + r;                     // lanes < 16 own values 0..7, lanes >= 16 own 8..15
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double r = __shfl_xor_sync(FULL, h8 ? a[i] : a[4 + i], 8);
        b[i] = (h8 ? a[4 + i] : a[i]) + r;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const double r = __shfl_xor_sync(FULL, h4 ? b[i] : b[2 + i], 4);
        c[i] = (h4 ? b[2 + i] : b[i]) + r;
    }
    {
        const double r = __shfl_xor_sync(FULL, h2 ? c[0] : c[1], 2);
        d1 = (h2 ? c[1] : c[0]) + r;
    }
    d1 += __shfl_xor_sync(FULL, d1, 1);
    // value index owned by this lane: bit 3 = h16, bit 2 = h8, bit 1 = h4, bit 0 = h2
    const int idx = (h16 ? 8 : 0) + (h8 ? 4 : 0) + (h4 ? 2 : 0) + (h2 ? 1 : 0);
    __syncwarp();
    if (!(lane & 1)) slots[idx] = d1;
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const double2 t = *reinterpret_cast<const double2 *>(slots + 2 * i);
        out[2 * i] = t.x;
        out[2 * i + 1] = t.y;
    }
}

template <class C>
struct UpdState {
    double c0[C::NT][C::JLAST + 1], c1[C::NT][C::JLAST + 1];   // accumulator tiles (I >= J)
    double px0[C::NK], px1[C::NK];                             // covXZ rows lane, lane + 32 (unscaled sum)
    double esum[C::NK], eep[C::NK * (C::NK + 1) / 2];          // sum e, sum e e^T over this lane's sigma points, e = Z - Z0
    double z0[C::NK];
    bool ok;
};

// ---- sigma points of columns 16 G .. 16 G + 15 through h, W' of those columns, covXZ += U W' ---------------------
template <class C, int G>
SLB_DEV void sigma_group(UpdState<C> &s, const double *mus_, const double *Lf, double *Wg, const double *dv, int lane) {
    constexpr int NK = C::NK, PS4 = C::PS4;
    // the mean scalars h needs, as 16-byte broadcast loads (statek pos|quat = mus[0..6], statek_i = mus[26..32])
    double mus[40 + NK];
    {
        const double2 *m2 = reinterpret_cast<const double2 *>(mus_);
#pragma unroll
        for (int e = 0; e < 4; ++e) { const double2 v = m2[e]; mus[2 * e] = v.x; mus[2 * e + 1] = v.y; }
#pragma unroll
        for (int e = 13; e < 17; ++e) { const double2 v = m2[e]; mus[2 * e] = v.x; mus[2 * e + 1] = v.y; }
#pragma unroll
        for (int c = 0; c < NK; ++c) mus[39 + c] = mus_[39 + c];
    }
    constexpr int col0 = 16 * G;
    constexpr int ncols = C::NCOL - col0 < 16 ? C::NCOL - col0 : 16;
    const bool neg = lane & 1, act = (lane >> 1) < ncols;
    const int jl = act ? lane >> 1 : 0;                    // idle lanes (last group) shadow column 0 with a zero scale
    const double rs = rsqrt_fast(dv[col0 + jl]);           // L(:,j) = U(:,j) / sqrt(d_j): the scale rides on the sign
    const double sgn = act ? (neg ? -rs : rs) : 0.0;
    // U(r, j) = col[2 r]; rows above the diagonal hold zeros
    const double *col = Lf + (jl >> 2) * PS4 + ((jl & 3) >> 1) * C::PL + (jl & 1);
    auto Lc = [&](int r) -> double { return sgn * col[2 * r]; };
    double xpk[3], xqk[4], xpi[3], xqi[4], xf[NK];
    if (col0 <= 5) {   // column j perturbs rows >= j only: statek is untouched from the second group on
#pragma unroll
        for (int c = 0; c < 3; ++c) xpk[c] = mus[c] + Lc(c);
        const double v[3] = {Lc(3), Lc(4), Lc(5)};
        const double q[4] = {mus[3], mus[4], mus[5], mus[6]};
        double e[4];
        so3_exp(v, 1.0, e);
        quat_mul(q, e, xqk);
    } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) xpk[c] = mus[c];
#pragma unroll
        for (int c = 0; c < 4; ++c) xqk[c] = mus[3 + c];
    }
    if (col0 <= 29) {
#pragma unroll
        for (int c = 0; c < 3; ++c) xpi[c] = mus[26 + c] + Lc(24 + c);
        const double v[3] = {Lc(27), Lc(28), Lc(29)};
        const double q[4] = {mus[29], mus[30], mus[31], mus[32]};
        double e[4];
        so3_exp(v, 1.0, e);
        quat_mul(q, e, xqi);
    } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) xpi[c] = mus[26 + c];
#pragma unroll
        for (int c = 0; c < 4; ++c) xqi[c] = mus[29 + c];
    }
#pragma unroll
    for (int c = 0; c < NK; ++c) xf[c] = mus[39 + c] + Lc(36 + c);
    double z[NK];
    vo_model<NK>(xpk, xqk, xpi, xqi, xf, z);
    double e[NK];
#pragma unroll
    for (int c = 0; c < NK; ++c) {
        e[c] = act ? z[c] - s.z0[c] : 0.0;
        s.esum[c] += e[c];
        const double zp = __shfl_xor_sync(FULL, z[c], 1);
        if (act && !neg) Wg[c * 16 + jl] = (0.5 * rs) * (z[c] - zp);   // W'_j = (Z+ - Z-) / (2 sqrt d_j)
    }
#pragma unroll
    for (int r = 0; r < NK; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) s.eep[tri(r, c)] = fma(e[r], e[c], s.eep[tri(r, c)]);
    __syncwarp();
    // covXZ rows `lane` and `lane + 32` += U(row, 16 cols) W'(16 cols, :)
    const bool has1 = lane + 32 < C::NPAD;
#pragma unroll
    for (int qq = 0; qq < ncols / 4; ++qq) {
        const double *pa = Lf + qq * PS4 + 2 * lane, *pb = pa + C::PL;
        const double2 x01 = *reinterpret_cast<const double2 *>(pa), x23 = *reinterpret_cast<const double2 *>(pb);
        double2 y01 = make_double2(0.0, 0.0), y23 = y01;
        if (has1) { y01 = *reinterpret_cast<const double2 *>(pa + 64); y23 = *reinterpret_cast<const double2 *>(pb + 64); }
#pragma unroll
        for (int c = 0; c < NK; ++c) {
            const double2 w01 = *reinterpret_cast<const double2 *>(Wg + c * 16 + 4 * qq);
            const double2 w23 = *reinterpret_cast<const double2 *>(Wg + c * 16 + 4 * qq + 2);
            s.px0[c] = fma(x23.y, w23.y, fma(x23.x, w23.x, fma(x01.y, w01.y, fma(x01.x, w01.x, s.px0[c]))));
            s.px1[c] = fma(y23.y, w23.y, fma(y23.x, w23.x, fma(y01.y, w01.y, fma(y01.x, w01.x, s.px1[c]))));
        }
    }
}

// ---- one panel of 4 columns (compile-time index P), then the group pass after every 4th ---------------------------
template <class C, int P>
struct PanelStep {
    SLB_DEV static void run(UpdState<C> &s, const double *mus, double *Lf, double *Lraw, double *Wg, double *dv, int lane) {
        constexpr int NT = C::NT, JLAST = C::JLAST, PS4 = C::PS4;
        constexpr int K = 4 * P, J0 = K >> 3, H = (K >> 2) & 1, Q = P & 3;
        constexpr int I1 = (K + 4) >> 3;     // first tile row / column with live entries after this panel
        const int a_ = lane & 3, b_ = lane >> 2;
        double *Lq = Lf + Q * PS4;
        constexpr int PL = C::PL;
        // (1) the 16 lanes holding columns K..K+3 park them (rows of tile row J0 and below), a column pair per plane
        if ((a_ >> 1) == H) {
#pragma unroll
            for (int I = J0; I < NT; ++I)
                *reinterpret_cast<double2 *>(Lraw + (a_ & 1) * PL + 2 * (8 * I + b_)) = make_double2(s.c0[I][J0], s.c1[I][J0]);
        }
        __syncwarp();
        // (2) the 4 x 4 diagonal block, eliminated redundantly by every lane
        double m10, m20, m30, m21, m31, m32, r0, r1, r2, r3, d0, d1, d2, d3;
        {
            const double *A = Lraw + 2 * K, *Bp = A + PL;
            const double a00 = A[0];
            const double2 a1 = *reinterpret_cast<const double2 *>(A + 2);
            const double2 a2 = *reinterpret_cast<const double2 *>(A + 4);
            const double a22 = Bp[4];
            const double2 a3 = *reinterpret_cast<const double2 *>(A + 6), a3b = *reinterpret_cast<const double2 *>(Bp + 6);
            d0 = a00;
            r0 = rcp_fast(d0);
            m10 = a1.x * r0; m20 = a2.x * r0; m30 = a3.x * r0;
            d1 = fma(-a1.x, m10, a1.y);
            const double u21 = fma(-a2.x, m10, a2.y), u31 = fma(-a3.x, m10, a3.y);
            r1 = rcp_fast(d1);
            m21 = u21 * r1; m31 = u31 * r1;
            d2 = fma(-u21, m21, fma(-a2.x, m20, a22));
            const double u32 = fma(-u31, m21, fma(-a3.x, m20, a3b.x));
            r2 = rcp_fast(d2);
            m32 = u32 * r2;
            d3 = fma(-u32, m32, fma(-u31, m31, fma(-a3.x, m30, a3b.y)));
            r3 = rcp_fast(d3);
            s.ok = s.ok && (d0 > 0.0) && (d1 > 0.0) && (d2 > 0.0) && (d3 > 0.0);
        }
        if (lane == 0) {
            *reinterpret_cast<double2 *>(dv + K) = make_double2(d0, d1);
            *reinterpret_cast<double2 *>(dv + K + 2) = make_double2(d2, d3);
        }
        // (3) lane i finishes rows i and i + 32 of the panel: U(i, K + c); zeros above the diagonal / above the panel.
        //     Branch-free: rows above the panel read whatever the planes hold and select zeros.
        auto finish = [&](int i, bool have) {
            const double2 x01 = *reinterpret_cast<const double2 *>(Lraw + 2 * i);
            const double2 x23 = *reinterpret_cast<const double2 *>(Lraw + PL + 2 * i);
            const double u0 = x01.x;
            const double u1 = fma(-u0, m10, x01.y);
            const double u2 = fma(-u1, m21, fma(-u0, m20, x23.x));
            const double u3 = fma(-u2, m32, fma(-u1, m31, fma(-u0, m30, x23.y)));
            double2 u01, u23;
            u01.x = i >= K ? u0 : 0.0;
            u01.y = i >= K + 1 ? u1 : 0.0;
            u23.x = i >= K + 2 ? u2 : 0.0;
            u23.y = i >= K + 3 ? u3 : 0.0;
            if (have) {
                *reinterpret_cast<double2 *>(Lq + 2 * i) = u01;
                *reinterpret_cast<double2 *>(Lq + PL + 2 * i) = u23;
            }
        };
        if (K < 32) {
            finish(lane, true);
        } else {   // rows 0..31 lie above the panel
            *reinterpret_cast<double2 *>(Lq + 2 * lane) = make_double2(0.0, 0.0);
            *reinterpret_cast<double2 *>(Lq + PL + 2 * lane) = make_double2(0.0, 0.0);
        }
        if (C::NPAD > 32) finish(lane + 32 < C::NPAD ? lane + 32 : lane, lane + 32 < C::NPAD);
        __syncwarp();
        // (4) rank-4 update of the live tiles: A fragment = U(8I + b, K + a), B fragment = -U(8J + b, K + a) / d_{K+a}
        if (I1 <= JLAST) {
            double f[NT];
            const double *fp = Lq + (a_ >> 1) * PL + (a_ & 1) + 2 * b_;
#pragma unroll
            for (int I = I1; I < NT; ++I) f[I] = fp[16 * I];
            const double rlo = (a_ & 1) ? r1 : r0, rhi = (a_ & 1) ? r3 : r2;
            const double nr = -((a_ & 2) ? rhi : rlo);
#pragma unroll
            for (int J = I1; J <= JLAST; ++J) {
                const double g = f[J] * nr;
#pragma unroll
                for (int I = J; I < NT; ++I) dmma884(s.c0[I][J], s.c1[I][J], f[I], g);
            }
        }
        if (Q == 3 || P == C::NPAN - 1) {
            sigma_group<C, (P >> 2)>(s, mus, Lf, Wg, dv, lane);
            SLB_CTA_SYNC();
        }
        PanelStep<C, P + 1>::run(s, mus, Lf, Lraw, Wg, dv, lane);
    }
};
template <class C>
struct PanelStep<C, C::NPAN> {
    SLB_DEV static void run(UpdState<C> &, const double *, double *, double *, double *, double *, int) {}
};

// update of the resident record.  Returns status bits; `dirty` is set when Ps / mus changed.
template <int NK, int NL>
SLB_DEV int usckf_update_smem(double *Ps, double *mus, double *scr, const double *zs, const slb::FilterArgs &a, int lane,
                              bool &dirty) {
    typedef StepCfg<NK, NL> C;
    constexpr int N = C::N, NT = C::NT, NPAD = C::NPAD, JLAST = C::JLAST, KP = C::KP;
    const int a_ = lane & 3, b_ = lane >> 2;
    double *Lf = scr, *Lraw = scr + C::LF, *Ks = scr, *Cs = scr + NPAD * KP, *Wg = scr + C::SCR0, *dv = Wg + C::WGS,
           *dl = Cs + NPAD * KP;
    UpdState<C> s;
    s.ok = true;
    // ---- the lower triangle of Pk from the resident record into the accumulator tiles; diagonal tiles are filled
    //      symmetrically, rows / columns beyond N (N % 8 != 0) are an identity block --------------------------------
    auto elem = [&](int i, int j) -> double {
        if (i >= N || j >= N) return i == j ? 1.0 : 0.0;
        return i >= j ? Ps[tri(i, j)] : Ps[tri(j, i)];
    };
#pragma unroll
    for (int I = 0; I < NT; ++I)
#pragma unroll
        for (int J = 0; J <= JLAST; ++J)
            if (J <= I) {
                const int i = 8 * I + b_, j = 8 * J + 2 * a_;
                s.c0[I][J] = elem(i, j);
                s.c1[I][J] = elem(i, j + 1);
            }
#pragma unroll
    for (int c = 0; c < NK; ++c) s.px0[c] = s.px1[c] = s.esum[c] = 0.0;
#pragma unroll
    for (int e = 0; e < NK * (NK + 1) / 2; ++e) s.eep[e] = 0.0;
    {   // Z0 = h(mu)
        double pk[3], qk[4], pi[3], qi[4], ft[NK];
#pragma unroll
        for (int c = 0; c < 3; ++c) { pk[c] = mus[c]; pi[c] = mus[26 + c]; }
#pragma unroll
        for (int c = 0; c < 4; ++c) { qk[c] = mus[3 + c]; qi[c] = mus[29 + c]; }
#pragma unroll
        for (int c = 0; c < NK; ++c) ft[c] = mus[39 + c];
        vo_model<NK>(pk, qk, pi, qi, ft, s.z0);
    }
    // ---- factorisation (Eigen::LLT of :577) interleaved with the sigma points (:275-283) -------------------------
    PanelStep<C, 0>::run(s, mus, Lf, Lraw, Wg, dv, lane);
    if (!s.ok) return SLB_ST_CHOL_FAIL;

    // ---- mean / innovation covariance (:280-282) from the deviations e = Z - Z0 -----------------------------------
    constexpr double NS = (double)(2 * N + 1);
    double ebar[NK], S[NK * (NK + 1) / 2];
    if (NK == 3) {   // 3 + 6 sums in one reduce-scatter pass (the d_k slots are dead: every sigma point has been drawn)
        double v[12], o[12];
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = s.esum[c];
#pragma unroll
        for (int e = 0; e < 6; ++e) v[3 + e] = s.eep[e];
        v[9] = v[10] = v[11] = 0.0;
        warp_sum12(v, o, dv, lane);
#pragma unroll
        for (int c = 0; c < 3; ++c) ebar[c] = o[c] * (1.0 / NS);
#pragma unroll
        for (int r = 0; r < NK; ++r)
#pragma unroll
            for (int c = 0; c <= r; ++c)
                S[tri(r, c)] = 0.5 * fma(-NS * ebar[r], ebar[c], o[3 + tri(r, c)]) + __ldg(a.R + r * NK + c);
    } else {
#pragma unroll
        for (int c = 0; c < NK; ++c) ebar[c] = warp_sum(s.esum[c]) * (1.0 / NS);
#pragma unroll
        for (int r = 0; r < NK; ++r)
#pragma unroll
            for (int c = 0; c <= r; ++c)
                S[tri(r, c)] = 0.5 * fma(-NS * ebar[r], ebar[c], warp_sum(s.eep[tri(r, c)])) + __ldg(a.R + r * NK + c);
    }
    // ---- K = covXZ S^-1 (:286-288), innovation, Mahalanobis gate (:290-294) --------------------------------------
    double Si[NK * (NK + 1) / 2];
    const bool sok = sym_inverse<NK>(S, Si);
    auto SiAt = [&](int r, int c) { return r >= c ? Si[tri(r, c)] : Si[tri(c, r)]; };
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncwarp();   // z has landed; every lane is done reading Lf / Wg: K | covXZ overlay them
    double nu[NK], sn[NK], m2 = 0.0;
#pragma unroll
    for (int c = 0; c < NK; ++c) nu[c] = zs[c] - (s.z0[c] + ebar[c]);
#pragma unroll
    for (int r = 0; r < NK; ++r) {
        double t = 0.0;
#pragma unroll
        for (int c = 0; c < NK; ++c) t = fma(SiAt(r, c), nu[c], t);
        sn[r] = t;   // S^-1 nu
        m2 = fma(nu[r], t, m2);
    }
    if (!sok) return SLB_ST_CHOL_FAIL;
    if (!chi2_accept(m2, a.gate)) return SLB_ST_GATE_REJECT;
    auto finish_row = [&](const double *px, int i) {
        double dsum = 0.0;
#pragma unroll
        for (int c = 0; c < KP; ++c) {
            double k = 0.0;
            if (c < NK) {
#pragma unroll
                for (int p = 0; p < NK; ++p) k = fma(px[p], SiAt(p, c), k);
                dsum = fma(px[c], sn[c], dsum);
            }
            Ks[i * KP + c] = k;
            Cs[i * KP + c] = c < NK ? -px[c] : 0.0;   // K S = covXZ: Pk -= covXZ K^T
        }
        dl[i] = dsum;   // (K nu)_i
    };
    finish_row(s.px0, lane);
    if (lane + 32 < NPAD) finish_row(s.px1, lane + 32);
    __syncwarp();
    dirty = true;
    // ---- mu = mu [+] K nu (:299-301): lane b < 12 owns block b of the three States, then the feature scalars ------
    bool finite = true;
    if (lane < 12) {
        const int mbw = lane & 3;
        const int mqo = 13 * (lane >> 2) + (mbw == 0 ? 0 : mbw == 1 ? 3 : mbw == 2 ? 7 : 10);
        const double v[3] = {dl[3 * lane], dl[3 * lane + 1], dl[3 * lane + 2]};
        if (mbw == 1) {
            const double q[4] = {mus[mqo], mus[mqo + 1], mus[mqo + 2], mus[mqo + 3]};
            double e[4], o[4];
            so3_exp(v, 1.0, e);
            quat_mul(q, e, o);
#pragma unroll
            for (int c = 0; c < 4; ++c) { mus[mqo + c] = o[c]; finite = finite && isfinite(o[c]); }
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const double o = mus[mqo + c] + v[c];
                mus[mqo + c] = o;
                finite = finite && isfinite(o);
            }
        }
    }
    for (int f = lane; f < NK + NL; f += 32) {
        const double o = mus[39 + f] + dl[36 + f];
        mus[39 + f] = o;
        finite = finite && isfinite(o);
    }
    // ---- Pk -= K S K^T (:296) = Pk - covXZ K^T on the resident record: one DMMA k-step per tile and 4 columns of K
    {
        double fa[NT][KP / 4], fb[NT][KP / 4];
#pragma unroll
        for (int I = 0; I < NT; ++I)
#pragma unroll
            for (int kk = 0; kk < KP / 4; ++kk) {
                fa[I][kk] = Cs[(8 * I + b_) * KP + 4 * kk + a_];
                fb[I][kk] = Ks[(8 * I + b_) * KP + 4 * kk + a_];
            }
#pragma unroll
        for (int I = 0; I < NT; ++I)
#pragma unroll
            for (int J = 0; J <= I; ++J) {
                const int i = 8 * I + b_, j = 8 * J + 2 * a_;
                const bool w0 = i < N && j <= i, w1 = i < N && j + 1 <= i;
                double p0 = w0 ? Ps[tri(i, j)] : 0.0, p1 = w1 ? Ps[tri(i, j + 1)] : 0.0;
#pragma unroll
                for (int kk = 0; kk < KP / 4; ++kk) dmma884(p0, p1, fa[I][kk], fb[J][kk]);
                if (w0) Ps[tri(i, j)] = p0;
                if (w1) Ps[tri(i, j + 1)] = p1;
            }
    }
    return __all_sync(FULL, finite) ? 0 : SLB_ST_NONFINITE;
}

// predict of the resident record (statek_i: rows 24..35, mean at q-offset 26); see predict12_kernel for the scheme.
// Returns status bits; `dirty` is set when Ps / mus changed.
template <int PM, int NF>
SLB_DEV int usckf_predict_smem(double *Ps, double *mus, double *scr, const double *u, double dt, const double *Qg, int lane,
                               bool &dirty) {
    typedef LayState12 L;
    typedef CycCfg<12> C;
    constexpr int ROW0 = 24, MU0 = 26;
    double *Ls = scr, *D = Ls + PRED_LS, *W = D + 28 * PRED_DS, *Fk = W + 144, *rsd = Fk + 12 * PRED_DS, *slots = rsd + 16;
    auto PR = [&](int r, int c) -> double & { return Ps[tri(ROW0 + r, c)]; };        // row 24 + r, col c
    auto PF = [&](int fr, int c) -> double & { return Ps[tri(36 + fr, 24 + c)]; };   // feature row fr, col 24 + c
    const int a_ = lane & 3, b_ = lane >> 2;
    double mu[13];
#pragma unroll
    for (int c = 0; c < 13; ++c) mu[c] = mus[MU0 + c];
    ProcessModel<PM> f;
    f.prepare(u, dt);
    // ---- Pk_i = U D^-1 U^T (Eigen::LLT of :572-577 up to the per-column scale 1/sqrt(d_k)) ---------------------
    double T[C::RT][C::CT];
#pragma unroll
    for (int r = 0; r < C::RT; ++r)
#pragma unroll
        for (int c = 0; c < C::CT; ++c)
            if (C::exists(r, c)) {
                const int i = a_ + 4 * r, j = b_ + 8 * c;
                T[r][c] = (i < 12 && j <= i) ? PR(i, ROW0 + j) : 0.0;
            }
    bool ok = true;
    {
        const double x0 = __shfl_sync(FULL, T[0][0], 0);
        CholStep<C, 0>::run(T, Ls, a_, b_, ok, x0, -rcp_fast(x0));
    }
    if (!ok) return SLB_ST_CHOL_FAIL;
    __syncwarp();
    if (lane < 12) {
        double sq_, rs_;
        sqrt_rsqrt(Ls[C::cb(lane)], sq_, rs_);
        rsd[lane] = rs_;  // 1 / L_kk
    }
    __syncwarp();
    // ---- sigma point `lane` (Usckf.hpp:572-598), process model (:141) ---------------------------------
    const bool act = lane < 25;
    const int j = lane >= 1 ? (lane - 1) >> 1 : 0, jc = j < 12 ? j : 11;
    double Y[13];
    {
        const double sgn = (lane & 1) ? rsd[jc] : -rsd[jc];
        const int cj = C::cb(jc) - jc;
        double d[12], X[13];
#pragma unroll
        for (int r = 0; r < 12; ++r) d[r] = (act && lane >= 1 && r >= jc) ? sgn * Ls[cj + r] : 0.0;
        boxplus<L>(mu, d, 1.0, X);
        f.apply(X, Y);
    }
    // ---- manifold mean (:601-627) --------------------------------------------------------------------
    double ref[13];
#pragma unroll
    for (int c = 0; c < 13; ++c) ref[c] = bcast(Y[c], 0);
    int it = 0;
    double nrm2;
    do {
        double dd[12], md[12], nr[13];
        boxminus<L>(Y, ref, dd);
#pragma unroll
        for (int r = 0; r < 12; ++r) dd[r] = act ? dd[r] : 0.0;
        warp_sum12(dd, md, slots, lane);
        nrm2 = 0.0;
#pragma unroll
        for (int r = 0; r < 12; ++r) {
            md[r] *= (1.0 / 25.0);
            nrm2 += md[r] * md[r];
        }
        boxplus<L>(ref, md, 1.0, nr);
#pragma unroll
        for (int c = 0; c < 13; ++c) ref[c] = nr[c];
    } while (sqrt(nrm2) > 1e-6 && ++it < 10000);
    int st = (it >= 10000) ? SLB_ST_MEAN_NOCONV : 0;
    {   // deviations, one row per sigma point; rows 25..27 are the zero padding of the DMMA k-dimension
        double dY[12];
        boxminus<L>(Y, ref, dY);
        if (lane < 28) {
#pragma unroll
            for (int r = 0; r < 12; r += 2)   // (row stride 160 B: 16-byte stores, 2-way instead of 8-way bank conflicts)
                *reinterpret_cast<double2 *>(D + lane * PRED_DS + r) = act ? make_double2(dY[r], dY[r + 1]) : make_double2(0.0, 0.0);
        }
    }
    __syncwarp();
    dirty = true;
    // ---- Pk_i = 0.5 D^T D + Q (:178): three lower 8x8 tiles, k = 28 ---------------------------------------------
    {
        double c00[2] = {0.0, 0.0}, c10[2] = {0.0, 0.0}, c11[2] = {0.0, 0.0};
        const double *p0 = D + a_ * PRED_DS + b_, *p1 = p0 + 8;  // A[m][k] = D[k][m]: lane (row b_, k a_)
#pragma unroll
        for (int k0 = 0; k0 < 28; k0 += 4) {
            const double f0 = p0[k0 * PRED_DS], f1 = p1[k0 * PRED_DS];  // columns 12..15 of D only feed dropped outputs
            dmma884(c00[0], c00[1], f0, f0);
            dmma884(c10[0], c10[1], f1, f0);
            dmma884(c11[0], c11[1], f1, f1);
        }
        for (int e = lane; e < 144; e += 32) {
            const int jj = e / 12, c = e - jj * 12;
            W[e] = 0.5 * (D[(1 + 2 * jj) * PRED_DS + c] - D[(2 + 2 * jj) * PRED_DS + c]);
        }
        __syncwarp();   // the old block (rows 24..35, cols 24..35) was consumed by the factorisation: overwrite it
        auto put = [&](int r, int c, double v) {
            if (r < 12 && c <= r) PR(r, ROW0 + c) = 0.5 * v + __ldg(Qg + r * 12 + c);
        };
        const int r = b_, c = 2 * a_;
        put(r, c, c00[0]); put(r, c + 1, c00[1]);
        put(8 + r, c, c10[0]); put(8 + r, c + 1, c10[1]);
        put(8 + r, 8 + c, c11[0]); put(8 + r, 8 + c + 1, c11[1]);
    }
    // ---- Fk = W^T L^-1  <=>  L^T Fk^T = W: lane c back-substitutes column c (:154); L(p,r) = U(p,r) / sqrt(d_r) --
    if (lane < 12) {
        double x[12];
#pragma unroll
        for (int r = 11; r >= 0; --r) {
            double sacc = 0.0;
#pragma unroll
            for (int p = r + 1; p < 12; ++p) sacc = fma(Ls[C::cb(r) - r + p], x[p], sacc);
            const double rs = rsd[r];
            x[r] = (W[r * 12 + lane] - rs * sacc) * rs;
        }
#pragma unroll
        for (int r = 0; r < 12; ++r) Fk[lane * PRED_DS + r] = x[r];
    }
    __syncwarp();
    // ---- cross-covariances with the clones (:191-208): rows 24..35 x cols 0..23  <- Fk * old: 2 x 3 tiles, k = 12 ----
    double oc[2][3][2];
    {
        const int r0 = min(b_, 11), r1 = min(8 + b_, 11);
#pragma unroll
        for (int I = 0; I < 2; ++I)
#pragma unroll
            for (int J = 0; J < 3; ++J) oc[I][J][0] = oc[I][J][1] = 0.0;
#pragma unroll
        for (int k0 = 0; k0 < 12; k0 += 4) {
            const double fa0 = Fk[r0 * PRED_DS + k0 + a_], fa1 = Fk[r1 * PRED_DS + k0 + a_];
#pragma unroll
            for (int J = 0; J < 3; ++J) {
                const double bv = PR(k0 + a_, 8 * J + b_);  // B[k][n] = old P(24 + k, n)
                dmma884(oc[0][J][0], oc[0][J][1], fa0, bv);
                dmma884(oc[1][J][0], oc[1][J][1], fa1, bv);
            }
        }
    }
    // ---- and with the features (:217-235): feature rows x cols 24..35  <- old * Fk^T: ceil(nf/8) x 2 tiles, k = 12 ----
    constexpr int NFT = (NF + 7) / 8;
    double of[NFT > 0 ? NFT : 1][2][2];
    {
        const int c0 = min(b_, 11), c1 = min(8 + b_, 11);
#pragma unroll
        for (int I = 0; I < NFT; ++I) {
            of[I][0][0] = of[I][0][1] = of[I][1][0] = of[I][1][1] = 0.0;
            const int fr = min(8 * I + b_, NF - 1);
#pragma unroll
            for (int k0 = 0; k0 < 12; k0 += 4) {
                const double av = PF(fr, k0 + a_);
                dmma884(of[I][0][0], of[I][0][1], av, Fk[c0 * PRED_DS + k0 + a_]);  // B[k][n] = Fk[n][k]
                dmma884(of[I][1][0], of[I][1][1], av, Fk[c1 * PRED_DS + k0 + a_]);
            }
        }
    }
    __syncwarp();
#pragma unroll
    for (int I = 0; I < 2; ++I)
#pragma unroll
        for (int J = 0; J < 3; ++J) {
            const int r = 8 * I + b_, c = 8 * J + 2 * a_;
            if (r < 12) { PR(r, c) = oc[I][J][0]; PR(r, c + 1) = oc[I][J][1]; }
        }
#pragma unroll
    for (int I = 0; I < NFT; ++I) {
#pragma unroll
        for (int J = 0; J < 2; ++J) {
            const int fr = 8 * I + b_, c = 8 * J + 2 * a_;
            if (fr < NF && c < 12) { PF(fr, c) = of[I][J][0]; PF(fr, c + 1) = of[I][J][1]; }
        }
    }
    bool finite = true;
#pragma unroll
    for (int c = 0; c < 13; ++c) {
        finite = finite && isfinite(ref[c]);
        if (lane == c) mus[MU0 + c] = ref[c];
    }
    if (!finite) st |= SLB_ST_NONFINITE;
    __syncwarp();
    return st;
}

// One instance per warp: record in (TMA bulk load), predict and / or update in shared memory, record out (TMA bulk
// store).  PRED && UPD is the fused step (slb_usckf_step, the *_step_host paths); UPD alone is slb_usckf_update.
template <int PM, int NK, int NL, bool PRED, bool UPD, int WPB, int MINB>
__global__ void __launch_bounds__(WPB * 32, MINB) usckf_step_kernel(slb::FilterArgs a) {
    typedef StepCfg<NK, NL> C;
    extern __shared__ __align__(128) double smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int inst = blockIdx.x * WPB + w;
    if (inst >= a.B) return;
    double *Ps = smem + (size_t)w * C::SM, *mus = Ps + C::PSTR, *scr = mus + C::QS, *zs = scr + C::SCR;
    uint64_t *bar = reinterpret_cast<uint64_t *>(zs + C::ZS);
    double *Pg = a.P + (size_t)inst * a.pstride;
    double *mug = a.mu + (size_t)inst * a.qstride;
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, (C::PSTR + C::QS) * 8);
        bulk_g2s(Ps, Pg, C::PSTR * 8, bar);
        bulk_g2s(mus, mug, C::QS * 8, bar);
        // the record a warp of a later wave of CTAs will want: pull it into L2 now (one wave = SMs x resident warps)
        if (a.prefetch > 0 && inst + a.prefetch < a.B) bulk_prefetch_l2(Pg + (size_t)a.prefetch * a.pstride, C::PSTR * 8);
    }
    // the measurement and the control input are fetched now: with the zero-copy *_step_host they live in mapped host
    // memory and their PCIe latency must not sit in the middle of the step
    if (UPD && lane < NK) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n cp.async.commit_group;\n" ::"r"(saddr(zs + lane)),
                     "l"(a.z + (size_t)inst * NK + lane)
                     : "memory");
    }
    double u[ProcessModel<PM>::NU];
    if (PRED) {
#pragma unroll
        for (int c = 0; c < ProcessModel<PM>::NU; ++c) u[c] = a.u[(size_t)inst * ProcessModel<PM>::NU + c];
    }
    __syncwarp();
    mbar_wait(bar, 0);
    SLB_CTA_SYNC();
    int st = 0;
    bool dirty = false;
    if (PRED) st |= usckf_predict_smem<PM, NK + NL>(Ps, mus, scr, u, a.dt, a.Q, lane, dirty);
    if (PRED) SLB_CTA_SYNC();
    if (UPD) st |= usckf_update_smem<NK, NL>(Ps, mus, scr, zs, a, lane, dirty);
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");   // (an update that bailed out early has not waited for z)
    // ---- record out ----------------------------------------------------------------------------------------------
    if (dirty) {
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) bulk_s2g(Pg, Ps, C::PSTR * 8);
        for (int e = lane; e < C::QD; e += 32) mug[e] = mus[e];
    }
    // optional instance-major copy of the posterior mean (mapped host memory in the zero-copy *_step_host): whatever
    // the step decided (accepted, gated, factorisation failed) the resident record holds the posterior
    if (a.mu_out) {
        __syncwarp();
        for (int e = lane; e < a.out_len; e += 32) a.mu_out[(size_t)inst * a.out_len + e] = mus[a.out_off + e];
    }
    if (st && lane == 0) a.status[inst] |= st;
    if (dirty && lane == 0) bulk_store_wait();
}

}  // namespace slbd
