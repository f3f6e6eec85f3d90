// slb_predict12.cuh -- warp-per-instance sigma-point predict of a 12-DOF State block inside a larger
// packed covariance record; shared by Usckf::predict (Usckf.hpp:113-244: block statek_i at rows 24..35,
// cross-covariances propagated) and Msckf::predict (Msckf.hpp:97-189: block statek at rows 0..11, cross
// blocks left stale, quirk Q5).
#pragma once
#include "slb_internal.h"
#include "slb_models.cuh"

namespace slbd {

constexpr unsigned FULL = 0xffffffffu;
SLB_DEV double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
SLB_DEV double bcast(double v, int src) { return __shfl_sync(FULL, v, src); }

// ---- TMA bulk copy helpers (global -> shared, completion on an mbarrier) ---------------------------
SLB_DEV unsigned saddr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
SLB_DEV void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(saddr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
SLB_DEV void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(saddr(bar)), "r"(bytes) : "memory");
}
SLB_DEV void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(saddr(dst)),
                 "l"(src), "r"(bytes), "r"(saddr(bar))
                 : "memory");
}
SLB_DEV void mbar_wait(uint64_t *bar, unsigned parity) {
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done)
            : "r"(saddr(bar)), "r"(parity)
            : "memory");
    }
}

// =====================================================================================================
// predict
// =====================================================================================================
constexpr int PRED_PR = 366;           // packed rows 24..35: T(36) - T(24) (USCKF; MSCKF uses 78 of it)
constexpr int PRED_PF = 28 * 12;       // feature-row segments, nk + nl <= 28
constexpr int PRED_SM = PRED_PR + PRED_PF + 78 + 25 * 13 + 144 + 144 + 3;  // doubles per warp (even: 16-B aligned
                                                                           // warps for the bulk copy; last slot = mbarrier)
SLB_DEV void pred_cp_async8(double *smem_dst, const double *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(saddr(smem_dst)), "l"(gsrc) : "memory");
}
SLB_DEV void pred_cp_async_wait_all() { asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;\n" ::: "memory"); }

// ROW0: first row of the 12x12 block; MU0: q-vector offset of the block's mean; CROSS: propagate the
// cross-covariances of the block's rows/columns with the rest of the state (USCKF) or not (MSCKF).
template <int PM, int WPB, int ROW0, int MU0, bool CROSS>
__global__ void __launch_bounds__(WPB * 32, 2) predict12_kernel(slb::FilterArgs a) {
    typedef LayState12 L;
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int inst = blockIdx.x * WPB + w;
    if (inst >= a.B) return;
    double *Pr = smem + (size_t)w * PRED_SM, *Pf = Pr + PRED_PR, *Ls = Pf + PRED_PF, *D = Ls + 78, *W = D + 25 * 13,
           *Fk = W + 144;
    const int nf = CROSS ? a.nk + a.nl : 0;
    constexpr int T0 = ROW0 * (ROW0 + 1) / 2, T1 = (ROW0 + 12) * (ROW0 + 13) / 2, SPAN = T1 - T0;
    static_assert(SPAN <= PRED_PR, "row span");
    double *Pg = a.P + (size_t)inst * a.pstride;
    double *mug = a.mu + (size_t)inst * a.qstride;
    auto PR = [&](int r, int c) -> double & { return Pr[tri(ROW0 + r, c) - T0]; };  // row ROW0+r, col c

    // The block's rows are one contiguous span of the packed record: one TMA bulk copy (UBLKCP) brings it in,
    // the 12-wide feature-row segments follow as LDGSTS; the mean and the control input load meanwhile.
    static_assert((T0 * 8) % 16 == 0 && (SPAN * 8) % 16 == 0 && PRED_SM % 2 == 0, "bulk copy needs 16-byte alignment");
    uint64_t *bar = reinterpret_cast<uint64_t *>(Fk + 144);
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, SPAN * 8);
        bulk_g2s(Pr, Pg + T0, SPAN * 8, bar);
    }
    for (int e = lane; e < nf * 12; e += 32) {
        const int r = e / 12, c = e - r * 12;
        pred_cp_async8(Pf + e, Pg + tri(36 + r, 24 + c));
    }
    double mu[13];
#pragma unroll
    for (int c = 0; c < 13; ++c) mu[c] = mug[MU0 + c];
    ProcessModel<PM> f;
    {
        double u[ProcessModel<PM>::NU];
#pragma unroll
        for (int c = 0; c < ProcessModel<PM>::NU; ++c) u[c] = a.u[(size_t)inst * ProcessModel<PM>::NU + c];
        f.prepare(u, a.dt);
    }
    pred_cp_async_wait_all();
    __syncwarp();
    mbar_wait(bar, 0);

    // ---- Cholesky of Pk_i (12x12): lane l < 12 owns row l; rows are broadcast with shuffles -----------
    double row[12];
#pragma unroll
    for (int p = 0; p < 12; ++p) row[p] = (lane < 12 && p <= lane) ? PR(lane, ROW0 + p) : 0.0;
    bool ok = true;
    double invd[12];  // 1 / L_kk (every lane computes the pivot: no IEEE division, whose slow path the zero
                      // numerators of the idle lanes would take)
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        double s = row[k];
#pragma unroll
        for (int p = 0; p < 12; ++p)
            if (p < k) s = fma(-row[p], bcast(row[p], k), s);
        const double x = bcast(s, k);
        ok = ok && (x > 0.0);
        double sx;
        sqrt_rsqrt(x, sx, invd[k]);
        row[k] = (lane == k) ? sx : s * invd[k];
    }
    if (!ok) {
        if (lane == 0) a.status[inst] |= SLB_ST_CHOL_FAIL;
        return;
    }
    if (lane < 12) {
#pragma unroll
        for (int p = 0; p < 12; ++p)
            if (p <= lane) Ls[tri(lane, p)] = row[p];
    }
    __syncwarp();

    // ---- sigma point `lane` (Usckf.hpp:572-598), process model (:141) ---------------------------------
    const bool act = lane < 25;
    const int j = (lane - 1) >> 1;
    const double sgn = (lane & 1) ? 1.0 : -1.0;
    double Y[13];
    {
        double d[12], X[13];
#pragma unroll
        for (int r = 0; r < 12; ++r) d[r] = (act && lane >= 1 && r >= j) ? sgn * Ls[tri(r, j < 0 ? 0 : (j > r ? r : j))] : 0.0;
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
        nrm2 = 0.0;
#pragma unroll
        for (int r = 0; r < 12; ++r) {
            md[r] = warp_sum(act ? dd[r] : 0.0) * (1.0 / 25.0);  // not "/ 25.0": zero numerators take div.rn's slow path
            nrm2 += md[r] * md[r];
        }
        boxplus<L>(ref, md, 1.0, nr);
#pragma unroll
        for (int c = 0; c < 13; ++c) ref[c] = nr[c];
    } while (sqrt(nrm2) > 1e-6 && ++it < 10000);
    int st = (it >= 10000) ? SLB_ST_MEAN_NOCONV : 0;

    {
        double dY[12];
        boxminus<L>(Y, ref, dY);
        if (act) {
#pragma unroll
            for (int r = 0; r < 12; ++r) D[lane * 13 + r] = dY[r];
        }
    }
    __syncwarp();
    // ---- Pk_i = cov + Q (:178) and W = 0.5 (dY+ - dY-) -------------------------------------------------
#pragma unroll
    for (int e0 = 0; e0 < 96; e0 += 32) {
        const int e = e0 + lane;
        if (e < 78) {
            int r = (int)((sqrtf(8.0f * e + 1.0f) - 1.0f) * 0.5f);  // packed index -> (r, c), exact for e < 78
            r += (tri(r + 1, 0) <= e) - (tri(r, 0) > e);
            const int c = e - tri(r, 0);
            double s = 0.0;
#pragma unroll
            for (int t = 0; t < 25; ++t) s = fma(D[t * 13 + r], D[t * 13 + c], s);
            PR(r, ROW0 + c) = 0.5 * s + __ldg(a.Q + r * 12 + c);
        }
    }
    if (CROSS) {
        for (int e = lane; e < 144; e += 32) {
            const int jj = e / 12, c = e - jj * 12;
            W[e] = 0.5 * (D[(1 + 2 * jj) * 13 + c] - D[(2 + 2 * jj) * 13 + c]);
        }
        __syncwarp();
        // ---- Fk = W^T L^-1  <=>  L^T Fk^T = W: lane c back-substitutes column c (:154) --------------------
        if (lane < 12) {
            double x[12];
    #pragma unroll
            for (int r = 11; r >= 0; --r) {
                double s = W[r * 12 + lane];
    #pragma unroll
                for (int p = r + 1; p < 12; ++p) s -= Ls[tri(p, r)] * x[p];
                x[r] = s * invd[r];
            }
    #pragma unroll
            for (int r = 0; r < 12; ++r) Fk[lane * 12 + r] = x[r];
        }
        __syncwarp();
        // ---- cross-covariances with the clones (:191-208): rows 24..35 x cols 0..23  <- Fk * old ----------
        double out[9];
    #pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int o = lane + 32 * t, r = o / 24, c = o - r * 24;
            double s = 0.0;
    #pragma unroll
            for (int p = 0; p < 12; ++p) s += Fk[r * 12 + p] * PR(p, c);
            out[t] = s;
        }
        // ---- and with the features (:217-235): feature rows x cols 24..35  <- old * Fk^T -------------------
        double of[11];
    #pragma unroll
        for (int t = 0; t < 11; ++t) {
            const int o = lane + 32 * t;
            double s = 0.0;
            if (o < nf * 12) {
                const int r = o / 12, c = o - r * 12;
    #pragma unroll
                for (int p = 0; p < 12; ++p) s += Pf[r * 12 + p] * Fk[c * 12 + p];
            }
            of[t] = s;
        }
        __syncwarp();
    #pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int o = lane + 32 * t, r = o / 24, c = o - r * 24;
            PR(r, c) = out[t];
        }
    #pragma unroll
        for (int t = 0; t < 11; ++t) {
            const int o = lane + 32 * t;
            if (o < nf * 12) Pf[o] = of[t];
        }
        __syncwarp();
    } else {
        __syncwarp();
    }
    // ---- write back --------------------------------------------------------------------------------------
    for (int e = lane; e < SPAN; e += 32) Pg[T0 + e] = Pr[e];
    for (int e = lane; e < nf * 12; e += 32) {
        const int r = e / 12, c = e - r * 12;
        Pg[tri(36 + r, 24 + c)] = Pf[e];
    }
    bool finite = true;
#pragma unroll
    for (int c = 0; c < 13; ++c) {
        finite = finite && isfinite(ref[c]);
        if (lane == c) mug[MU0 + c] = ref[c];
    }
    if (!finite) st |= SLB_ST_NONFINITE;
    if (st && lane == 0) a.status[inst] |= st;
}

}  // namespace slbd
