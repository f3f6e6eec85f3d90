// slb_msckf.cu -- localization::Msckf<MultiState<State,SensorState>,State> (src/filters/Msckf.hpp):
// predict (warp per instance, shared with the USCKF predict) and the UKF-flavoured update with the
// per-feature chi-square gate, one filter instance per CTA.
//
// update (Msckf.hpp:220-277, removeOutliers :723-754, applyDelta :659-666) at BASELINE config 3
// (10 clones -> N = 72, 145 sigma points, 50 features -> m = 100) keeps everything of one instance in
// shared memory (~224 KB: one CTA per SM, persistent over the batch):
//   A  5056 doubles : P -> L (Cholesky in place) ... later S -> Ls ... later P_new -> L2 -> P_out
//   B 14645 doubles : Z (145 x 101) ... later compacted S'/Pxz' (outliers) ... later X / D (145 x 83)
//   C  7272 doubles : W -> covXZ (in-place TRMM) -> Y = covXZ Ls^-T (in-place TRSM)
// Algebra: with S = Ls Ls^T and Y = covXZ Ls^-T the reference's  K = covXZ S^-1, Pk -= K S K^T,
// delta = K nu  become  Pk -= Y Y^T,  delta = Y (Ls^-1 nu)  -- same maths as :257-263 without forming
// the explicit inverse (quirk Q9).  covXZ = L W with W_j = 0.5 (Z+_j - Z-_j), as in the other kernels.
#include "slb_predict12.cuh"

namespace slbd {

constexpr int MS_T = 256;          // threads per CTA
constexpr int MS_NMAX = 72, MS_MMAX = 100, MS_NSMAX = 145, MS_QMAX = 83;
constexpr int MS_A = 5056;                       // >= 100*101/2 and >= 72*73/2
constexpr int MS_ZS = MS_MMAX + 1;               // odd row stride of Z and of covXZ / Y
constexpr int MS_B = MS_NSMAX * MS_ZS;           // 14645
constexpr int MS_C = MS_NMAX * MS_ZS;            // 7272
constexpr int MS_D = 768;                        // small vectors
constexpr int MS_SMEM_DOUBLES = MS_A + MS_B + MS_C + MS_D;
static_assert(MS_SMEM_DOUBLES * 8 <= 227 * 1024, "MSCKF update working set exceeds shared memory");
static_assert(MS_NSMAX * MS_QMAX <= MS_B, "sigma points do not fit region B");

// In-place Cholesky of a packed lower matrix in shared memory, right-looking, all threads of the CTA.
// ok_flag (shared int) is cleared on a non-positive pivot.
SLB_DEV void chol_smem(double *A, int n, int *ok_flag) {
    const int tid = threadIdx.x;
    const int ti = tid >> 4, tj = tid & 15;
    for (int k = 0; k < n; ++k) {
        if (tid == 0) {
            const double x = A[tri(k, k)];
            if (!(x > 0.0)) *ok_flag = 0;
            A[tri(k, k)] = sqrt(x);
        }
        __syncthreads();
        const double inv = 1.0 / A[tri(k, k)];
        for (int i = k + 1 + tid; i < n; i += MS_T) A[tri(i, k)] *= inv;
        __syncthreads();
        for (int i = k + 1 + ti; i < n; i += 16) {
            const double lik = A[tri(i, k)];
            for (int j = k + 1 + tj; j <= i; j += 16) A[tri(i, j)] -= lik * A[tri(j, k)];
        }
        __syncthreads();
    }
}

// Multi-state q-vector: statek (pos quat velo angvelo) then k sensor poses (pos quat).
// Block b of the tangent space: 0..3 = statek blocks, 4+2c / 5+2c = pos / orient of clone c.
SLB_DEV int ms_qoff(int b) { return b < 4 ? (b == 0 ? 0 : b == 1 ? 3 : b == 2 ? 7 : 10) : 13 + 7 * ((b - 4) >> 1) + (((b - 4) & 1) ? 3 : 0); }
SLB_DEV bool ms_so3(int b) { return b < 4 ? b == 1 : ((b - 4) & 1); }

__global__ void __launch_bounds__(MS_T, 1) msckf_update_kernel(slb::FilterArgs a) {
    extern __shared__ __align__(16) double sm[];
    double *RA = sm, *RB = RA + MS_A, *RC = RB + MS_B, *RD = RC + MS_C;
    double *mu = RD, *zbar = mu + 84, *nu = zbar + 100, *wv = nu + 100, *dl = wv + 100, *acc = dl + 72, *ref = acc + 72;
    int *kept = reinterpret_cast<int *>(ref + 84);  // 100 ints
    int *flags = kept + 100;                        // [0] chol ok, [1] kept count, [2] outliers, [3] loop, [4] iters
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = a.k, N = 12 + 6 * k, QD = 13 + 7 * k, NS = 2 * N + 1, M = a.m, NF = M / 2, NB = 4 + 2 * k;
    const int NP = N * (N + 1) / 2;

    for (int inst = blockIdx.x; inst < a.B; inst += gridDim.x) {
        double *Pg = a.P + (size_t)inst * a.pstride;
        double *mug = a.mu + (size_t)inst * a.qstride;
        const double *zg = a.z + (size_t)inst * M;
        __syncthreads();
        for (int e = tid; e < NP; e += MS_T) RA[e] = Pg[e];
        for (int e = tid; e < QD; e += MS_T) mu[e] = mug[e];
        if (tid == 0) { flags[0] = 1; flags[1] = M; flags[2] = 0; }
        __syncthreads();
        // ---- L = chol(Pk) (:229 -> :412) -------------------------------------------------------------
        chol_smem(RA, N, flags);
        int st = 0;
        if (!flags[0]) {
            if (tid == 0) a.status[inst] |= SLB_ST_CHOL_FAIL;
            continue;
        }
        // ---- sigma points through h (:229-232), thread per sigma point ---------------------------------
        // h = SLB_MM_MSCKF_REPROJ: feature f is landmark f seen from clone f % k (statek is not observed)
        if (tid < NS) {
            const int s = tid, j = s >= 1 ? (s - 1) >> 1 : 0;
            const double sgn = (s & 1) ? 1.0 : -1.0;
            double *Zr = RB + s * MS_ZS;
            for (int c = 0; c < k; ++c) {
                const int r0 = 12 + 6 * c;
                double d[6];
#pragma unroll
                for (int r = 0; r < 6; ++r) d[r] = (s >= 1 && r0 + r >= j) ? sgn * RA[tri(r0 + r, j)] : 0.0;
                const double *mp = mu + 13 + 7 * c;
                const double p[3] = {mp[0] + d[0], mp[1] + d[1], mp[2] + d[2]};
                double e[4], q[4];
                so3_exp(d + 3, 1.0, e);
                quat_mul(mp + 3, e, q);
                for (int f = c; f < NF; f += k) {
                    const double dv[3] = {__ldg(a.params + 3 * f) - p[0], __ldg(a.params + 3 * f + 1) - p[1],
                                          __ldg(a.params + 3 * f + 2) - p[2]};
                    double pc[3];
                    quat_rotate_inv(q, dv, pc);
                    Zr[2 * f] = pc[0] / pc[2];
                    Zr[2 * f + 1] = pc[1] / pc[2];
                }
            }
        }
        __syncthreads();
        // ---- mean_z, innovation (:234-236) -------------------------------------------------------------
        if (tid < M) {
            double s = 0.0;
            for (int t = 0; t < NS; ++t) s += RB[t * MS_ZS + tid];
            const double zb = s * (1.0 / (double)NS);
            zbar[tid] = zb;
            nu[tid] = zg[tid] - zb;
        }
        // ---- W_j = 0.5 (Z+_j - Z-_j) into region C ---------------------------------------------------
        for (int e = tid; e < N * M; e += MS_T) {
            const int j = e / M, c = e - j * M;
            RC[j * MS_ZS + c] = 0.5 * (RB[(1 + 2 * j) * MS_ZS + c] - RB[(2 + 2 * j) * MS_ZS + c]);
        }
        __syncthreads();
        // centre Z for the covariance
        for (int e = tid; e < NS * M; e += MS_T) {
            const int s = e / M, c = e - s * M;
            RB[s * MS_ZS + c] -= zbar[c];
        }
        // ---- covXZ = L W (:239 -> :635-657): in-place TRMM on region C, 8 rows per pass, bottom up --------
        for (int i0 = ((N - 1) / 8) * 8; i0 >= 0; i0 -= 8) {
            const int rows = min(8, N - i0);
            double out[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int o = tid + MS_T * t;
                double s = 0.0;
                if (o < rows * M) {
                    const int i = i0 + o / M, c = o % M;
                    for (int j = 0; j <= i; ++j) s += RA[tri(i, j)] * RC[j * MS_ZS + c];
                }
                out[t] = s;
            }
            __syncthreads();
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int o = tid + MS_T * t;
                if (o < rows * M) RC[(i0 + o / M) * MS_ZS + o % M] = out[t];
            }
            __syncthreads();
        }
        // ---- S = 0.5 Zc^T Zc + R (:238) into region A (L is dead), 5x5 register tiles -------------------
        {
            const int nt = (M + 4) / 5;
            int tr = 0, rem = tid;
            while (rem > tr) { rem -= tr + 1; ++tr; }  // tid -> (tr, tc) lower tile index
            const int tc = rem;
            if (tr < nt) {
                double accS[5][5];
#pragma unroll
                for (int x = 0; x < 5; ++x)
#pragma unroll
                    for (int y = 0; y < 5; ++y) accS[x][y] = 0.0;
                const int r0 = 5 * tr, c0 = 5 * tc;
                for (int s = 0; s < NS; ++s) {
                    const double *Zr = RB + s * MS_ZS;
                    double zr[5], zc[5];
#pragma unroll
                    for (int x = 0; x < 5; ++x) { zr[x] = (r0 + x < M) ? Zr[r0 + x] : 0.0; zc[x] = (c0 + x < M) ? Zr[c0 + x] : 0.0; }
#pragma unroll
                    for (int x = 0; x < 5; ++x)
#pragma unroll
                        for (int y = 0; y < 5; ++y) accS[x][y] += zr[x] * zc[y];
                }
#pragma unroll
                for (int x = 0; x < 5; ++x)
#pragma unroll
                    for (int y = 0; y < 5; ++y) {
                        const int r = r0 + x, c = c0 + y;
                        if (r < M && c <= r) RA[tri(r, c)] = 0.5 * accS[x][y] + __ldg(a.R + r * M + c);
                    }
            }
        }
        __syncthreads();
        // ---- removeOutliers (:241 -> :723-754), index quirk Q6 reproduced ---------------------------------
        if (a.gate && tid == 0) {
            int len = M, out = 0, i = 0;
            for (int e = 0; e < M; ++e) kept[e] = e;
            while (i < len / 2) {
                const int ia = kept[2 * i], ib = kept[2 * i + 1];
                const double s00 = RA[tri(ia, ia)], s11 = RA[tri(ib, ib)], s10 = ib > ia ? RA[tri(ib, ia)] : RA[tri(ia, ib)];
                const double det = s00 * s11 - s10 * s10;
                const double v0 = nu[ia], v1 = nu[ib];
                const double m2 = (v0 * (s11 * v0 - s10 * v1) + v1 * (s00 * v1 - s10 * v0)) / det;
                if (!(m2 < 5.99)) {
                    // removeRow(2i); removeRow(2i+1) -- the second index is NOT re-based (Q6)
                    for (int pass = 0; pass < 2; ++pass) {
                        const int pos = 2 * i + pass, num = len - 1;
                        if (pos < num)
                            for (int e = pos; e < num; ++e) kept[e] = kept[e + 1];
                        len = num;
                    }
                    ++out;
                } else {
                    ++i;
                }
            }
            flags[1] = len;
            flags[2] = out;
        }
        __syncthreads();
        const int mk = flags[1];
        if (tid == 0) a.outliers[inst] = flags[2];
        if (mk <= 0) continue;  // :250 nothing left to update with
        double *Sp = RA, *Xz = RC;  // S' (packed, stride implicit) and covXZ' (row stride MS_ZS)
        if (mk < M) {
            // compact S -> region B (Z is dead), covXZ in place (kept[] is increasing, kept[p] >= p)
            for (int e = tid; e < mk * (mk + 1) / 2; e += MS_T) {
                int r = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
                while (tri(r, 0) > e) --r;
                while (tri(r + 1, 0) <= e) ++r;
                const int c = e - tri(r, 0);
                RB[e] = RA[tri(kept[r], kept[c])];
            }
            for (int i = tid; i < N; i += MS_T)
                for (int p = 0; p < mk; ++p) RC[i * MS_ZS + p] = RC[i * MS_ZS + kept[p]];
            if (tid < mk) wv[tid] = nu[kept[tid]];
            __syncthreads();
            if (tid < mk) nu[tid] = wv[tid];
            Sp = RB;
            __syncthreads();
        }
        // ---- Ls = chol(S') -------------------------------------------------------------------------------
        chol_smem(Sp, mk, flags);
        if (!flags[0]) {
            if (tid == 0) a.status[inst] |= SLB_ST_CHOL_FAIL;
            continue;
        }
        // ---- Y = covXZ' Ls^-T: right-looking TRSM, 4 columns per pass ---------------------------------
        for (int q0 = 0; q0 < mk; q0 += 4) {
            const int qb = min(4, mk - q0);
            if (tid < N) {
                double *row = Xz + tid * MS_ZS;
                for (int q = q0; q < q0 + qb; ++q) {
                    double s = row[q];
                    for (int p = q0; p < q; ++p) s -= Sp[tri(q, p)] * row[p];
                    row[q] = s * rcp_fast(Sp[tri(q, q)]);
                }
            }
            __syncthreads();
            const int rest = mk - q0 - qb;
            for (int e = tid; e < N * rest; e += MS_T) {
                const int i = e / rest, c = q0 + qb + e % rest;
                double s = Xz[i * MS_ZS + c];
                for (int p = q0; p < q0 + qb; ++p) s -= Xz[i * MS_ZS + p] * Sp[tri(c, p)];
                Xz[i * MS_ZS + c] = s;
            }
            __syncthreads();
        }
        // ---- w = Ls^-1 nu (warp 0), delta = Y w ------------------------------------------------------
        if (warp == 0) {
            for (int q = 0; q < mk; ++q) {
                const double wq = nu[q] * rcp_fast(Sp[tri(q, q)]);
                __syncwarp();
                if (lane == 0) wv[q] = wq;
                for (int c = q + 1 + lane; c < mk; c += 32) nu[c] -= Sp[tri(c, q)] * wq;
                __syncwarp();
            }
        }
        __syncthreads();
        for (int i = warp; i < N; i += MS_T / 32) {
            double s = 0.0;
            for (int q = lane; q < mk; q += 32) s += Xz[i * MS_ZS + q] * wv[q];
            s = warp_sum(s);
            if (lane == 0) dl[i] = s;
        }
        // ---- P_new = Pk - Y Y^T (:262) into region A from the HBM record, 4x4 register tiles ------------
        __syncthreads();
        {
            const int nt = (N + 3) / 4;
            int tr = 0, rem = tid;
            while (rem > tr) { rem -= tr + 1; ++tr; }
            const int tc = rem;
            if (tr < nt) {
                double ac[4][4];
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) ac[x][y] = 0.0;
                const int r0 = 4 * tr, c0 = 4 * tc;
                for (int q = 0; q < mk; ++q) {
                    double yr[4], yc[4];
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        yr[x] = (r0 + x < N) ? Xz[(r0 + x) * MS_ZS + q] : 0.0;
                        yc[x] = (c0 + x < N) ? Xz[(c0 + x) * MS_ZS + q] : 0.0;
                    }
#pragma unroll
                    for (int x = 0; x < 4; ++x)
#pragma unroll
                        for (int y = 0; y < 4; ++y) ac[x][y] += yr[x] * yc[y];
                }
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) {
                        const int r = r0 + x, c = c0 + y;
                        if (r < N && c <= r) RA[tri(r, c)] = Pg[tri(r, c)] - ac[x][y];
                    }
            }
        }
        __syncthreads();
        // ---- applyDelta(K nu) (:263 -> :659-666): L2 = chol(P_new), X = mu [+] (delta +- L2 e_j) ----------
        chol_smem(RA, N, flags);
        if (!flags[0]) {
            if (tid == 0) a.status[inst] |= SLB_ST_CHOL_FAIL;
            continue;
        }
        if (tid < NS) {
            const int s = tid, j = s >= 1 ? (s - 1) >> 1 : 0;
            const double sgn = (s & 1) ? 1.0 : -1.0;
            double *X = RB + s * MS_QMAX;
            for (int b = 0; b < NB; ++b) {
                double d[3];
#pragma unroll
                for (int r = 0; r < 3; ++r) d[r] = dl[3 * b + r] + ((s >= 1 && 3 * b + r >= j) ? sgn * RA[tri(3 * b + r, j)] : 0.0);
                const int qo = ms_qoff(b);
                if (ms_so3(b)) {
                    double e[4], q[4];
                    so3_exp(d, 1.0, e);
                    quat_mul(mu + qo, e, q);
                    X[qo] = q[0]; X[qo + 1] = q[1]; X[qo + 2] = q[2]; X[qo + 3] = q[3];
                } else {
                    X[qo] = mu[qo] + d[0]; X[qo + 1] = mu[qo + 1] + d[1]; X[qo + 2] = mu[qo + 2] + d[2];
                }
            }
        }
        __syncthreads();
        // manifold mean (:499-525): ref = X0; do { d = mean(Xi [-] ref); ref [+]= d } while (|d| > 1e-6 ...)
        for (int e = tid; e < QD; e += MS_T) ref[e] = RB[e];
        if (tid == 0) flags[4] = 0;
        __syncthreads();
        // per-warp partial sums go to region C (Y is dead) and are added in warp order: the result of an
        // instance must not depend on scheduling (bitwise reproducible across batch sizes and GPUs)
        double *part = RC;
        const int nwa = (NS + 31) / 32;
        while (true) {
            for (int b = 0; b < NB; ++b) {
                double d[3] = {0.0, 0.0, 0.0};
                if (tid < NS) {
                    const double *X = RB + tid * MS_QMAX;
                    const int qo = ms_qoff(b);
                    if (ms_so3(b)) {
                        double r[4];
                        quat_cmul(ref + qo, X + qo, r);
                        so3_log(r, d);
                    } else {
                        d[0] = X[qo] - ref[qo]; d[1] = X[qo + 1] - ref[qo + 1]; d[2] = X[qo + 2] - ref[qo + 2];
                    }
                }
                if (warp < nwa) {
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        const double s = warp_sum(d[r]);
                        if (lane == 0) part[warp * MS_NMAX + 3 * b + r] = s;
                    }
                }
            }
            __syncthreads();
            if (tid < N) {
                double s = 0.0;
                for (int w2 = 0; w2 < nwa; ++w2) s += part[w2 * MS_NMAX + tid];
                acc[tid] = s;
            }
            __syncthreads();
            if (tid < NB) {
                const int b = tid, qo = ms_qoff(b);
                const double wn = 1.0 / (double)NS;
                const double d[3] = {acc[3 * b] * wn, acc[3 * b + 1] * wn, acc[3 * b + 2] * wn};
                dl[3 * b] = d[0]; dl[3 * b + 1] = d[1]; dl[3 * b + 2] = d[2];
                if (ms_so3(b)) {
                    double e[4], q[4];
                    so3_exp(d, 1.0, e);
                    quat_mul(ref + qo, e, q);
                    ref[qo] = q[0]; ref[qo + 1] = q[1]; ref[qo + 2] = q[2]; ref[qo + 3] = q[3];
                } else {
                    ref[qo] += d[0]; ref[qo + 1] += d[1]; ref[qo + 2] += d[2];
                }
            }
            __syncthreads();
            if (tid == 0) {
                double n2 = 0.0;
                for (int e = 0; e < N; ++e) n2 += dl[e] * dl[e];
                flags[3] = (sqrt(n2) > 1e-6 && ++flags[4] < 10000) ? 1 : 0;
            }
            __syncthreads();
            if (!flags[3]) break;
        }
        if (flags[4] >= 10000) st |= SLB_ST_MEAN_NOCONV;
        // deviations d_s = X_s [-] mean, in place (72 <= 83 slots per sigma point)
        if (tid < NS) {
            double *X = RB + tid * MS_QMAX;
            for (int b = 0; b < NB; ++b) {
                const int qo = ms_qoff(b);
                double d[3];
                if (ms_so3(b)) {
                    double r[4];
                    quat_cmul(ref + qo, X + qo, r);
                    so3_log(r, d);
                } else {
                    d[0] = X[qo] - ref[qo]; d[1] = X[qo + 1] - ref[qo + 1]; d[2] = X[qo + 2] - ref[qo + 2];
                }
                // 3b <= qoff(b): writing the tangent block never overtakes the q-blocks still to be read
                X[3 * b] = d[0]; X[3 * b + 1] = d[1]; X[3 * b + 2] = d[2];
            }
        }
        __syncthreads();
        // ---- Pk = 0.5 sum d d^T (:574-589) straight to the HBM record, 4x4 register tiles ----------------
        {
            const int nt = (N + 3) / 4;
            int tr = 0, rem = tid;
            while (rem > tr) { rem -= tr + 1; ++tr; }
            const int tc = rem;
            if (tr < nt) {
                double ac[4][4];
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) ac[x][y] = 0.0;
                const int r0 = 4 * tr, c0 = 4 * tc;
                for (int s = 0; s < NS; ++s) {
                    const double *Dr = RB + s * MS_QMAX;
                    double yr[4], yc[4];
#pragma unroll
                    for (int x = 0; x < 4; ++x) { yr[x] = (r0 + x < N) ? Dr[r0 + x] : 0.0; yc[x] = (c0 + x < N) ? Dr[c0 + x] : 0.0; }
#pragma unroll
                    for (int x = 0; x < 4; ++x)
#pragma unroll
                        for (int y = 0; y < 4; ++y) ac[x][y] += yr[x] * yc[y];
                }
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) {
                        const int r = r0 + x, c = c0 + y;
                        if (r < N && c <= r) Pg[tri(r, c)] = 0.5 * ac[x][y];
                    }
            }
        }
        bool finite = true;
        for (int e = tid; e < QD; e += MS_T) {
            mug[e] = ref[e];
            finite = finite && isfinite(ref[e]);
        }
        if (!finite) st |= SLB_ST_NONFINITE;
        if (st) atomicOr(a.status + inst, st);
    }
}

}  // namespace slbd

namespace slb {

template <int PM>
static int launch_ms_predict_t(const FilterArgs &a, cudaStream_t s) {
    constexpr int WPB = 8;
    constexpr size_t smem = (size_t)WPB * slbd::PRED_SM * sizeof(double);
    auto kern = slbd::predict12_kernel<PM, WPB, 0, 0, false>;
    SLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(a.B + WPB - 1) / WPB, WPB * 32, smem, s>>>(a);
    count_launch();
    SLB_CUDA(cudaGetLastError());
    return SLB_OK;
}

int launch_msckf_predict(int pm, const FilterArgs &a, cudaStream_t s) {
    if (pm == SLB_PM_MSCKF_DELTAPOSE) return launch_ms_predict_t<SLB_PM_MSCKF_DELTAPOSE>(a, s);
    if (pm == SLB_PM_USCKF_TEST) return launch_ms_predict_t<SLB_PM_USCKF_TEST>(a, s);
    return set_error(SLB_ERR_INVALID, "msckf: unsupported process model");
}

int launch_msckf_update(int mm, const FilterArgs &a, cudaStream_t s) {
    if (mm != SLB_MM_MSCKF_REPROJ) return set_error(SLB_ERR_INVALID, "msckf: unsupported measurement model");
    if (a.k < 1 || a.k > 10 || a.m > slbd::MS_MMAX || a.m < 2)
        return set_error(SLB_ERR_INVALID, "msckf update: needs 1..10 clones and 2 <= m <= 100");
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t smem = (size_t)slbd::MS_SMEM_DOUBLES * sizeof(double);
    SLB_CUDA(cudaFuncSetAttribute(slbd::msckf_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = a.B < sms ? a.B : sms;  // one persistent CTA per SM
    slbd::msckf_update_kernel<<<grid, slbd::MS_T, smem, s>>>(a);
    count_launch();
    SLB_CUDA(cudaGetLastError());
    return SLB_OK;
}

}  // namespace slb
