// slb_internal.h -- shared declarations between the translation units of libslb.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/slb.h"

constexpr int SLB_NXS = 3;  // internal streams of the *_step_host pipelines

struct slb_batch_s {
    slb_config cfg;
    int N;        // tangent dimension of the full filter state
    int QD;       // q-vector length
    int NP;       // packed lower-triangle length N(N+1)/2
    int B;        // instances
    int stride;   // UKF: SoA stride (B rounded up to 32); USCKF/MSCKF: unused
    int pstride;  // USCKF/MSCKF: doubles per instance record of P (NP rounded up to 16 -> 128 B)
    int qstride;  // USCKF/MSCKF: doubles per instance record of mu (QD rounded up to 2 -> 16 B)
    double *mu;   // UKF: [QD][stride] SoA.  USCKF/MSCKF: [B][qstride]
    double *P;    // UKF: [NP][stride] SoA packed lower.  USCKF/MSCKF: [B][pstride] packed lower
    int32_t *status;
    int32_t *outliers;
    // staging for the *_host entry points and upload/download
    double *stage;        // device scratch, stage_bytes long
    size_t stage_bytes;
    double *shared_small; // device copy of Q / R / params passed from host pointers
    int64_t *counts_dev;  // 4 counters for slb_status
    int32_t *misc_dev;    // 16 ints of per-launch scratch flags (e.g. "R is diagonal" for the MSCKF EKF update)
    cudaStream_t xs[SLB_NXS];  // chunk ring of the *_step_host entry points
    // the *_step_host pipeline captured as a CUDA graph (replayed while the caller keeps passing the same pinned
    // buffers and parameters: one cudaGraphLaunch instead of ~30 stream calls per step)
    cudaStream_t cap_stream;
    cudaGraphExec_t step_exec;
    int step_kernels;          // kernel launches inside the captured step
    int step_key_i[8];
    double step_key_dt;
    const void *step_key_p[6];
    cudaEvent_t ev_start, ev_done[SLB_NXS];
    int out_off, out_len;      // slb_set_output_slice: part of the q-vector the *_step_host entry points copy back
    // zero-copy *_step_host: host copy of the Q | R last uploaded to shared_small (the upload is skipped while the caller
    // keeps passing the same values: two tiny DMA operations per step are 10 % of a UKFoM step)
    double qr_cache[144 + 81];
    int qr_nq, qr_m;           // 0 = nothing cached
};

namespace slb {

int set_error(int code, const char *what, cudaError_t ce = cudaSuccess);
void count_launch(int n = 1);

#define SLB_CUDA(call)                                                        \
    do {                                                                      \
        cudaError_t e_ = (call);                                              \
        if (e_ != cudaSuccess) return slb::set_error(SLB_ERR_CUDA, #call, e_); \
    } while (0)

// Every entry point that takes a handle runs on the handle's device and leaves the caller's current device as it
// found it (several handles on several devices may be driven from one thread).
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};

struct FilterArgs {
    double *mu;
    double *P;
    int32_t *status;
    int32_t *outliers;
    int B;
    int stride, pstride, qstride;
    const double *u;
    double dt;
    const double *Q;
    const double *z;
    const double *R;
    const double *params;
    int m;
    int gate;
    int nk, nl, k;
    int32_t *misc;
    int prefetch;     // USCKF step: L2-prefetch the record of instance (i + prefetch); 0 = off
    double *mu_out;   // optional instance-major copy of the posterior means (may be mapped host memory), B x out_len
    int out_off, out_len;  // the slice [out_off, out_off + out_len) of the q-vector that goes to mu_out
};

// slb_ukf.cu
int launch_ukf(int layout, int pm, int mm, bool predict, bool update, const FilterArgs &a, cudaStream_t s);
// slb_usckf.cu
int launch_usckf(int pm, int mm, bool predict, bool update, const FilterArgs &a, cudaStream_t s);
bool usckf_shape_supported(int nk, int nl);
int launch_usckf_clone(int mode, const FilterArgs &a, cudaStream_t s);
int launch_usckf_set_measurement(int mode, const FilterArgs &a, cudaStream_t s);
// slb_msckf.cu
int launch_msckf_predict(int pm, const FilterArgs &a, cudaStream_t s);
int launch_msckf_update(int mm, const FilterArgs &a, cudaStream_t s);
int launch_msckf_update_ekf(int mm, const FilterArgs &a, cudaStream_t s);
// slb_check.cu
int launch_check_sigma_points(const slb_batch_s *h, int32_t *flags, double *diff, cudaStream_t s);
// slb_fusion.cu
int launch_fusion(int d, int64_t n, int op, const double *x1, const double *C1, const double *x2,
                  const double *C2, double *xo, double *Co, cudaStream_t s);

}  // namespace slb
