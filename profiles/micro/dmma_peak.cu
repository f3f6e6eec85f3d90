// dmma_peak.cu -- is the FP64 tensor path (mma.sync m8n8k4 DMMA) faster than the FP64 SIMT pipe on B200?
// north_star allows DMMA tiles "only where ncu shows they beat the FP64 SIMT pipe": this measures both rates
// with register-resident operands.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_peak dmma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) dfma_kernel(double *sink, int iters, double a, double b) {
    double x[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = a + k + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = fma(x[k], b, a);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += x[k];
    if (s == 12345.678) sink[0] = s;
}

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(256) dmma_kernel(double *sink, int iters, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int k = 0; k < 8; ++k) { c[k][0] = k; c[k][1] = threadIdx.x; }
    const double av = a + (threadIdx.x & 3), bv = b + (threadIdx.x >> 2);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) dmma(c[k][0], c[k][1], av, bv);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += c[k][0] + c[k][1];
    if (s == 12345.678) sink[0] = s;
}

int main() {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *sink;
    cudaMalloc(&sink, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 4096, grid = sms * 8;
    for (int which = 0; which < 2; ++which) {
        double best = 0;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            if (which == 0) dfma_kernel<<<grid, 256>>>(sink, iters, 1.0000001, 0.9999999);
            else dmma_kernel<<<grid, 256>>>(sink, iters, 1.0000001, 0.9999999);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            // DFMA: 16 FMA per thread per iter; DMMA: 8 mma per warp per iter, 8*8*4 FMA each
            const double fl = which == 0 ? 2.0 * 16 * iters * 256.0 * grid : 2.0 * 8 * 256 * iters * (256.0 / 32) * grid;
            const double tf = fl / (ms * 1e-3) * 1e-12;
            if (rep > 0 && tf > best) best = tf;
        }
        printf("%s %.2f TFLOP/s\n", which == 0 ? "DFMA (SIMT FP64)" : "DMMA m8n8k4 (tensor FP64)", best);
    }
    printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
