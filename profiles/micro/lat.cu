// lat.cu -- dependent-issue latencies (cycles) of the instructions the sigma-point kernels' serial chains are made
// of, one warp on one SM:   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu && ./lat
#include <cstdio>
#include <cuda_runtime.h>

#define REP 256
template <int ILP>
__global__ void k_dfma(double *out, long long *cyc, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = a + i + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < REP / 16; ++r) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], b, a);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_ffma(float *out, long long *cyc, float a, float b) {
    float x = a + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < REP / 16; ++r) {
#pragma unroll
        for (int u = 0; u < 16; ++u) x = fmaf(x, b, a);
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_shfl(double *out, long long *cyc, double a) {
    double x = a + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < REP / 16; ++r) {
#pragma unroll
        for (int u = 0; u < 16; ++u) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_shfl32(float *out, long long *cyc, float a) {
    float x = a + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < REP / 16; ++r) {
#pragma unroll
        for (int u = 0; u < 16; ++u) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lds(double *out, long long *cyc) {
    __shared__ int nxt[1024];
    for (int i = threadIdx.x; i < 1024; i += 32) nxt[i] = (i + 32) & 1023;
    __syncwarp();
    int p = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < REP / 16; ++r) {
#pragma unroll
        for (int u = 0; u < 16; ++u) p = nxt[p];
    }
    long long t1 = clock64();
    out[threadIdx.x] = p;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// STS -> __syncwarp -> LDS round trip through shared memory (what a column publication costs)
__global__ void k_sts_lds(double *out, long long *cyc, double a) {
    __shared__ double buf[64];
    double x = a + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < REP / 16; ++r) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            buf[threadIdx.x] = x;
            __syncwarp();
            x = buf[(threadIdx.x + 1) & 31];
            __syncwarp();
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_rsq(double *out, long long *cyc, double a) {
    double x = a + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < REP / 16; ++r) {
#pragma unroll
        for (int u = 0; u < 16; ++u) asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(x));
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_rcp(double *out, long long *cyc, double a) {
    double x = a + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < REP / 16; ++r) {
#pragma unroll
        for (int u = 0; u < 16; ++u) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(x));
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int ILP>
__global__ void k_dmma(double *out, long long *cyc, double a, double b) {
    double c[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = i; c[i][1] = threadIdx.x; }
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < REP / 16; ++r) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int i = 0; i < ILP; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(c[i][0]), "+d"(c[i][1])
                             : "d"(a), "d"(b));
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// DMMA whose A operand depends on the previous result (the panel -> trailing-update dependency)
__global__ void k_dmma_a(double *out, long long *cyc, double a, double b) {
    double c0 = a, c1 = b;
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < REP / 16; ++r) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%0}, {%2}, {%0,%1};\n"
                         : "+d"(c0), "+d"(c1)
                         : "d"(b));
    }
    long long t1 = clock64();
    out[threadIdx.x] = c0 + c1;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
    double *out;
    long long *cyc, h;
    cudaMalloc(&out, 8 * 64);
    cudaMalloc(&cyc, 8);
#define RUN(name, launch, n)                                                    \
    launch;                                                                     \
    launch;                                                                     \
    cudaDeviceSynchronize();                                                    \
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                             \
    printf("%-44s %7.2f cycles per op\n", name, (double)h / (double)(n));
    RUN("DFMA dependent chain", (k_dfma<1><<<1, 32>>>(out, cyc, 1.0, 0.5)), REP);
    RUN("DFMA 2 independent chains (per DFMA)", (k_dfma<2><<<1, 32>>>(out, cyc, 1.0, 0.5)), REP * 2);
    RUN("DFMA 4 independent chains (per DFMA)", (k_dfma<4><<<1, 32>>>(out, cyc, 1.0, 0.5)), REP * 4);
    RUN("DFMA 8 independent chains (per DFMA)", (k_dfma<8><<<1, 32>>>(out, cyc, 1.0, 0.5)), REP * 8);
    RUN("FFMA dependent chain", (k_ffma<<<1, 32>>>((float *)out, cyc, 1.0f, 0.5f)), REP);
    RUN("SHFL.IDX f64 (2 x 32-bit) dependent", (k_shfl<<<1, 32>>>(out, cyc, 1.0)), REP);
    RUN("SHFL.IDX 32-bit dependent", (k_shfl32<<<1, 32>>>((float *)out, cyc, 1.0f)), REP);
    RUN("LDS dependent (pointer chase)", (k_lds<<<1, 32>>>(out, cyc)), REP);
    RUN("STS + syncwarp + LDS + syncwarp round trip", (k_sts_lds<<<1, 32>>>(out, cyc, 1.0)), REP);
    RUN("MUFU.RSQ64H (rsqrt.approx.f64) dependent", (k_rsq<<<1, 32>>>(out, cyc, 1.5)), REP);
    RUN("MUFU.RCP64H (rcp.approx.f64) dependent", (k_rcp<<<1, 32>>>(out, cyc, 1.5)), REP);
    RUN("DMMA m8n8k4 dependent on C", (k_dmma<1><<<1, 32>>>(out, cyc, 1.0, 0.5)), REP);
    RUN("DMMA 2 independent accumulators (per DMMA)", (k_dmma<2><<<1, 32>>>(out, cyc, 1.0, 0.5)), REP * 2);
    RUN("DMMA 4 independent accumulators (per DMMA)", (k_dmma<4><<<1, 32>>>(out, cyc, 1.0, 0.5)), REP * 4);
    RUN("DMMA dependent on A", (k_dmma_a<<<1, 32>>>(out, cyc, 1.0, 0.5)), REP);
    printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
