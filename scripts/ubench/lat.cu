// micro-benchmark: dependent-chain latencies on the target GPU (developer tool)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double a, double b, int n)
{
    __shared__ double sm[1024];
    __shared__ unsigned int msk;
    int lane = threadIdx.x;
    sm[lane] = a + lane; if (lane == 0) msk = 0;
    __syncthreads();
    double x = a + lane * 1e-9, y = b;
    long long t0, t1;
    // DFMA chain
    t0 = clock64();
    #pragma unroll 16
    for (int i = 0; i < n; ++i) x = fma(x, y, y);
    t1 = clock64(); if (lane == 0 && blockIdx.x==0) cyc[0] = t1 - t0;
    // DADD chain
    t0 = clock64();
    #pragma unroll 16
    for (int i = 0; i < n; ++i) x = x + y;
    t1 = clock64(); if (lane == 0&& blockIdx.x==0) cyc[1] = t1 - t0;
    // DMUL chain
    t0 = clock64();
    #pragma unroll 16
    for (int i = 0; i < n; ++i) x = x * y;
    t1 = clock64(); if (lane == 0&& blockIdx.x==0) cyc[2] = t1 - t0;
    // sqrt chain
    x = fabs(x) + 2.0;
    t0 = clock64();
    #pragma unroll 4
    for (int i = 0; i < n; ++i) x = sqrt(x + 3.0);
    t1 = clock64(); if (lane == 0&& blockIdx.x==0) cyc[3] = t1 - t0;
    // LDS dependent chain
    int idx = lane;
    t0 = clock64();
    #pragma unroll 16
    for (int i = 0; i < n; ++i) idx = ((int)sm[idx & 1023]) & 1023;
    t1 = clock64(); if (lane == 0&& blockIdx.x==0) cyc[4] = t1 - t0;
    // syncthreads
    t0 = clock64();
    for (int i = 0; i < n; ++i) __syncthreads();
    t1 = clock64(); if (lane == 0&& blockIdx.x==0) cyc[5] = t1 - t0;
    // ballot chain
    unsigned m = lane;
    t0 = clock64();
    #pragma unroll 16
    for (int i = 0; i < n; ++i) m = __ballot_sync(0xffffffffu, (m >> (lane & 7)) & 1) + i;
    t1 = clock64(); if (lane == 0&& blockIdx.x==0) cyc[6] = t1 - t0;
    // smem atomicOr
    t0 = clock64();
    for (int i = 0; i < n; ++i) if (lane == 0) atomicOr(&msk, 1u << (i & 31));
    t1 = clock64(); if (lane == 0&& blockIdx.x==0) cyc[7] = t1 - t0;
    // FFMA chain
    float f = (float)a;
    t0 = clock64();
    #pragma unroll 16
    for (int i = 0; i < n; ++i) f = fmaf(f, 1.0001f, 0.5f);
    t1 = clock64(); if (lane == 0&& blockIdx.x==0) cyc[8] = t1 - t0;
    // independent DFMA throughput: 8 chains
    double z[8]; for (int j = 0; j < 8; ++j) z[j] = a + j;
    t0 = clock64();
    for (int i = 0; i < n; ++i) { 
        #pragma unroll
        for (int j = 0; j < 8; ++j) z[j] = fma(z[j], y, y); }
    t1 = clock64(); if (lane == 0&& blockIdx.x==0) cyc[9] = t1 - t0;
    double s = x + idx + m + f; for (int j = 0; j < 8; ++j) s += z[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + msk;
}
int main()
{
    double* out; long long* cyc; int n = 4096;
    cudaMalloc(&out, 8 * 1024 * 1024); cudaMalloc(&cyc, 16 * 8);
    const char* names[] = {"DFMA dep", "DADD dep", "DMUL dep", "dsqrt(+DADD) dep", "LDS.64 dep", "syncthreads", "ballot dep", "smem atomicOr", "FFMA dep", "8x indep DFMA (per 8)"};
    for (int threads : {32, 256, 1024})
    {
        k<<<1, threads>>>(out, cyc, 1.0000001, 0.9999999, n);
        k<<<1, threads>>>(out, cyc, 1.0000001, 0.9999999, n);
        long long h[16]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("threads=%d:", threads);
        for (int i = 0; i < 10; ++i) printf("  %s=%.1f", names[i], (double)h[i] / n);
        printf("\n");
    }
    // full chip fp64 throughput
    k<<<148 * 4, 256>>>(out, cyc, 1.0000001, 0.9999999, n);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<<<148 * 8, 256>>>(out, cyc, 1.0000001, 0.9999999, n); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); printf("full-chip kernel %.3f ms (%s)\n", ms, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
