// Latency microbenchmarks: dependent DFMA chain, dependent LDS chain, LDS.64 throughput per warp.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dfma(double *out, long long *cyc, int n)
{
    double a = threadIdx.x * 1e-9 + 1.0, b = 1.0000001, c = 1e-7;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) a = fma(a, b, c);
    }
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_dfma4(double *out, long long *cyc, int n)
{
    double a0 = threadIdx.x * 1e-9 + 1.0, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b = 1.0000001, c = 1e-7;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) { a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c); }
    }
    long long t1 = clock64();
    out[threadIdx.x] = a0 + a1 + a2 + a3;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_ffma(float *out, long long *cyc, int n)
{
    float a = threadIdx.x * 1e-9f + 1.0f, b = 1.0000001f, c = 1e-7f;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) a = fmaf(a, b, c);
    }
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lds(int *out, long long *cyc, int n)
{
    __shared__ int idx[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i + 33) & 1023;
    __syncthreads();
    int p = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) p = idx[p];
    }
    long long t1 = clock64();
    out[threadIdx.x] = p;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lds64_tp(double *out, long long *cyc, int n)
{
    __shared__ double buf[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) buf[i] = i;
    __syncthreads();
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    const double *p = buf + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 16; u += 4) { s0 += p[u * 32]; s1 += p[(u + 1) * 32]; s2 += p[(u + 2) * 32]; s3 += p[(u + 3) * 32]; }
    }
    long long t1 = clock64();
    out[threadIdx.x] = s0 + s1 + s2 + s3;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main()
{
    double *d; long long *c; cudaMalloc(&d, 1 << 16); cudaMalloc(&c, 64);
    long long h; int n = 1000;
    for (int threads : {32, 128, 512}) {
        k_dfma<<<1, threads>>>(d, c, n); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("threads %d: dependent DFMA: %.2f clk/op\n", threads, (double)h / (16.0 * n));
        k_dfma4<<<1, threads>>>(d, c, n); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("threads %d: 4-chain DFMA: %.2f clk/op\n", threads, (double)h / (16.0 * n));
        k_ffma<<<1, threads>>>((float *)d, c, n); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("threads %d: dependent FFMA: %.2f clk/op\n", threads, (double)h / (16.0 * n));
        k_lds<<<1, threads>>>((int *)d, c, n); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("threads %d: dependent LDS: %.2f clk/op\n", threads, (double)h / (16.0 * n));
        k_lds64_tp<<<1, threads>>>(d, c, n); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("threads %d: LDS.64+DADD stream (4 chains): %.2f clk/op\n", threads, (double)h / (16.0 * n));
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
