// DFMA throughput with realistic operand patterns (developer microbenchmark).
//   mode 0: a = fma(a, b, c) with loop-invariant b, c (the classic peak kernel)
//   mode 1: acc_i = fma(x_j, y_k, acc_i), 6 chains, x/y: 16 distinct registers each, rotating
// run with 256 threads/CTA, 1 CTA/SM (2 warps per scheduler) and with 1024 threads.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void k(double *out, int iters, double seed)
{
    double x[16], y[16], a[6];
    for (int i = 0; i < 16; ++i) { x[i] = seed + i * 1e-3 + threadIdx.x * 1e-9; y[i] = 1.0 - i * 1e-4; }
    for (int i = 0; i < 6; ++i) a[i] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                if (MODE == 0) a[c] = fma(a[c], x[0], y[0]);
                else a[c] = fma(x[(u + c) & 15], y[(u * 3 + c * 5) & 15], a[c]);
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a[0] + a[1] + a[2] + a[3] + a[4] + a[5];
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *out; cudaMalloc(&out, 8 * 2048 * p.multiProcessorCount);
    for (int threads : {128, 256, 512, 1024})
        for (int mode = 0; mode < 2; ++mode) {
            const int iters = 20000;
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) k<0><<<p.multiProcessorCount, threads>>>(out, iters, 1.0);
                else k<1><<<p.multiProcessorCount, threads>>>(out, iters, 1.0);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double fmas = (double)iters * 96 * threads * p.multiProcessorCount;
            printf("threads/SM %4d mode %d: %.2f TFLOP/s, %.2f DFMA lanes/clk/SM @1.965GHz\n", threads, mode,
                   2 * fmas / (ms * 1e-3) / 1e12, fmas / (ms * 1e-3) / p.multiProcessorCount / 1.965e9);
        }
    return 0;
}
