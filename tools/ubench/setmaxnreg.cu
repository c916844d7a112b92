// Cost of a setmaxnreg.dec / setmaxnreg.inc pair per warpgroup (developer microbenchmark).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(384, 1) k(long long *out, int iters)
{
    const int wg = threadIdx.x / 128;
    asm volatile("bar.sync %0, 128;" ::"r"(wg + 1));
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (wg < 2) { // two warpgroups trade, the third stays low
            asm volatile("bar.sync %0, 128;" ::"r"(wg + 1));
            asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
            asm volatile("bar.sync %0, 128;" ::"r"(wg + 1));
            asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        }
    }
    long long t1 = clock64();
    if ((threadIdx.x & 127) == 0) out[blockIdx.x * 3 + wg] = t1 - t0;
}
int main()
{
    long long *out; cudaMallocManaged(&out, 8 * 3 * 148);
    const int iters = 2000;
    k<<<148, 384>>>(out, iters); cudaDeviceSynchronize();
    k<<<148, 384>>>(out, iters); cudaDeviceSynchronize();
    printf("clocks per inc+dec pair (with 2 bar.sync): wg0 %.1f wg1 %.1f (wg2 idle %.1f) err=%s\n", out[0] / (double)iters,
           out[1] / (double)iters, out[2] / (double)iters, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
