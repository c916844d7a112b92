// How fast does the register-resident pair recurrence run WITHOUT the exchange?
// 20 unrolled orders of nbr_pair_order per "step", d[K+1] faked from t[K] in registers.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../heyoka.py_b200/csrc/hy_nbody_reg.cuh"
using namespace hy;
template <int K> struct Ord {
    static __device__ __forceinline__ void run(double (&d0)[20], double (&d1)[20], double (&d2)[20], double (&r2)[20],
                                               double (&c)[20], double &inv, double k0, double k1, double k2, double &s)
    {
        double t0, t1, t2;
        nbr_pair_order<double, K, 20>(d0, d1, d2, r2, c, inv, k0, k1, k2, t0, t1, t2);
        s += t0 + t1 + t2;
        if constexpr (K + 1 < 20) Ord<K + 1>::run(d0, d1, d2, r2, c, inv, t0 * 1e-3 + k1, t1 * 1e-3 + k2, t2 * 1e-3 + k0, s);
    }
};
__global__ void __launch_bounds__(256, 1) k(double *out, int iters)
{
    double d0[20], d1[20], d2[20], r2[20], c[20], inv = 0, s = 0;
    double x = 1.0 + threadIdx.x * 1e-6;
    for (int it = 0; it < iters; ++it) {
        Ord<0>::run(d0, d1, d2, r2, c, inv, x, x * 0.5, x * 0.25, s);
        x += s * 1e-30;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *out; cudaMalloc(&out, 8 * 1024 * p.multiProcessorCount);
    for (int threads : {128, 256}) {
        const int iters = 4000;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            k<<<p.multiProcessorCount, threads>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("threads/SM %d: %.3f ms, %.1f clk per step per warp @1.965GHz\n", threads, ms, ms * 1e-3 * 1.965e9 / iters);
    }
    return 0;
}
