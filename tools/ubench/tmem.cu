// Can TMEM serve as per-thread scratch?  tcgen05.ld/st.32x32b: lane i of warp w
// reads/writes TMEM lane 32*(w%4)+i at a warp-uniform column.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void tm_st(uint32_t taddr, double v)
{
    uint32_t lo = (uint32_t)__double2loint(v), hi = (uint32_t)__double2hiint(v);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(lo), "r"(hi) : "memory");
}
__device__ __forceinline__ double tm_ld(uint32_t taddr)
{
    uint32_t lo, hi;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(taddr) : "memory");
    return __hiloint2double((int)hi, (int)lo);
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tm_ld16(uint32_t taddr, double (&v)[8])
{
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr)
                 : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
}
__global__ void k16(double *out, long long *cyc, int *err)
{
    __shared__ uint32_t s_base;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = s_base;
    const uint32_t nslots = blockDim.x / 128;
    const uint32_t ncol = 512 / (nslots ? nslots : 1);
    const uint32_t tb = base + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * ncol;
    for (uint32_t c = 0; c < ncol; c += 2) tm_st(tb + c, (double)(threadIdx.x * 1000 + c));
    tm_wait_st();
    __syncthreads();
    int bad = 0;
    // unaligned start columns (multiples of 2): verify
    for (uint32_t c = 0; c + 16 <= ncol; c += 6) {
        double v[8];
        tm_ld16(tb + c, v);
        tm_wait_ld();
        for (int i = 0; i < 8; ++i)
            if (v[i] != (double)(threadIdx.x * 1000 + c + 2 * i)) bad++;
    }
    if (bad) atomicAdd(err, bad);
    double acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < 128; ++i) {
        double v[8];
        tm_ld16(tb + ((i * 6) % (ncol - 16)), v);
        tm_wait_ld();
        acc += v[0] + v[7];
    }
    long long t1 = clock64();
    for (int i = 0; i < 128; i += 4) {
        double v0[8], v1[8], v2[8], v3[8];
        tm_ld16(tb + (((i + 0) * 6) % (ncol - 16)), v0);
        tm_ld16(tb + (((i + 1) * 6) % (ncol - 16)), v1);
        tm_ld16(tb + (((i + 2) * 6) % (ncol - 16)), v2);
        tm_ld16(tb + (((i + 3) * 6) % (ncol - 16)), v3);
        tm_wait_ld();
        acc += v0[1] + v1[2] + v2[3] + v3[4];
    }
    long long t2 = clock64();
    out[threadIdx.x] = acc;
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base));
}
__global__ void k(double *out, long long *cyc, int *err)
{
    __shared__ uint32_t s_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = s_base;
    // this warp's lane window: lane field = 32*(warp%4) in bits [31:16]; columns split by warp/4
    const uint32_t nslots = blockDim.x / 128;              // warps sharing a lane quadrant
    const uint32_t ncol = 512 / (nslots ? nslots : 1);     // columns per thread
    const uint32_t tb = base + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * ncol;
    // write
    for (uint32_t c = 0; c < ncol; c += 2) tm_st(tb + c, (double)(threadIdx.x * 1000 + c));
    tm_wait_st();
    __syncthreads();
    // read back + verify
    int bad = 0;
    for (uint32_t c = 0; c < ncol; c += 2) {
        double v = tm_ld(tb + c);
        tm_wait_ld();
        if (v != (double)(threadIdx.x * 1000 + c)) bad++;
    }
    if (bad) atomicAdd(err, bad);
    // latency: dependent ld -> wait -> ld ...
    double acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < 256; ++i) {
        double v = tm_ld(tb + ((i * 2) & (ncol - 1)));
        tm_wait_ld();
        acc += v;
    }
    long long t1 = clock64();
    // throughput: 8 loads then one wait
    double a2 = 0;
    for (int i = 0; i < 256; i += 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = tm_ld(tb + (((i + u) * 2) & (ncol - 1)));
        tm_wait_ld();
#pragma unroll
        for (int u = 0; u < 8; ++u) a2 += v[u];
    }
    long long t2 = clock64();
    // store latency
    for (int i = 0; i < 256; ++i) {
        tm_st(tb + ((i * 2) & (ncol - 1)), acc + i);
        tm_wait_st();
    }
    long long t3 = clock64();
    out[threadIdx.x] = acc + a2;
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base));
}
int main()
{
    double *d; long long *c; int *e; cudaMalloc(&d, 1 << 16); cudaMalloc(&c, 64); cudaMalloc(&e, 4);
    for (int threads : {32, 128, 256, 512}) {
        cudaMemset(e, 0, 4);
        k<<<1, threads>>>(d, c, e);
        cudaError_t er = cudaDeviceSynchronize();
        long long h[3]; int bad;
        cudaMemcpy(h, c, 24, cudaMemcpyDeviceToHost); cudaMemcpy(&bad, e, 4, cudaMemcpyDeviceToHost);
        printf("threads %d: err=%s mismatches=%d  ld+wait %.1f clk  8xld+wait %.1f clk/ld  st+wait %.1f clk\n", threads,
               cudaGetErrorString(er), bad, h[0] / 256.0, h[1] / 256.0, h[2] / 256.0);
    }
    for (int threads : {32, 128, 256, 512}) {
        cudaMemset(e, 0, 4);
        k16<<<1, threads>>>(d, c, e);
        cudaError_t er = cudaDeviceSynchronize();
        long long h[3]; int bad;
        cudaMemcpy(h, c, 24, cudaMemcpyDeviceToHost); cudaMemcpy(&bad, e, 4, cudaMemcpyDeviceToHost);
        printf("x16 threads %d: err=%s mismatches=%d  ld16+wait %.1f clk  4xld16+wait %.1f clk/ld16\n", threads,
               cudaGetErrorString(er), bad, h[0] / 128.0, h[1] / 128.0);
    }
    // all SMs busy
    cudaMemset(e, 0, 4);
    k<<<148, 256>>>(d, c, e);
    cudaError_t er = cudaDeviceSynchronize();
    int bad; cudaMemcpy(&bad, e, 4, cudaMemcpyDeviceToHost);
    printf("148 CTAs x 256: err=%s mismatches=%d\n", cudaGetErrorString(er), bad);
    return 0;
}
