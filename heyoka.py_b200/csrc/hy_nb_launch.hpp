// hy_nb_launch.hpp - launch entry points of the register-resident N-body kernels.
// Every body count is instantiated in its own translation unit (hy_nb3.cu ... hy_nb6.cu:
// the fully unrolled order loops are long, nvcc compiles the units in parallel).
#pragma once
#include <cuda_runtime.h>

#include "hy_kernels.cuh"

namespace hy {

// fx: the build with the extended features (hy_kernels.cuh, template parameter FX)
template <typename R, int NB> cudaError_t launch_nbody_kernel(const KParams<R> &P, const hy_launch_info &li, cudaStream_t s, bool fx);
template <typename R, int NB> int regs_nbody_kernel();
// warpgroup-rotation variant (experimental, HY_CUDA_WGX=1): FP64, 6 bodies, order 20 only
cudaError_t launch_nbody_kernel_wgx(const KParams<double> &P, const hy_launch_info &li, cudaStream_t s);

// 6 bodies, FP64, unrolled to order NBR_LMAX = 22 (tol = 1e-18, the reference's benchmark configuration)
cudaError_t launch_nbody_kernel_p22(const KParams<double> &P, const hy_launch_info &li, cudaStream_t s, bool fx);
int regs_nbody_kernel_p22();
// register-resident CR3BP kernel (hy_cr3bp_reg.cuh, instantiated in hy_cr3bp.cu)
template <typename R> cudaError_t launch_cr3bp_kernel(const KParams<R> &P, const hy_launch_info &li, cudaStream_t s, bool fx);
template <typename R> int regs_cr3bp_kernel();
// FP64 build unrolled to order 22 (kernel variant CRB_VARIANT_P22)
cudaError_t launch_cr3bp_kernel_p22(const KParams<double> &P, const hy_launch_info &li, cudaStream_t s, bool fx);
int regs_cr3bp_kernel_p22();

// launch one instantiation of the persistent kernel
template <typename K, typename R>
inline cudaError_t launch_kernel_fn(K kern, const KParams<R> &P, const hy_launch_info &li, cudaStream_t s)
{
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)li.smem_bytes);
    if (e != cudaSuccess) return e;
    kern<<<li.ctas, li.threads, li.smem_bytes, s>>>(P);
    return cudaGetLastError();
}

#define HY_NB_INSTANTIATE(NB)                                                                            \
    template <typename R, int N> cudaError_t launch_nbody_kernel(const KParams<R> &P, const hy_launch_info &li, \
                                                                 cudaStream_t s, bool fx)               \
    {                                                                                                    \
        if (fx) return launch_kernel_fn(propagate_kernel<R, 16, true, N, false, NBR_PMAX, true>, P, li, s); \
        return launch_kernel_fn(propagate_kernel<R, 16, true, N, false, NBR_PMAX, false>, P, li, s);     \
    }                                                                                                    \
    template <typename R, int N> int regs_nbody_kernel()                                                 \
    {                                                                                                    \
        cudaFuncAttributes a{};                                                                          \
        if (cudaFuncGetAttributes(&a, propagate_kernel<R, 16, true, N, false, NBR_PMAX, false>) != cudaSuccess) return 0; \
        return a.numRegs;                                                                                \
    }                                                                                                    \
    template cudaError_t launch_nbody_kernel<double, NB>(const KParams<double> &, const hy_launch_info &, cudaStream_t, bool); \
    template cudaError_t launch_nbody_kernel<float, NB>(const KParams<float> &, const hy_launch_info &, cudaStream_t, bool);  \
    template int regs_nbody_kernel<double, NB>();                                                        \
    template int regs_nbody_kernel<float, NB>();

} // namespace hy
