// Register-resident N-body kernel for 6 bodies (see hy_nbody_reg.cuh, hy_nb_launch.hpp).
#include "hy_nb_launch.hpp"

namespace hy {
HY_NB_INSTANTIATE(6)

cudaError_t launch_nbody_kernel_p22(const KParams<double> &P, const hy_launch_info &li, cudaStream_t s, bool fx)
{
    if (fx) return launch_kernel_fn(propagate_kernel<double, 16, true, 6, false, NBR_LMAX, true>, P, li, s);
    return launch_kernel_fn(propagate_kernel<double, 16, true, 6, false, NBR_LMAX, false>, P, li, s);
}
int regs_nbody_kernel_p22()
{
    cudaFuncAttributes a{};
    if (cudaFuncGetAttributes(&a, propagate_kernel<double, 16, true, 6, false, NBR_LMAX, false>) != cudaSuccess) return 0;
    return a.numRegs;
}

cudaError_t launch_nbody_kernel_wgx(const KParams<double> &P, const hy_launch_info &li, cudaStream_t s)
{
    auto kern = propagate_kernel<double, 16, true, 6, true, NBR_PMAX, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)li.smem_bytes);
    if (e != cudaSuccess) return e;
    kern<<<li.ctas, li.threads, li.smem_bytes, s>>>(P);
    return cudaGetLastError();
}
}
