// Register-resident CR3BP kernel (see hy_cr3bp_reg.cuh, hy_nb_launch.hpp): two lanes per trajectory.
#include "hy_nb_launch.hpp"

namespace hy {

template <typename R> cudaError_t launch_cr3bp_kernel(const KParams<R> &P, const hy_launch_info &li, cudaStream_t s, bool fx)
{
    if (fx) return launch_kernel_fn(propagate_kernel<R, 2, true, -1, false, NBR_PMAX, true>, P, li, s);
    return launch_kernel_fn(propagate_kernel<R, 2, true, -1, false, NBR_PMAX, false>, P, li, s);
}
template <typename R> int regs_cr3bp_kernel()
{
    cudaFuncAttributes a{};
    if (cudaFuncGetAttributes(&a, propagate_kernel<R, 2, true, -1, false, NBR_PMAX, false>) != cudaSuccess) return 0;
    return a.numRegs;
}
cudaError_t launch_cr3bp_kernel_p22(const KParams<double> &P, const hy_launch_info &li, cudaStream_t s, bool fx)
{
    if (fx) return launch_kernel_fn(propagate_kernel<double, 2, true, -1, false, CRB_PMAX_HI, true>, P, li, s);
    return launch_kernel_fn(propagate_kernel<double, 2, true, -1, false, CRB_PMAX_HI, false>, P, li, s);
}
int regs_cr3bp_kernel_p22()
{
    cudaFuncAttributes a{};
    if (cudaFuncGetAttributes(&a, propagate_kernel<double, 2, true, -1, false, CRB_PMAX_HI, false>) != cudaSuccess) return 0;
    return a.numRegs;
}
template cudaError_t launch_cr3bp_kernel<double>(const KParams<double> &, const hy_launch_info &, cudaStream_t, bool);
template cudaError_t launch_cr3bp_kernel<float>(const KParams<float> &, const hy_launch_info &, cudaStream_t, bool);
template int regs_cr3bp_kernel<double>();
template int regs_cr3bp_kernel<float>();

} // namespace hy
