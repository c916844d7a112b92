/* hy_cuda_cout.h - continuous-output records rebuilt from host data.
 *
 * The reference's continuous_output_batch<T> is a value: it can be copied, deep-copied and pickled
 * (/root/reference/heyoka/taylor_expose_c_output.cpp:260-526, pickle support through
 * pickle_wrappers.hpp:35-73; exercised by /root/reference/heyoka/test.py:1414-1771 and by
 * process-based ensembles with c_output=True, _ensemble_impl.py:102-138).  A record of libhy_cuda lives in
 * device memory; hy_cout_get exports it (tcs [S, n, order+1, B], times (hi, lo) [S+1, B], NaN past a
 * lane's own step count) and this call is the way back: a new record, on `device`, that evaluates
 * exactly like the one the arrays came from.  Free it with hy_cout_free(rec, NULL). */
#ifndef HY_CUDA_COUT_H
#define HY_CUDA_COUT_H
#include "hy_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* n_steps [B]: recorded steps of every lane (<= S). */
int hy_cout_from_host(hy_cout **out, int device, int fp_bits, uint32_t n_state, uint32_t order, uint32_t batch,
                      const uint64_t *n_steps, const void *tcs, const void *times_hi, const void *times_lo, uint64_t S);

#ifdef __cplusplus
}
#endif
#endif /* HY_CUDA_COUT_H */
