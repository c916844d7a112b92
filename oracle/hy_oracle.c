/*
 * hy_oracle.c - CPU restatement of the reference's batch Taylor integrator,
 * evaluated over the opcode tape of include/hy_cuda.h.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product (heyoka.py_b200/) never does.
 *
 * Algorithm source.  The arithmetic of this path lives in the heyoka C++
 * library, pinned 7.11.0 by /root/reference/CMakeLists.txt:127 and absent
 * from /root/reference (not vendored; needs LLVM/Boost/TBB/fmt, none of which
 * are installed; no network) - so it is "unbuildable here" and this file
 * restates its published algorithm (papers cited in
 * /root/reference/README.md: arXiv:2105.00800, arXiv:2204.09948), anchored on
 * the reference's own call sites:
 *   step ............ /root/reference/heyoka/expose_batch_integrators.cpp:233-241
 *   propagate_* ..... /root/reference/heyoka/expose_batch_integrators.cpp:243-314
 *   time (hi, lo) ... /root/reference/heyoka/expose_batch_integrators.cpp:407-449
 *   ensemble ........ /root/reference/heyoka/_ensemble_impl.py:23-68 (OpenMP here)
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this oracle (and
 * its numpy twin oracle/np_oracle.py, which does not use the tape) against
 * the absolute values printed in the reference's notebooks.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/hy_cuda.h"

#define HY_MAX_ORDER 64

/* ---- double ---- */
#define REAL double
#define SUF _f64
#define R_FMA fma
#define R_SQRT sqrt
#define R_POW pow
#define R_EXP exp
#define R_LOG log
#define R_SIN sin
#define R_COS cos
#define R_ASIN asin
#define R_ACOS acos
#define R_ATAN atan
#define R_ERF erf
#define R_ABS fabs
#define R_COPYSIGN copysign
#include "hy_oracle_impl.h"
#undef REAL
#undef SUF
#undef R_FMA
#undef R_SQRT
#undef R_POW
#undef R_EXP
#undef R_LOG
#undef R_SIN
#undef R_COS
#undef R_ASIN
#undef R_ACOS
#undef R_ATAN
#undef R_ERF
#undef R_ABS
#undef R_COPYSIGN
#undef JET
#undef FN
#undef CAT
#undef CAT_

/* ---- float ---- */
#define REAL float
#define SUF _f32
#define R_FMA fmaf
#define R_SQRT sqrtf
#define R_POW powf
#define R_EXP expf
#define R_LOG logf
#define R_SIN sinf
#define R_COS cosf
#define R_ASIN asinf
#define R_ACOS acosf
#define R_ATAN atanf
#define R_ERF erff
#define R_ABS fabsf
#define R_COPYSIGN copysignf
#include "hy_oracle_impl.h"

int ora_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
