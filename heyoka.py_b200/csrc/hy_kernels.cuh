// hy_kernels.cuh - sm_100a kernels of the batch Taylor integrator.
//
// Mapping:
//  * a GROUP of G threads (G = 1..32, a power of two, always inside one warp)
//    cooperates on one trajectory; a CTA keeps T trajectories resident, one CTA
//    per SM, persistent: finished trajectories are retired and the group refills
//    itself from a global work counter (warp-level vote + one atomic per warp).
//  * the jets of a trajectory live in a per-CTA workspace ws[row][T] (trajectory
//    index fastest, row stride TS odd => conflict-free 64-bit accesses) held in
//    SHARED MEMORY: nothing but the initial/final state crosses HBM.
//  * the tape (ops, terms, level boundaries) is staged once per CTA in shared
//    memory; the ops of one dependency level are spread over the G lanes of the
//    group and levels are separated by __syncwarp(group mask).  The groups of a
//    warp run the same program in lockstep (re-aligned at every iteration).
//  * matched tapes skip the interpreter: register-resident jets for N-body systems
//    (hy_nbody_reg.cuh, NB > 0) and for the CR3BP (hy_cr3bp_reg.cuh, NB < 0); the
//    persistent loop, the tail of the step and every API feature are shared code.
//
// This replaces the reference's JIT-compiled taylor_step + C++ propagate loop
// (/root/reference/heyoka/expose_batch_integrators.cpp:233-314 -> [UPSTREAM]).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hy_cuda.h"
#ifdef __CUDACC_RTC__
#include "hy_devprog.h" // run-time compiled kernels (hy_jit.hpp): no host code in the translation unit
#else
#include "hy_schedule.hpp"
#endif
#include "hy_events.cuh"
#include "hy_evtape.cuh"
#include "hy_nbody_reg.cuh"
#include "hy_cr3bp_reg.cuh"

// Run-time compiled kernels (hy_jit.hpp) define HY_JIT and HY_WS = 32: one thread per trajectory,
// workspace interleaved over the lanes of a warp (row r of lane l at [r * 32 + l]), the order
// sweep generated from the tape (hy_gen_jets) instead of interpreted.
#ifndef HY_WS
#define HY_WS 1
#endif

namespace hy {

#ifdef HY_JIT_EVT
// Event functions of a register-resident kernel as generated code (hy_jit.hpp): all events on the
// first lane of the trajectory's group.  norms: the ops needed at every order over all orders, the
// rest at orders 0, p-1, p; order: the rest at one more order (rare path: an event may happen);
// interval: enclosures of the event functions over the step, true if one of them contains 0.
template <typename R, int XS>
__device__ __forceinline__ void hy_gen_evt_norms(const EvtCtx<R, XS> &C, const ETerm *terms, R *scr, uint32_t sub);
template <typename R, int XS>
__device__ __forceinline__ void hy_gen_evt_order(const EvtCtx<R, XS> &C, const ETerm *terms, uint32_t k);
template <typename R, int XS>
__device__ __forceinline__ bool hy_gen_evt_interval(const R *w, R *iv, const ETerm *terms, const double *imm, R t0, R h,
                                                    uint32_t sub);
#endif
#ifdef HY_JIT
// generated per tape: orders 0 .. p-1 of every op / the event-function ops at order p
template <typename R> __device__ __forceinline__ void hy_gen_jets(R *__restrict__ w, const R *__restrict__ rk, const R tm);
template <typename R> __device__ __forceinline__ void hy_gen_ev_sweep(R *__restrict__ w, const R *__restrict__ rk, const R tm);
#endif

// ---- precision-generic math wrappers ----
__device__ __forceinline__ double r_fma(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float r_fma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double r_sqrt(double a) { return sqrt(a); }
__device__ __forceinline__ float r_sqrt(float a) { return sqrtf(a); }
// libm calls are kept out of line: one copy each, so the hot loop stays small.
static __device__ __noinline__ double r_pow(double a, double b) { return pow(a, b); }
static __device__ __noinline__ float r_pow(float a, float b) { return powf(a, b); }
static __device__ __noinline__ double r_exp(double a) { return exp(a); }
static __device__ __noinline__ float r_exp(float a) { return expf(a); }
static __device__ __noinline__ double r_log(double a) { return log(a); }
static __device__ __noinline__ float r_log(float a) { return logf(a); }
static __device__ __noinline__ void r_sincos(double a, double *s, double *c) { sincos(a, s, c); }
static __device__ __noinline__ void r_sincos(float a, float *s, float *c) { sincosf(a, s, c); }
__device__ __forceinline__ double r_abs(double a) { return fabs(a); }
__device__ __forceinline__ float r_abs(float a) { return fabsf(a); }
__device__ __forceinline__ double r_copysign(double a, double b) { return copysign(a, b); }
__device__ __forceinline__ float r_copysign(float a, float b) { return copysignf(a, b); }
// x^e for the step-size radii (x > 0, e = 1/p): exp(log(x) * e).  The error of log() is
// scaled by e <= 1/2, so the result is good to ~1 ulp - at a third of the cost of pow().
static __device__ __noinline__ double r_root(double x, double e) { return exp(log(x) * e); }
static __device__ __noinline__ float r_root(float x, float e) { return expf(logf(x) * e); }
// order 0 of the INTG functions (HY_OP_INTG: imm = 0 asin, 1 acos, 2 atan, 3 erf)
static __device__ __noinline__ double r_intg0(double a, int code)
{
    return code == 0 ? asin(a) : (code == 1 ? acos(a) : (code == 2 ? atan(a) : erf(a)));
}
static __device__ __noinline__ float r_intg0(float a, int code)
{
    return code == 0 ? asinf(a) : (code == 1 ? acosf(a) : (code == 2 ? atanf(a) : erff(a)));
}
template <typename R> __device__ __forceinline__ R r_inf();
template <> __device__ __forceinline__ double r_inf<double>() { return __longlong_as_double(0x7ff0000000000000LL); }
template <> __device__ __forceinline__ float r_inf<float>() { return __int_as_float(0x7f800000); }

template <typename R> __device__ __forceinline__ R shfl_xor(unsigned mask, R v, int lanemask);
template <> __device__ __forceinline__ double shfl_xor<double>(unsigned mask, double v, int lm)
{
    return __shfl_xor_sync(mask, v, lm);
}
template <> __device__ __forceinline__ float shfl_xor<float>(unsigned mask, float v, int lm)
{
    return __shfl_xor_sync(mask, v, lm);
}

// NaN-propagating max (once NaN, stays NaN).
template <typename R> __device__ __forceinline__ R nan_max(R m, R a) { return (a > m || a != a) ? a : m; }

enum { MODE_UNTIL = 0, MODE_FOR = 1, MODE_STEP = 2, MODE_GRID = 3 };

struct ProgDims {
    uint32_t n_slots, n_tslots, n_imm, n_phases;
    uint32_t ws_len, par_off, one_off, n_spill; // device workspace layout (hy_schedule.hpp)
    uint32_t evt_bytes;                          // event tape staged in shared memory (hy_evtape.cuh), or 0
};

// Internal outcome: the lane stopped inside a launch and can be resumed by the next one (per-launch
// step budget, a non-terminal event whose callback must run on the host, recorder pool exhausted).
// Never shown to callers of the Python API.
// HY_OUTCOME_PAUSED is declared in include/hy_cuda.h (the host loop of the front end sees it);
// HY_OUTCOME_PAUSED_POOL never leaves libhy_cuda: the lane ran out of recorder pool and is
// resumed by hy_propagate_ex itself after it has grown the pool.
#define HY_OUTCOME_PAUSED_POOL (-4294967401LL)

// Continuous-output recorder (single pass): every lane appends its steps to a linked list of
// fixed-size CHUNKS taken from a pool by bump allocation.  A chunk holds HY_REC_CH step records
// of rec_len elements (n*(p+1) Taylor coefficients, then the end time (hi, lo) of the step),
// preceded by a 2-element header (element 0: id of the next chunk, as an integer).
constexpr uint32_t HY_REC_CH = 8;
constexpr uint32_t HY_REC_NONE = 0xffffffffu;
template <typename R> struct RecDev {
    R *const *seg;          // [n_seg] segment base pointers (device array)
    uint32_t seg_chunks;    // chunks per segment
    uint32_t cap_chunks;    // chunks in all segments
    unsigned int *next;     // bump counter
    uint32_t *head, *tail;  // [B] first / last chunk of the lane
    uint32_t *count;        // [B] recorded steps
    R *t0_hi, *t0_lo;       // [B] time at which the lane's record starts
    uint32_t rec_len, chunk_len;
    // element (variable i, order k) of a record is at i * si + k * sk: (p + 1, 1) - the reference's tc
    // layout - or (1, n), order-major: the layout of the CR3BP kernel's column in shared memory, whose
    // FP64 records leave the SM as ONE bulk copy (TMA) per step instead of 126 element copies
    uint32_t si, sk;
    int on, append;
    __device__ __forceinline__ R *chunk(uint32_t cid) const
    {
        return seg[cid / seg_chunks] + (size_t)(cid % seg_chunks) * chunk_len;
    }
};

template <typename R> struct KParams {
    hy_dims d;
    const void *prog;            // [ops | terms | imm] in the smem layout
    const uint32_t *phase_slot;  // [n_phases + 1]
    const uint32_t *ev_ref;      // [n_events] device row of each event jet
    const uint32_t *state_row;   // [n_state] device row of each state variable
    const int32_t *state_spill;  // [n_state] spill slot or -1
    R *gjet;                     // spilled state jets: [ctas][T][n_spill][p+1]
    ProgDims pd;
    R *state;      // [n][B]
    const R *pars; // [m][B]
    R *t_hi, *t_lo, *last_h;
    const R *tf_hi, *tf_lo;      // [B] absolute final time, double-length (MODE_UNTIL / MODE_FOR)
    const R *mdt; // [B] or nullptr
    const unsigned char *active; // [B] lanes taking part in this launch (nullptr: all)
    uint32_t *gidx;              // [B] next grid point of the lane (MODE_GRID; kept for resumed launches)
    int resume;                  // lanes continue from the counters / grid index of the previous launch
    int pause_on_nt;             // stop a lane (PAUSED) after a step that logged a non-terminal event
    unsigned long long launch_steps; // per-launch step budget (0: none): the lane stops with PAUSED
    const uint32_t *rec_off;     // [n (p+1)] register-resident kernels: column offset of record element (i, k)
    const uint32_t *red_idx;     // [n_red] state variables reduced to [0, 2 pi) after every step
    uint32_t n_red;
    RecDev<R> rec;               // continuous-output recorder (rec.on)
    EvtDev evt;                  // event tape of the register-resident kernels (evt.n_events)
    long long *outcome;
    R *min_h, *max_h;
    unsigned long long *n_steps;
    R *tc; // [n][p+1][B] or nullptr
    // propagate_grid: grid [K][B] in, gout [K][n][B] out
    const R *grid;
    R *gout;
    uint32_t grid_k;
    EvParams<R> ev; // event detection state (n_events > 0)
    unsigned int *counter;
    R *gws; // global workspace fallback (ws_in_smem == 0)
    uint32_t B, T, TS;
    uint32_t nb_tb_off; // register-resident N-body path: offset of the exchange buffer in the column
    uint32_t wgx_wgs;   // WGX: warpgroups that take work (3; fewer for experiments)
    unsigned long long max_steps;
    int mode, backward, write_tc, high_accuracy, ws_in_smem;
    R rhofac, inv_p, inv_pm1;
};

// ---------------------------------------------------------------------------
// Device-side program (built by hy_schedule.hpp): per-lane op streams in a
// lane-interleaved layout, 16-byte ops, 16-byte terms.
//
// Workspace of ONE trajectory (contiguous, RS elements, RS odd):
//   [0, n_rows)                 jets / cur rows of the u-variables
//   [n_rows, n_rows+n_par)      runtime parameters
//   [n_rows+n_par, +P1)         unit jet [1, 0, ..., 0]   (also "par = 1.0")
// Orders of a jet are contiguous, so a convolution walks a[+j], b[-j] with
// immediate offsets.  Trajectory t of the CTA lives at ws + t*RS.
// ---------------------------------------------------------------------------
// device reference -> element offset at order k: jets advance by k, spilled
// (ping-pong) state variables by k & 1, single rows not at all.
__device__ __forceinline__ uint32_t roff(uint32_t ref, uint32_t k)
{
    return (ref & 0x3fffffffu) + ((ref & HY_DREF_JET) ? k : ((ref & HY_DREF_PP) ? (k & 1u) : 0u));
}
__device__ __forceinline__ uint32_t rbase(uint32_t ref) { return ref & 0x3fffffffu; }
// operand offset of a DOp field
__device__ __forceinline__ uint32_t ooff(uint32_t off, bool jet, bool pp, uint32_t k)
{
    return off + (jet ? k : (pp ? (k & 1u) : 0u));
}

template <typename R> __device__ __noinline__ R pow0(R x, double alpha)
{
    if (alpha == -1.5) return (R)1 / (x * r_sqrt(x));
    if (alpha == -0.5) return (R)1 / r_sqrt(x);
    if (alpha == 1.5) return x * r_sqrt(x);
    if (alpha == -1.0) return (R)1 / x;
    if (alpha == -2.0) return (R)1 / (x * x);
    return r_pow(x, (R)alpha);
}

// ---------------------------------------------------------------------------
// Convolution bodies.  A single warp has to hide its own LDS latency (29 clk)
// and DFMA latency (8 clk): every body is a loop over BLOCKS of terms; inside a
// block all loads are predicated (no control flow), issued back-to-back, and
// feed independent accumulator chains.  No per-term branches, no per-term
// address arithmetic (immediate offsets from two base registers).
// ---------------------------------------------------------------------------
// sum_{j=0}^{n-1} pa[j] * pb[-j]
// (S: element stride of the workspace - 1 for a trajectory column, HY_WS for the warp-interleaved
//  workspace of the run-time compiled kernels)
template <typename R, int S = 1>
__device__ __forceinline__ R conv(const R *__restrict__ pa, const R *__restrict__ pb, int n)
{
    R s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll 1
    for (int j0 = 0; j0 < n; j0 += 8) {
        R a[8], b[8];
        const int m = n - j0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const bool v = u < m;
            a[u] = v ? pa[(j0 + u) * S] : (R)0;
            b[u] = v ? pb[-(j0 + u) * S] : (R)0;
        }
        s0 = r_fma(a[0], b[0], s0);
        s1 = r_fma(a[1], b[1], s1);
        s2 = r_fma(a[2], b[2], s2);
        s3 = r_fma(a[3], b[3], s3);
        s0 = r_fma(a[4], b[4], s0);
        s1 = r_fma(a[5], b[5], s1);
        s2 = r_fma(a[6], b[6], s2);
        s3 = r_fma(a[7], b[7], s3);
    }
    return (s0 + s1) + (s2 + s3);
}

#ifdef HY_JIT
// Run-time compiled kernels: the operands come from global memory (L1/L2/HBM), so a convolution
// must have ALL its loads in flight before the first FMA waits on one - every block of the loop
// above costs one memory latency.  Same terms, same four chains in the same order as conv():
// identical rounding.  Out of line: the generated sweep calls it once per product (small code,
// the instruction cache holds the whole sweep).
template <typename R, int S, int NT> __device__ __forceinline__ R conv_wide_n(const R *__restrict__ pa, const R *__restrict__ pb, int n)
{
    R a[NT], b[NT];
#pragma unroll
    for (int u = 0; u < NT; ++u) {
        const bool v = u < n;
        a[u] = v ? pa[u * S] : (R)0;
        b[u] = v ? pb[-u * S] : (R)0;
    }
    R s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
    for (int u = 0; u < NT; u += 4) {
        s0 = r_fma(a[u], b[u], s0);
        s1 = r_fma(a[u + 1], b[u + 1], s1);
        s2 = r_fma(a[u + 2], b[u + 2], s2);
        s3 = r_fma(a[u + 3], b[u + 3], s3);
    }
    return (s0 + s1) + (s2 + s3);
}
template <typename R, int S> __device__ __noinline__ R conv_wide(const R *__restrict__ pa, const R *__restrict__ pb, int n)
{
    if (n <= 8) return conv_wide_n<R, S, 8>(pa, pb, n);
    if (n <= 16) return conv_wide_n<R, S, 16>(pa, pb, n);
    if (n <= 24) return conv_wide_n<R, S, 24>(pa, pb, n);
    return conv<R, S>(pa, pb, n);
}
#endif

// Three products sharing the operand b:  o_i = sum_{j<n} a_i[j] * pb[-j]
template <typename R, int S = 1>
__device__ __forceinline__ void conv3(const R *__restrict__ a0, const R *__restrict__ a1, const R *__restrict__ a2,
                                      const R *__restrict__ pb, int n, R &o0, R &o1, R &o2)
{
    R s0a = 0, s1a = 0, s2a = 0, s0b = 0, s1b = 0, s2b = 0;
#pragma unroll 1
    for (int j0 = 0; j0 < n; j0 += 4) {
        R x[4], y[4], z[4], b[4];
        const int m = n - j0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const bool v = u < m;
            b[u] = v ? pb[-(j0 + u) * S] : (R)0;
            x[u] = v ? a0[(j0 + u) * S] : (R)0;
            y[u] = v ? a1[(j0 + u) * S] : (R)0;
            z[u] = v ? a2[(j0 + u) * S] : (R)0;
        }
        s0a = r_fma(x[0], b[0], s0a);
        s1a = r_fma(y[0], b[0], s1a);
        s2a = r_fma(z[0], b[0], s2a);
        s0b = r_fma(x[1], b[1], s0b);
        s1b = r_fma(y[1], b[1], s1b);
        s2b = r_fma(z[1], b[1], s2b);
        s0a = r_fma(x[2], b[2], s0a);
        s1a = r_fma(y[2], b[2], s1a);
        s2a = r_fma(z[2], b[2], s2a);
        s0b = r_fma(x[3], b[3], s0b);
        s1b = r_fma(y[3], b[3], s1b);
        s2b = r_fma(z[3], b[3], s2b);
    }
    o0 = s0a + s0b;
    o1 = s1a + s1b;
    o2 = s2a + s2b;
}

#ifdef HY_JIT
// conv3 with 12-term blocks (48 loads in flight); chains as in conv3: even / odd terms.
template <typename R, int S>
__device__ __noinline__ void conv3_wide(const R *__restrict__ a0, const R *__restrict__ a1, const R *__restrict__ a2,
                                        const R *__restrict__ pb, int n, R &o0, R &o1, R &o2)
{
    R s0a = 0, s1a = 0, s2a = 0, s0b = 0, s1b = 0, s2b = 0;
#pragma unroll 1
    for (int j0 = 0; j0 < n; j0 += 12) {
        R x[12], y[12], z[12], b[12];
        const int m = n - j0;
#pragma unroll
        for (int u = 0; u < 12; ++u) {
            const bool v = u < m;
            b[u] = v ? pb[-(j0 + u) * S] : (R)0;
            x[u] = v ? a0[(j0 + u) * S] : (R)0;
            y[u] = v ? a1[(j0 + u) * S] : (R)0;
            z[u] = v ? a2[(j0 + u) * S] : (R)0;
        }
#pragma unroll
        for (int u = 0; u < 12; u += 2) {
            s0a = r_fma(x[u], b[u], s0a);
            s1a = r_fma(y[u], b[u], s1a);
            s2a = r_fma(z[u], b[u], s2a);
            s0b = r_fma(x[u + 1], b[u + 1], s0b);
            s1b = r_fma(y[u + 1], b[u + 1], s1b);
            s2b = r_fma(z[u + 1], b[u + 1], s2b);
        }
    }
    o0 = s0a + s0b;
    o1 = s1a + s1b;
    o2 = s2a + s2b;
}
#endif

#ifdef HY_JIT
// NT convolutions that share one operand, every element of it loaded ONCE: o_t = sum_{j<n} ps[+-j] * pt[t][-+j]
// (FWD: the shared operand is the one read forwards).  Each sum is chained exactly like conv_wide / conv():
// term j on chain j mod 4, folded (s0 + s1) + (s2 + s3) - the results are bit-identical to NT separate calls.
// Blocks of 24 (NT <= 2) or 12 terms: all loads of a block are in flight before its first FMA.
// (CTAs of more than 256 threads have 128 registers per thread: at most two products per call, 12-term blocks)
#if defined(HY_JIT_THREADS) && HY_JIT_THREADS > 256
#define HY_JIT_SMALLREG 1
#else
#define HY_JIT_SMALLREG 0
#endif
template <typename R, int S, int NT, bool FWD>
__device__ __forceinline__ void convn_wide(const R *__restrict__ ps, const R *const (&pt)[NT], int n, R (&o)[NT])
{
    constexpr int BL = (NT <= 2 && !HY_JIT_SMALLREG) ? 24 : 12;
    R s[NT][4];
#pragma unroll
    for (int t = 0; t < NT; ++t) s[t][0] = s[t][1] = s[t][2] = s[t][3] = 0;
#pragma unroll 1
    for (int j0 = 0; j0 < n; j0 += BL) {
        R sv[BL], tv[NT][BL];
        const int m = n - j0;
#pragma unroll
        for (int u = 0; u < BL; ++u) {
            const bool v = u < m;
            sv[u] = v ? ps[(FWD ? (j0 + u) : -(j0 + u)) * S] : (R)0;
#pragma unroll
            for (int t = 0; t < NT; ++t) tv[t][u] = v ? pt[t][(FWD ? -(j0 + u) : (j0 + u)) * S] : (R)0;
        }
#pragma unroll
        for (int u = 0; u < BL; ++u)
#pragma unroll
            for (int t = 0; t < NT; ++t) s[t][u & 3] = r_fma(tv[t][u], sv[u], s[t][u & 3]);
    }
#pragma unroll
    for (int t = 0; t < NT; ++t) o[t] = (s[t][0] + s[t][1]) + (s[t][2] + s[t][3]);
}

// Software prefetch of the rows an upcoming op will read (rows [0, n) of a jet): needs no
// registers, so the memory round trip of op i + D overlaps the arithmetic of ops i .. i + D - 1.
template <typename R, int S, int LVL> __device__ __forceinline__ void pf_rows(const R *p, int n)
{
#pragma unroll 4
    for (int u = 0; u < n; ++u) {
        if (LVL == 1)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(p + u * S));
        else
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p + u * S));
    }
}
// ---- order-blocked Cauchy products (run-time compiled kernels) ----
// The jets of a large system stream from L2 / HBM: a plain sweep reads the whole history of both
// operands of every product at every order.  Blocked in groups of M orders, the history is read once
// per block: at the first order k0 of a block (k0 a multiple of M, k0 >= M) ONE pass over
// s[jlo .. jhi] and the window operand computes, for i = 0 .. M-1, the part of the order-(k0+i) sum
// whose two factors are both known already,
//     O_i = sum_{j = jlo+i}^{jhi} s[j] * w[K + i - j],
// O_0 is the complete sum of order k0; O_1 .. O_{M-1} are parked in scratch rows and completed at
// their own order with the 2 i products that involve the new coefficients (shortsum below).
//   product  c = a b :  s = a, w = b, jlo = 0, jhi = k0, K = k0
//   quotient c = a / b (sum_{j>=1} b[j] c[k-j]) :  s = b, w = c, jlo = 1, jhi = k0, K = k0
// NT streams share the window operand (MULSH).  Memory traffic of the products drops ~2.5x at
// M = 4, the number of dependent memory round trips per step ~4x.  The summation order differs from
// conv(): results agree with the interpreter to rounding, not bit for bit.
template <typename R, int S, int M, int NT, int CH>
__device__ __forceinline__ void blockconv_core(const R *const *ps, const R *__restrict__ pw, int n, R (&O)[NT][M])
{
    R wv[M];
#pragma unroll
    for (int i = 0; i < M; ++i) wv[i] = 0;
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int i = 0; i < M; ++i) O[t][i] = 0;
#pragma unroll 1
    for (int j0 = 0; j0 < n; j0 += CH) {
        R sv[NT][CH], wn[CH];
        const int m = n - j0;
#pragma unroll
        for (int u = 0; u < CH; ++u) {
            const bool v = u < m;
            wn[u] = v ? pw[-(j0 + u) * S] : (R)0;
#pragma unroll
            for (int t = 0; t < NT; ++t) sv[t][u] = v ? ps[t][(j0 + u) * S] : (R)0;
        }
#pragma unroll
        for (int u = 0; u < CH; ++u) {
#pragma unroll
            for (int i = M - 1; i > 0; --i) wv[i] = wv[i - 1];
            wv[0] = wn[u];
#pragma unroll
            for (int t = 0; t < NT; ++t)
#pragma unroll
                for (int i = 0; i < M; ++i) O[t][i] = r_fma(sv[t][u], wv[i], O[t][i]);
        }
    }
}
// sum_{j<i} x[j] y[-j], i < M (all loads first, nothing branches)
template <typename R, int S, int M> __device__ __forceinline__ R shortsum(const R *__restrict__ x, const R *__restrict__ y, uint32_t i)
{
    R xs[M > 1 ? M - 1 : 1], ys[M > 1 ? M - 1 : 1]; // (M = 1: never called, must still compile)
#pragma unroll
    for (int u = 0; u < M - 1; ++u) {
        const bool v = (uint32_t)u < i;
        xs[u] = v ? x[u * S] : (R)0;
        ys[u] = v ? y[-u * S] : (R)0;
    }
    R acc = 0;
#pragma unroll
    for (int u = 0; u < M - 1; ++u) acc = r_fma(xs[u], ys[u], acc);
    return acc;
}
#endif

// pow recurrence sum:  sum_{j<n} (kal - j*al1) * ak[-j] * c[j]
template <typename R, int S = 1>
__device__ __forceinline__ R conv_pow(const R *__restrict__ ak, const R *__restrict__ c, int n, const R al1, const R kal)
{
    R s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    R jr = 0;
#pragma unroll 1
    for (int j0 = 0; j0 < n; j0 += 8) {
        R a[8], b[8];
        const int m = n - j0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const bool v = u < m;
            a[u] = v ? ak[-(j0 + u) * S] : (R)0;
            b[u] = v ? c[(j0 + u) * S] : (R)0;
        }
#pragma unroll
        for (int u = 0; u < 8; u += 4) {
            s0 = r_fma(r_fma(-(jr + (R)u), al1, kal) * a[u], b[u], s0);
            s1 = r_fma(r_fma(-(jr + (R)(u + 1)), al1, kal) * a[u + 1], b[u + 1], s1);
            s2 = r_fma(r_fma(-(jr + (R)(u + 2)), al1, kal) * a[u + 2], b[u + 2], s2);
            s3 = r_fma(r_fma(-(jr + (R)(u + 3)), al1, kal) * a[u + 3], b[u + 3], s3);
        }
        jr += (R)8;
    }
    return (s0 + s1) + (s2 + s3);
}

// Same, terms j = 1 .. n-1 only (the caller supplies the j = 0 term from a register).
template <typename R>
__device__ __forceinline__ R conv_pow_from1(const R *__restrict__ ak, const R *__restrict__ c, int n, const R al1,
                                            const R kal)
{
    R s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    R jr = 1;
#pragma unroll 1
    for (int j0 = 1; j0 < n; j0 += 8) {
        R a[8], b[8];
        const int m = n - j0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const bool v = u < m;
            a[u] = v ? ak[-(j0 + u)] : (R)0;
            b[u] = v ? c[j0 + u] : (R)0;
        }
#pragma unroll
        for (int u = 0; u < 8; u += 4) {
            s0 = r_fma(r_fma(-(jr + (R)u), al1, kal) * a[u], b[u], s0);
            s1 = r_fma(r_fma(-(jr + (R)(u + 1)), al1, kal) * a[u + 1], b[u + 1], s1);
            s2 = r_fma(r_fma(-(jr + (R)(u + 2)), al1, kal) * a[u + 2], b[u + 2], s2);
            s3 = r_fma(r_fma(-(jr + (R)(u + 3)), al1, kal) * a[u + 3], b[u + 3], s3);
        }
        jr += (R)8;
    }
    return (s0 + s1) + (s2 + s3);
}

#ifdef HY_JIT
// ---------------------------------------------------------------------------
// Out-of-line op bodies of the run-time compiled kernels.  The generated sweep is a list of calls
// with literal row numbers ("threaded code"): a few instructions per op, so the whole sweep stays
// in the instruction cache however many ops the system has; the bodies below are shared by all
// ops and all warps.  Rows are element rows of the interleaved workspace (element r at w[r * S]).
// bi: position of the order inside a block of M (0xffffffff: this order is not blocked).
// The unblocked paths repeat exec_op's arithmetic statement by statement (identical rounding).
// ---------------------------------------------------------------------------
#define HY_NOBLK 0xffffffffu
template <typename R, int S, int M>
__device__ __noinline__ void jop_mul(R *__restrict__ w, uint32_t k, uint32_t bi, uint32_t a, uint32_t b, uint32_t dst, uint32_t q)
{
    const R *pa = w + a * S, *pb = w + b * S;
    if (M < 2 || bi == HY_NOBLK) {
        w[dst * S] = conv_wide<R, S>(pa, pb + k * S, (int)k + 1);
    } else if (bi == 0) {
        const R *const ps[1] = {pa};
        R O[1][M];
        blockconv_core<R, S, M, 1, 16>(ps, pb + k * S, (int)k + 1, O);
        w[dst * S] = O[0][0];
#pragma unroll
        for (int i = 1; i < M; ++i) w[(q + i - 1) * S] = O[0][i];
    } else {
        w[dst * S] = (w[(q + bi - 1) * S] + shortsum<R, S, M>(pa, pb + k * S, bi)) + shortsum<R, S, M>(pb, pa + k * S, bi);
    }
}
// NT products sharing the operand b (MULSH): rows a[t] -> dst[t]; scratch rows q + t (M - 1)
template <typename R, int S, int M, int NT> struct JRows {
    uint32_t a[NT], dst[NT];
};
template <typename R, int S, int M, int NT>
__device__ __noinline__ void jop_mulsh(R *__restrict__ w, uint32_t k, uint32_t bi, uint32_t b, const JRows<R, S, M, NT> r, uint32_t q)
{
    const R *pb = w + b * S;
    if (M < 2 || bi == HY_NOBLK) {
        if (NT == 3) {
            R s0, s1, s2;
            conv3_wide<R, S>(w + r.a[0] * S, w + r.a[1] * S, w + r.a[2 % NT] * S, pb + k * S, (int)k + 1, s0, s1, s2);
            w[r.dst[0] * S] = s0;
            w[r.dst[1] * S] = s1;
            w[r.dst[2 % NT] * S] = s2;
#ifndef HY_JIT_NOSHARE
        } else if (NT == 2 || NT == 4) {
            // (the shared operand's elements are loaded once for the NT products)
            const R *pt[NT];
#pragma unroll
            for (int t = 0; t < NT; ++t) pt[t] = w + r.a[t] * S;
            R o[NT];
            if (NT == 4 && HY_JIT_SMALLREG) {
                // (two pairs: the register budget of a 512-thread CTA)
                const R *p0[2] = {pt[0], pt[1]}, *p1[2] = {pt[2 % NT], pt[3 % NT]};
                R o0[2], o1[2];
                convn_wide<R, S, 2, false>(pb + k * S, p0, (int)k + 1, o0);
                convn_wide<R, S, 2, false>(pb + k * S, p1, (int)k + 1, o1);
                o[0] = o0[0];
                o[1] = o0[1];
                o[2 % NT] = o1[0];
                o[3 % NT] = o1[1];
            } else {
                convn_wide<R, S, NT, false>(pb + k * S, pt, (int)k + 1, o);
            }
#pragma unroll
            for (int t = 0; t < NT; ++t) w[r.dst[t] * S] = o[t];
#endif
        } else {
#pragma unroll
            for (int t = 0; t < NT; ++t) w[r.dst[t] * S] = conv_wide<R, S>(w + r.a[t] * S, pb + k * S, (int)k + 1);
        }
    } else if (bi == 0) {
        const R *ps[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) ps[t] = w + r.a[t] * S;
        R O[NT][M];
        blockconv_core<R, S, M, NT, (NT == 1 ? 16 : (NT == 2 ? 12 : 8))>(ps, pb + k * S, (int)k + 1, O);
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            w[r.dst[t] * S] = O[t][0];
#pragma unroll
            for (int i = 1; i < M; ++i) w[(q + t * (M - 1) + i - 1) * S] = O[t][i];
        }
    } else {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const R *pa = w + r.a[t] * S;
            w[r.dst[t] * S] = (w[(q + t * (M - 1) + bi - 1) * S] + shortsum<R, S, M>(pa, pb + k * S, bi)) +
                              shortsum<R, S, M>(pb, pa + k * S, bi);
        }
    }
}
// c = a / b: num = row of a at this order, b / c = jet bases, inv = row holding 1 / b[0]
template <typename R, int S, int M>
__device__ __noinline__ void jop_div(R *__restrict__ w, uint32_t k, uint32_t bi, uint32_t num, uint32_t b, uint32_t c,
                                     uint32_t inv, uint32_t q)
{
    const R *pb = w + b * S;
    R *pc = w + c * S;
    if (k == 0) w[inv * S] = (R)1 / pb[0];
    if (M < 2 || bi == HY_NOBLK) {
        R acc = w[num * S];
        if (k > 0) acc -= conv_wide<R, S>(pb + S, pc + (k - 1) * S, (int)k);
        pc[k * S] = acc * w[inv * S];
    } else if (bi == 0) {
        const R *const ps[1] = {pb + S};
        R O[1][M];
        blockconv_core<R, S, M, 1, 16>(ps, pc + (k - 1) * S, (int)k, O);
#pragma unroll
        for (int i = 1; i < M; ++i) w[(q + i - 1) * S] = O[0][i];
        const R acc = w[num * S] - O[0][0];
        pc[k * S] = acc * w[inv * S];
    } else {
        const R s = (w[(q + bi - 1) * S] + shortsum<R, S, M>(pb + S, pc + (k - 1) * S, bi)) + shortsum<R, S, M>(pc, pb + k * S, bi);
        pc[k * S] = (w[num * S] - s) * w[inv * S];
    }
}
// NT quotients by the same denominator b (grouped by the generator: DIV ops of one dependency level): b's
// coefficients are loaded once.  Per quotient the arithmetic is jop_div's, statement by statement.
template <int NT> struct JDivRows {
    uint32_t num[NT], c[NT], inv[NT];
};
template <typename R, int S, int NT>
__device__ __noinline__ void jop_divsh(R *__restrict__ w, uint32_t k, uint32_t b, const JDivRows<NT> r)
{
    const R *pb = w + b * S;
    if (k == 0) {
        const R iv = (R)1 / pb[0];
#pragma unroll
        for (int t = 0; t < NT; ++t) w[r.inv[t] * S] = iv;
    }
    R cv[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) cv[t] = 0;
    if (k > 0) {
        const R *pt[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) pt[t] = w + (r.c[t] + k - 1) * S;
        convn_wide<R, S, NT, true>(pb + S, pt, (int)k, cv);
    }
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        R acc = w[r.num[t] * S];
        if (k > 0) acc -= cv[t];
        w[(r.c[t] + k) * S] = acc * w[r.inv[t] * S];
    }
}
template <typename R, int S> __device__ __noinline__ void jop_square(R *__restrict__ w, uint32_t k, uint32_t a, uint32_t dst)
{
    const R *pa = w + a * S;
    R acc = conv_wide<R, S>(pa, pa + k * S, (int)((k + 1) >> 1));
    acc = acc + acc;
    if ((k & 1u) == 0) {
        const R m = pa[(k >> 1) * S];
        acc = r_fma(m, m, acc);
    }
    w[dst * S] = acc;
}
// one term of a sum of squares: acc += sum_{j<half} a[j] a[k-j]; acc2 = fma(a[k/2], a[k/2], acc2) for even k
template <typename R, int S> __device__ __noinline__ void jop_sumsq_term(const R *__restrict__ w, uint32_t k, uint32_t a, R &acc, R &acc2)
{
    const R *pa = w + a * S;
    acc += conv_wide<R, S>(pa, pa + k * S, (int)((k + 1) >> 1));
    if ((k & 1u) == 0) {
        const R m = pa[(k >> 1) * S];
        acc2 = r_fma(m, m, acc2);
    }
}
template <typename R, int S>
__device__ __noinline__ void jop_pow(R *__restrict__ w, const R *__restrict__ rk, uint32_t k, uint32_t a, uint32_t c, uint32_t inv,
                                     double alpha, int is_sqrt)
{
    const R *pa = w + a * S;
    R *pc = w + c * S;
    if (k == 0) {
        w[inv * S] = (R)1 / pa[0];
        pc[0] = is_sqrt ? r_sqrt(pa[0]) : pow0<R>(pa[0], alpha);
    } else {
        const R al1 = (R)(alpha + 1.0), kal = (R)k * (R)alpha;
        pc[k * S] = (conv_pow<R, S>(pa + k * S, pc, (int)k, al1, kal) * rk[k]) * w[inv * S];
    }
}
template <typename R, int S> __device__ __noinline__ void jop_exp(R *__restrict__ w, const R *__restrict__ rk, uint32_t k, uint32_t a, uint32_t c)
{
    const R *pa = w + a * S;
    R *pc = w + c * S;
    if (k == 0) {
        pc[0] = r_exp(pa[0]);
    } else {
        R acc = 0, jr = 1;
#pragma unroll 1
        for (uint32_t j = 1; j <= k; ++j, jr += (R)1) acc = r_fma(jr * pa[j * S], pc[(k - j) * S], acc);
        pc[k * S] = acc * rk[k];
    }
}
template <typename R, int S>
__device__ __noinline__ void jop_log(R *__restrict__ w, const R *__restrict__ rk, uint32_t k, uint32_t a, uint32_t c, uint32_t inv)
{
    const R *pa = w + a * S;
    R *pc = w + c * S;
    if (k == 0) {
        w[inv * S] = (R)1 / pa[0];
        pc[0] = r_log(pa[0]);
    } else {
        R acc = 0, jr = 1;
#pragma unroll 1
        for (uint32_t j = 1; j < k; ++j, jr += (R)1) acc = r_fma(jr * pc[j * S], pa[(k - j) * S], acc);
        pc[k * S] = r_fma(-acc, rk[k], pa[k * S]) * w[inv * S];
    }
}
template <typename R, int S>
__device__ __noinline__ void jop_intg(R *__restrict__ w, const R *__restrict__ rk, uint32_t k, uint32_t a, uint32_t b, uint32_t dst,
                                      int code)
{
    const R *pa = w + a * S, *pb = w + b * S;
    if (k == 0) {
        w[dst * S] = r_intg0(pa[0], code);
    } else {
        R acc = 0, jr = 1;
#pragma unroll 1
        for (uint32_t j = 1; j <= k; ++j, jr += (R)1) acc = r_fma(jr * pa[j * S], pb[(k - j) * S], acc);
        w[dst * S] = acc * rk[k];
    }
}
template <typename R, int S>
__device__ __noinline__ void jop_sincos(R *__restrict__ w, const R *__restrict__ rk, uint32_t k, uint32_t a, uint32_t sr, uint32_t cr)
{
    const R *pa = w + a * S;
    R *s = w + sr * S, *c = w + cr * S;
    if (k == 0) {
        R sv, cv;
        r_sincos(pa[0], &sv, &cv);
        s[0] = sv;
        c[0] = cv;
    } else {
        R sa0 = 0, ca0 = 0, sa1 = 0, ca1 = 0, jr = 1;
        uint32_t j = 1;
#pragma unroll 1
        for (; j + 1 <= k; j += 2, jr += (R)2) {
            const R ja0 = jr * pa[j * S], ja1 = (jr + (R)1) * pa[(j + 1) * S];
            const R c0 = c[(k - j) * S], s0 = s[(k - j) * S], c1 = c[(k - j - 1) * S], s1 = s[(k - j - 1) * S];
            sa0 = r_fma(ja0, c0, sa0);
            ca0 = r_fma(ja0, s0, ca0);
            sa1 = r_fma(ja1, c1, sa1);
            ca1 = r_fma(ja1, s1, ca1);
        }
        if (j <= k) {
            const R ja = jr * pa[j * S];
            sa0 = r_fma(ja, c[(k - j) * S], sa0);
            ca0 = r_fma(ja, s[(k - j) * S], ca0);
        }
        s[k * S] = (sa0 + sa1) * rk[k];
        c[k * S] = -((ca0 + ca1) * rk[k]);
    }
}
#endif

// ---------------------------------------------------------------------------
// Order-specialised body of the fused pair interaction (nd = 3): K is a
// compile-time constant, every loop is fully unrolled, every operand is an
// immediate-offset load from four base registers, nothing is predicated.
// The arithmetic (term order within each chain apart) is the generic path's.
// ---------------------------------------------------------------------------
template <typename R, int K>
__device__ __forceinline__ void pair3_k(R *__restrict__ d0, R *__restrict__ d1, R *__restrict__ d2, R *__restrict__ r2,
                                        R *__restrict__ c, R *__restrict__ inv, const R dk0, const R dk1, const R dk2,
                                        const R rkK, const double alpha, R &t0, R &t1, R &t2)
{
    d0[K] = dk0;
    d1[K] = dk1;
    d2[K] = dk2;
    constexpr int half = (K + 1) / 2;
    // Term order: every chain adds the term that involves the newest value (d[K], r2[K],
    // c[K]) LAST.  hy_nbody_reg.cuh uses the same order: the two paths agree bit for bit.
    // r2[K] = 2 * sum_i sum_{j<half} d_i[j] d_i[K-j]  (+ sum_i d_i[K/2]^2 for even K)
    R q0 = 0, q1 = 0, q2 = 0;
    if (half > 1) {
        q0 = d0[1] * d0[K - 1];
        q1 = d1[1] * d1[K - 1];
        q2 = d2[1] * d2[K - 1];
    }
#pragma unroll
    for (int j = 2; j < half; ++j) {
        q0 = r_fma(d0[j], d0[K - j], q0);
        q1 = r_fma(d1[j], d1[K - j], q1);
        q2 = r_fma(d2[j], d2[K - j], q2);
    }
    R e = 0;
    if ((K & 1) == 0 && K > 0) {
        const R m0 = d0[K / 2], m1 = d1[K / 2], m2 = d2[K / 2];
        e = r_fma(m2, m2, r_fma(m1, m1, m0 * m0));
    }
    if (half > 0) {
        q0 = r_fma(d0[0], dk0, q0);
        q1 = r_fma(d1[0], dk1, q1);
        q2 = r_fma(d2[0], dk2, q2);
    }
    R acc = (q0 + q1) + q2;
    acc = acc + acc;
    if (K == 0) e = r_fma(dk2, dk2, r_fma(dk1, dk1, dk0 * dk0));
    if ((K & 1) == 0) acc += e;
    r2[K] = acc;
    R ck;
    if (K == 0) {
        *inv = (R)1 / acc;
        ck = pow0<R>(acc, alpha);
    } else {
        const R al1 = (R)(alpha + 1.0), kal = (R)K * (R)alpha;
        R s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
        for (int j = 1; j < K; ++j) {
            const R wgt = r_fma((R)(-j), al1, kal);
            const R pr = wgt * r2[K - j];
            if ((j & 3) == 0) s0 = r_fma(pr, c[j], s0);
            if ((j & 3) == 1) s1 = r_fma(pr, c[j], s1);
            if ((j & 3) == 2) s2 = r_fma(pr, c[j], s2);
            if ((j & 3) == 3) s3 = r_fma(pr, c[j], s3);
        }
        const R tot = r_fma(kal * acc, c[0], (s0 + s1) + (s2 + s3));
        ck = (tot * rkK) * (*inv);
    }
    c[K] = ck;
    // t_i[K] = sum_{j<=K} d_i[j] c[K-j]; j = K uses d[K] from the register, j = 0 (newest c) goes last
    R a0 = 0, a1 = 0, a2 = 0, b0 = 0, b1 = 0, b2 = 0;
    if (K >= 1) {
        if (K & 1) {
            b0 = dk0 * c[0];
            b1 = dk1 * c[0];
            b2 = dk2 * c[0];
        } else {
            a0 = dk0 * c[0];
            a1 = dk1 * c[0];
            a2 = dk2 * c[0];
        }
    }
#pragma unroll
    for (int j = K - 1; j >= 1; --j) {
        const R cj = c[K - j];
        if (j & 1) {
            b0 = r_fma(d0[j], cj, b0);
            b1 = r_fma(d1[j], cj, b1);
            b2 = r_fma(d2[j], cj, b2);
        } else {
            a0 = r_fma(d0[j], cj, a0);
            a1 = r_fma(d1[j], cj, a1);
            a2 = r_fma(d2[j], cj, a2);
        }
    }
    if (K == 0) {
        t0 = dk0 * ck;
        t1 = dk1 * ck;
        t2 = dk2 * ck;
    } else {
        t0 = r_fma(d0[0], ck, a0 + b0);
        t1 = r_fma(d1[0], ck, a1 + b1);
        t2 = r_fma(d2[0], ck, a2 + b2);
    }
}

// One op of the program at order k on the trajectory column `w`.  `lt` is the
// lane's term stream (term c at lt[c * G]).
template <typename R, int G>
__device__ __forceinline__ void exec_op(const DOp o, const DTerm *__restrict__ lt, R *__restrict__ w,
                                        const R *__restrict__ rk, const double *__restrict__ s_imm, const uint32_t k,
                                        const R tm, R *__restrict__ gj, const uint32_t P1)
{
    // State recurrence store: x[k+1] = v.  Resident jets: dst + k + 1.  Spilled state
    // variables keep orders (k, k+1) on chip and stream the jet to the global scratch.
#define HY_STORE_SV(v)                                                                                  \
    do {                                                                                                \
        const R v_ = (v);                                                                               \
        if (o.pad & DP_DST) {                                                                           \
            w[o.dst + ((k + 1u) & 1u)] = v_;                                                            \
            gj[(uint32_t)(o.dst2 - 1) * P1 + k + 1] = v_;                                               \
        } else {                                                                                        \
            w[o.dst + k + 1] = v_;                                                                      \
        }                                                                                               \
    } while (0)
    switch (o.opcode) {
    case DOP_PAIR: {
        // Fused pair interaction (hy_schedule.hpp): d_i = +-A_i +- B_i, r2 = sum d_i^2,
        // wj = r2^alpha, t_i = d_i * wj.  Same arithmetic as the separate ops.
        const DTerm *t = lt + (uint32_t)o.b * G;
        const uint32_t nd = o.n;
        // all term records first (independent LDS.128), then all operands
        DTerm ta[3], tb[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const uint32_t ii = (uint32_t)i < nd ? (uint32_t)i : 0u;
            ta[i] = t[(2 * ii) * G];
            tb[i] = t[(2 * ii + 1) * G];
        }
        R av[3], bv[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            av[i] = w[roff(ta[i].src, k)];
            bv[i] = w[roff(ta[i].aux, k)];
        }
        R *d0 = w + rbase(tb[0].src), *d1 = w + rbase(tb[1].src), *d2 = w + rbase(tb[2].src);
        R dk[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int sc = (int)ta[i].coef;
            const R a = (sc & 1) ? -av[i] : av[i], b = (sc & 2) ? -bv[i] : bv[i];
            dk[i] = a + b;
        }
        if (nd == 3 && k < 24) {
            R t0, t1, t2;
            R *r2 = w + o.a, *c = w + o.dst, *inv = w + o.dst2;
            const double alpha = s_imm[o.imm];
            const R rkK = rk[k];
            switch (k) {
#define HY_PK(K) case K: pair3_k<R, K>(d0, d1, d2, r2, c, inv, dk[0], dk[1], dk[2], rkK, alpha, t0, t1, t2); break;
                HY_PK(0) HY_PK(1) HY_PK(2) HY_PK(3) HY_PK(4) HY_PK(5) HY_PK(6) HY_PK(7) HY_PK(8) HY_PK(9) HY_PK(10)
                HY_PK(11) HY_PK(12) HY_PK(13) HY_PK(14) HY_PK(15) HY_PK(16) HY_PK(17) HY_PK(18) HY_PK(19) HY_PK(20)
                HY_PK(21) HY_PK(22) HY_PK(23)
#undef HY_PK
            default: t0 = t1 = t2 = 0; break;
            }
            w[roff(tb[0].aux, k)] = t0;
            w[roff(tb[1].aux, k)] = t1;
            w[roff(tb[2].aux, k)] = t2;
            break;
        }
        d0[k] = dk[0];
        if (nd > 1) d1[k] = dk[1];
        if (nd > 2) d2[k] = dk[2];
        // r2[k] = sum_i ( 2 * sum_{j<half} d_i[j] d_i[k-j]  (+ d_i[k/2]^2 for even k) )
        // The j = 0 term uses d_i[k], which is still in a register.
        const int half = (int)((k + 1) >> 1);
        R q0 = 0, q1 = 0, q2 = 0;
        if (half > 0) {
            q0 = d0[0] * dk[0];
            if (nd > 1) q1 = d1[0] * dk[1];
            if (nd > 2) q2 = d2[0] * dk[2];
        }
#pragma unroll 1
        for (int j0 = 1; j0 < half; j0 += 4) {
            R x[4], xr[4], y[4], yr[4], z[4], zr[4];
            const int m = half - j0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool v = u < m;
                x[u] = v ? d0[j0 + u] : (R)0;
                xr[u] = v ? d0[k - (j0 + u)] : (R)0;
                y[u] = (v && nd > 1) ? d1[j0 + u] : (R)0;
                yr[u] = (v && nd > 1) ? d1[k - (j0 + u)] : (R)0;
                z[u] = (v && nd > 2) ? d2[j0 + u] : (R)0;
                zr[u] = (v && nd > 2) ? d2[k - (j0 + u)] : (R)0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                q0 = r_fma(x[u], xr[u], q0);
                q1 = r_fma(y[u], yr[u], q1);
                q2 = r_fma(z[u], zr[u], q2);
            }
        }
        R acc = (q0 + q1) + q2;
        acc = acc + acc;
        if ((k & 1u) == 0) {
            const R m0 = k ? d0[k >> 1] : dk[0];
            R e = m0 * m0;
            if (nd > 1) {
                const R m1 = k ? d1[k >> 1] : dk[1];
                e = r_fma(m1, m1, e);
            }
            if (nd > 2) {
                const R m2 = k ? d2[k >> 1] : dk[2];
                e = r_fma(m2, m2, e);
            }
            acc += e;
        }
        R *r2 = w + o.a, *c = w + o.dst;
        r2[k] = acc;
        const double alpha = s_imm[o.imm];
        R ck;
        if (k == 0) {
            w[o.dst2] = (R)1 / acc;
            ck = pow0<R>(acc, alpha);
        } else {
            const R al1 = (R)(alpha + 1.0), kal = (R)k * (R)alpha;
            // j = 0 term uses r2[k] = acc from the register
            R s = (kal * acc) * c[0];
            if (k > 1) s += conv_pow_from1<R>(r2 + k, c, (int)k, al1, kal);
            ck = (s * rk[k]) * w[o.dst2];
        }
        c[k] = ck;
        // t_i[k] = sum_{j<=k} d_i[j] c[k-j]; the j = 0 term uses c[k] from the register
        R s0 = d0[0] * ck, s1 = nd > 1 ? d1[0] * ck : (R)0, s2 = nd > 2 ? d2[0] * ck : (R)0;
        if (k > 0) {
            R u0, u1, u2;
            conv3<R>(d0 + 1, nd > 1 ? d1 + 1 : d0 + 1, nd > 2 ? d2 + 1 : d0 + 1, c + k - 1, (int)k, u0, u1, u2);
            s0 += u0;
            s1 += u1;
            s2 += u2;
        }
        w[roff(tb[0].aux, k)] = s0;
        if (nd > 1) w[roff(tb[1].aux, k)] = s1;
        if (nd > 2) w[roff(tb[2].aux, k)] = s2;
    } break;
    case HY_OP_LINCOMB: {
        // Specialised on the number of terms (exact, fully unrolled, nothing predicated):
        // all term records first, then all operands (and multipliers, if any term has a
        // runtime parameter), then one FMA chain in term order.
        const DTerm *t = lt + (uint32_t)o.b * G;
        const bool nopar = (o.pad & DP_NOPAR) != 0;
        R acc = 0;
#define HY_LC(N)                                                                                        \
    {                                                                                                   \
        DTerm tt[N];                                                                                    \
        _Pragma("unroll") for (int u = 0; u < N; ++u) tt[u] = t[u * G];                                 \
        R vv[N];                                                                                        \
        _Pragma("unroll") for (int u = 0; u < N; ++u) vv[u] = w[tt[u].src + (k & (tt[u].aux >> 16))];   \
        if (nopar) {                                                                                    \
            _Pragma("unroll") for (int u = 0; u < N; ++u) acc = r_fma((R)tt[u].coef, vv[u], acc);       \
        } else {                                                                                        \
            R mv[N];                                                                                    \
            _Pragma("unroll") for (int u = 0; u < N; ++u) mv[u] = w[tt[u].aux & 0xffffu];               \
            _Pragma("unroll") for (int u = 0; u < N; ++u) acc = r_fma((R)tt[u].coef * mv[u], vv[u], acc); \
        }                                                                                               \
    }
        uint32_t n = o.n;
        while (n > 8) { // long combinations: peel blocks of 8
            HY_LC(8)
            t += 8 * G;
            n -= 8;
        }
        switch (n) {
        case 1: HY_LC(1) break;
        case 2: HY_LC(2) break;
        case 3: HY_LC(3) break;
        case 4: HY_LC(4) break;
        case 5: HY_LC(5) break;
        case 6: HY_LC(6) break;
        case 7: HY_LC(7) break;
        case 8: HY_LC(8) break;
        default: break;
        }
#undef HY_LC
        if (o.flags & HY_OPF_SVD)
            HY_STORE_SV(acc * rk[k + 1]);
        else
            w[o.dst + ((o.flags & DF_JDST) ? k : 0u)] = acc;
    } break;
    case HY_OP_ADDSUB: {
        R a = w[ooff(o.a, o.flags & DF_JA, o.pad & DP_A, k)], b = w[ooff(o.b, o.flags & DF_JB, o.pad & DP_B, k)];
        if (o.flags & HY_OPF_NEGA) a = -a;
        if (o.flags & HY_OPF_NEGB) b = -b;
        const R acc = a + b;
        if (o.flags & HY_OPF_SVD)
            HY_STORE_SV(acc * rk[k + 1]);
        else
            w[o.dst + ((o.flags & DF_JDST) ? k : 0u)] = acc;
    } break;
    case HY_OP_SVD: {
        HY_STORE_SV(w[ooff(o.a, o.flags & DF_JA, o.pad & DP_A, k)] * rk[k + 1]);
    } break;
    case HY_OP_MUL: {
        w[o.dst + ((o.flags & DF_JDST) ? k : 0u)] = conv<R>(w + o.a, w + o.b + k, (int)k + 1);
    } break;
    case HY_OP_SQUARE: {
        const R *a = w + o.a;
        R acc = conv<R>(a, a + k, (int)((k + 1) >> 1));
        acc = acc + acc;
        if ((k & 1u) == 0) {
            const R m = a[k >> 1];
            acc = r_fma(m, m, acc);
        }
        w[o.dst + ((o.flags & DF_JDST) ? k : 0u)] = acc;
    } break;
    case HY_OP_SUMSQ: {
        const int half = (int)((k + 1) >> 1);
        const DTerm *t = lt + (uint32_t)o.b * G;
        const uint32_t n = o.n;
        R acc = 0, acc2 = 0;
#pragma unroll 1
        for (uint32_t i = 0; i < n; ++i) {
            const R *a = w + rbase(t[i * G].src);
            acc += conv<R>(a, a + k, half);
            if ((k & 1u) == 0) {
                const R m = a[k >> 1];
                acc2 = r_fma(m, m, acc2);
            }
        }
        w[o.dst + ((o.flags & DF_JDST) ? k : 0u)] = (acc + acc) + acc2;
    } break;
    case HY_OP_MULSH: {
        const R *b = w + o.a + k;
        const DTerm *t = lt + (uint32_t)o.b * G;
        const int n = (int)k + 1;
        if (o.n == 3) {
            const DTerm t0 = t[0], t1 = t[G], t2 = t[2 * G];
            R s0, s1, s2;
            conv3<R>(w + rbase(t0.src), w + rbase(t1.src), w + rbase(t2.src), b, n, s0, s1, s2);
            w[roff(t0.aux, k)] = s0;
            w[roff(t1.aux, k)] = s1;
            w[roff(t2.aux, k)] = s2;
        } else {
#pragma unroll 1
            for (uint32_t i = 0; i < o.n; ++i) {
                const DTerm ti = t[i * G];
                w[roff(ti.aux, k)] = conv<R>(w + rbase(ti.src), b, n);
            }
        }
    } break;
    case HY_OP_DIV: {
        const R *b = w + o.b;
        R *c = w + o.dst;
        if (k == 0) w[o.dst2] = (R)1 / b[0];
        R acc = w[ooff(o.a, o.flags & DF_JA, o.pad & DP_A, k)];
        if (k > 0) acc -= conv<R>(b + 1, c + k - 1, (int)k);
        c[k] = acc * w[o.dst2];
    } break;
    case HY_OP_POW:
    case HY_OP_SQRT: {
        const R *a = w + o.a;
        R *c = w + o.dst;
        const double alpha = o.opcode == HY_OP_SQRT ? 0.5 : s_imm[o.imm];
        if (k == 0) {
            w[o.dst2] = (R)1 / a[0];
            c[0] = o.opcode == HY_OP_SQRT ? r_sqrt(a[0]) : pow0<R>(a[0], alpha);
        } else {
            const R al1 = (R)(alpha + 1.0), kal = (R)k * (R)alpha;
            c[k] = (conv_pow<R>(a + k, c, (int)k, al1, kal) * rk[k]) * w[o.dst2];
        }
    } break;
    case HY_OP_EXP: {
        const R *a = w + o.a;
        R *c = w + o.dst;
        if (k == 0) {
            c[0] = r_exp(a[0]);
        } else {
            R acc = 0, jr = 1;
#pragma unroll 1
            for (uint32_t j = 1; j <= k; ++j, jr += (R)1) acc = r_fma(jr * a[j], c[k - j], acc);
            c[k] = acc * rk[k];
        }
    } break;
    case HY_OP_LOG: {
        const R *a = w + o.a;
        R *c = w + o.dst;
        if (k == 0) {
            w[o.dst2] = (R)1 / a[0];
            c[0] = r_log(a[0]);
        } else {
            R acc = 0, jr = 1;
#pragma unroll 1
            for (uint32_t j = 1; j < k; ++j, jr += (R)1) acc = r_fma(jr * c[j], a[k - j], acc);
            c[k] = r_fma(-acc, rk[k], a[k]) * w[o.dst2];
        }
    } break;
    case HY_OP_INTG: {
        // dst = F(a), dF/da = b:  dst[k] = (1/k) sum_{j=1..k} j a[j] b[k-j]
        const R *a = w + o.a, *b = w + o.b;
        if (k == 0) {
            w[o.dst] = r_intg0(a[0], (int)s_imm[o.imm]);
        } else {
            R acc = 0, jr = 1;
#pragma unroll 1
            for (uint32_t j = 1; j <= k; ++j, jr += (R)1) acc = r_fma(jr * a[j], b[k - j], acc);
            w[o.dst + ((o.flags & DF_JDST) ? k : 0u)] = acc * rk[k];
        }
    } break;
    case HY_OP_SINCOS: {
        const R *a = w + o.a;
        R *s = w + o.dst, *c = w + o.dst2;
        if (k == 0) {
            R sv, cv;
            r_sincos(a[0], &sv, &cv);
            s[0] = sv;
            c[0] = cv;
        } else {
            R sa0 = 0, ca0 = 0, sa1 = 0, ca1 = 0, jr = 1;
            uint32_t j = 1;
#pragma unroll 1
            for (; j + 1 <= k; j += 2, jr += (R)2) {
                const R ja0 = jr * a[j], ja1 = (jr + (R)1) * a[j + 1];
                const R c0 = c[k - j], s0 = s[k - j], c1 = c[k - j - 1], s1 = s[k - j - 1];
                sa0 = r_fma(ja0, c0, sa0);
                ca0 = r_fma(ja0, s0, ca0);
                sa1 = r_fma(ja1, c1, sa1);
                ca1 = r_fma(ja1, s1, ca1);
            }
            if (j <= k) {
                const R ja = jr * a[j];
                sa0 = r_fma(ja, c[k - j], sa0);
                ca0 = r_fma(ja, s[k - j], ca0);
            }
            s[k] = (sa0 + sa1) * rk[k];
            c[k] = -((ca0 + ca1) * rk[k]);
        }
    } break;
    case HY_OP_TIME: {
        w[o.dst + k] = k == 0 ? tm : (k == 1 ? (R)1 : (R)0);
    } break;
    default: break; // OP_NOP
    }
#undef HY_STORE_SV
}

// Error-free time arithmetic.  __dadd_rn & co. forbid
// contraction/reassociation.
__device__ __forceinline__ double ef_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float ef_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double ef_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float ef_sub(float a, float b) { return __fsub_rn(a, b); }

template <typename R> __device__ __forceinline__ void time_add(R &hi, R &lo, R h)
{
    R s = ef_add(hi, h);
    R bb = ef_sub(s, hi);
    R err = ef_add(ef_sub(hi, ef_sub(s, bb)), ef_sub(h, bb));
    err = ef_add(err, lo);
    R nh = ef_add(s, err);
    R nl = ef_sub(err, ef_sub(nh, s));
    hi = nh;
    lo = nl;
}
template <typename R> __device__ __forceinline__ R time_sub(R ahi, R alo, R bhi, R blo)
{
    R s = ef_sub(ahi, bhi);
    R bb = ef_sub(s, ahi);
    R err = ef_add(ef_sub(ahi, ef_sub(s, bb)), ef_sub(-bhi, bb));
    err = ef_add(err, ef_sub(alo, blo));
    return ef_add(s, err);
}

// Shared-memory carve-up (dynamic smem):
//   [ops | terms | imm | phase_slot | ev_ref | rk | ws]
struct SmemLayout {
    uint32_t off_ops, off_terms, off_imm, off_phase, off_ev, off_srow, off_ssp, off_ns0, off_rk, off_evt, off_ws, total;
};

__host__ __device__ inline uint32_t align_up(uint32_t x, uint32_t a) { return (x + a - 1) / a * a; }

// Elements of one trajectory's workspace column (before padding to odd).

__host__ __device__ inline SmemLayout make_layout(const hy_dims &d, const ProgDims &pd, uint32_t G, uint32_t T,
                                                  uint32_t RS, uint32_t real_bytes, int ws_in_smem)
{
    SmemLayout L;
    uint32_t o = 0;
    L.off_ops = o;
    o += pd.n_slots * G * 16u;
    L.off_terms = o;
    o += pd.n_tslots * G * 16u;
    L.off_imm = o;
    o += pd.n_imm * 8u;
    L.off_phase = o;
    o += (pd.n_phases + 1) * 4u;
    L.off_ev = o;
    o += d.n_events * 4u;
    L.off_srow = o;
    o += d.n_state * 4u;
    L.off_ssp = o;
    o += d.n_state * 4u;
    L.off_ns0 = o;
    o += T * 4u; // step count of every resident trajectory at the start of the launch
    o = align_up(o, 8);
    L.off_rk = o;
    o += (d.order + 2) * real_bytes;
    o = align_up(o, 16);
    L.off_evt = o;
    o += pd.evt_bytes;
    o = align_up(o, 16);
    L.off_ws = o;
    if (ws_in_smem) o += T * RS * real_bytes;
    L.total = o;
    return L;
}

// Threads per CTA (= the register budget: 65536 / threads).  The interpreter's small-group
// variants (G < 16: small systems, short jets) trade registers for resident trajectories;
// G = 16 and the register-resident N-body kernels need all 255 registers.
#ifndef HY_CRB_F32_THREADS
#define HY_CRB_F32_THREADS 768
#endif
// (rb: bytes per real.  The FP32 CR3BP kernel holds 5 x 9 jet registers: more warps instead.)
__host__ __device__ constexpr int hy_max_threads(int G, bool smem, int NB, int rb = 8)
{
#ifdef HY_JIT_THREADS
    return HY_JIT_THREADS; // run-time compiled kernels: the CTA size is a compile-time constant of the build
#endif
    return (NB < 0 && rb == 4) ? HY_CRB_F32_THREADS : ((NB == 0 && smem && G < 16) ? 512 : 256);
}

// NB > 0: register-resident jets for a matched N-body tape (hy_nbody_reg.cuh); NB < 0: for the
// CR3BP tape (hy_cr3bp_reg.cuh, G = 2); the tape interpreter is not instantiated.  NB = 0: tape
// interpreter.  NB != 0: the trajectories of a warp step in lockstep.
// WGX: warpgroup rotation (hy_nbody_reg.cuh): 384 threads, 24 trajectories, registers traded
// between the warpgroups at the phase boundaries of the step.
// PM: the order the register-resident N-body jets are unrolled to (NBR_PMAX, or NBR_LMAX for the
// high-accuracy 6-body build).
// FX: extended features - active-lane mask, resumed launches, per-launch step budget / pause on
// non-terminal events, the continuous-output recorder, the device-side angle reducer, events on
// the register-resident kernels.  The plain build (FX = false) of the register-resident kernels
// is what an uninterrupted propagate_for/until/grid or step() runs: it carries none of that code.
template <typename R, int G, bool SMEM, int NB = 0, bool WGX = false, int PM = NBR_PMAX, bool FX = (NB == 0)>
__global__ void __launch_bounds__(WGX ? 384 : hy_max_threads(G, SMEM, NB, (int)sizeof(R)), 1) propagate_kernel(const KParams<R> P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const hy_dims &d = P.d;
    const SmemLayout L = make_layout(d, P.pd, G, P.T, P.TS, sizeof(R), SMEM ? 1 : 0);
    DOp *s_ops = reinterpret_cast<DOp *>(smem_raw + L.off_ops);
    DTerm *s_terms = reinterpret_cast<DTerm *>(smem_raw + L.off_terms);
    double *s_imm = reinterpret_cast<double *>(smem_raw + L.off_imm);
    uint32_t *s_phase = reinterpret_cast<uint32_t *>(smem_raw + L.off_phase);
    uint32_t *s_ev = reinterpret_cast<uint32_t *>(smem_raw + L.off_ev);
    uint32_t *s_srow = reinterpret_cast<uint32_t *>(smem_raw + L.off_srow);
    int32_t *s_ssp = reinterpret_cast<int32_t *>(smem_raw + L.off_ssp);
    uint32_t *s_ns0 = reinterpret_cast<uint32_t *>(smem_raw + L.off_ns0);
    R *s_rk = reinterpret_cast<R *>(smem_raw + L.off_rk);
    const uint32_t RS = P.TS; // workspace stride between trajectories (odd)

    // ---- stage the program ----
    {
        const uint32_t nw = L.off_phase / 4; // ops + terms + imm are contiguous
        const uint32_t *src = reinterpret_cast<const uint32_t *>(P.prog);
        uint32_t *dst = reinterpret_cast<uint32_t *>(smem_raw);
        for (uint32_t i = threadIdx.x; i < nw; i += blockDim.x) dst[i] = src[i];
        for (uint32_t i = threadIdx.x; i <= P.pd.n_phases; i += blockDim.x) s_phase[i] = P.phase_slot[i];
        for (uint32_t i = threadIdx.x; i < d.n_events; i += blockDim.x) s_ev[i] = P.ev_ref[i];
        for (uint32_t i = threadIdx.x; i < d.n_state; i += blockDim.x) {
            s_srow[i] = P.state_row[i];
            s_ssp[i] = P.state_spill[i];
        }
        for (uint32_t i = threadIdx.x; i < d.order + 2; i += blockDim.x)
            s_rk[i] = i == 0 ? (R)0 : (R)(1.0 / (double)i);
        if constexpr (FX && NB != 0) {
            const uint32_t *es = reinterpret_cast<const uint32_t *>(P.evt.blob);
            uint32_t *ed = reinterpret_cast<uint32_t *>(smem_raw + L.off_evt);
            for (uint32_t i = threadIdx.x; i < P.pd.evt_bytes / 4; i += blockDim.x) ed[i] = es[i];
        }
    }
    __syncthreads();
    // event tape of the register-resident kernels (FX builds): ops | terms | imm | op ranges | event slots
    const EOp *s_eops = reinterpret_cast<const EOp *>(smem_raw + L.off_evt);
    const ETerm *s_eterms = reinterpret_cast<const ETerm *>(s_eops + P.evt.n_ops);
    const double *s_eimm = reinterpret_cast<const double *>(s_eterms + P.evt.n_terms);
    const uint32_t *s_estart = reinterpret_cast<const uint32_t *>(s_eimm + P.evt.n_imm);
    const uint32_t *s_eslot = s_estart + P.evt.n_events + 1;
    const uint32_t *s_eused = s_eslot + P.evt.n_events;
    const bool reg_events = FX && NB != 0 && P.evt.n_events != 0;

    const uint32_t p = d.order, P1 = p + 1, n = d.n_state;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t sub = threadIdx.x & (G - 1);
    const uint32_t slot = threadIdx.x / G; // trajectory slot in this CTA
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(uint32_t)(G - 1)));
    const uint32_t par_off = P.pd.par_off, one_off = P.pd.one_off;
    // whole groups only: safe w.r.t. group-mask syncs (NB != 0: T is such that whole warps run)
    // (wmask: the lanes of this warp that own a trajectory slot - the interpreter's warp-level syncs)
    const unsigned wmask = __ballot_sync(0xffffffffu, slot < P.T);
    if (slot >= P.T) return;
    R *w;
#ifdef HY_JIT
    // warp-interleaved workspace: warp q of the CTA owns rows [q * RS, (q + 1) * RS) x 32 lanes
    w = (SMEM ? reinterpret_cast<R *>(smem_raw + L.off_ws) : P.gws + (size_t)blockIdx.x * P.T * RS) +
        (size_t)(threadIdx.x >> 5) * RS * 32u + lane;
#else
    if (SMEM)
        w = reinterpret_cast<R *>(smem_raw + L.off_ws) + (size_t)slot * RS;
    else
        w = P.gws + ((size_t)blockIdx.x * P.T + slot) * RS;
#endif
    // global scratch of this trajectory slot's spilled state jets
    R *gj = P.gjet + ((size_t)blockIdx.x * P.T + slot) * (size_t)P.pd.n_spill * P1;
    // event workspace of the register-resident kernels (hy_evtape.cuh): inside the column, or a slab in
    // global memory; wev: the base that the event-jet references (s_ev, relative to the column) apply to
    R *const ews = P.evt.gws ? reinterpret_cast<R *>(P.evt.gws) + ((size_t)blockIdx.x * P.T + slot) * P.evt.gstride
                             : w + P.evt.ews_off;
    R *const wev = ews - P.evt.ews_off;
    // order j of state variable i (resident jet or spilled copy; .cg: written by other lanes of the group)
// (NB > 0: the orders of a state variable are NBR_JS elements apart, see hy_nbody_reg.cuh)
    // (HY_JIT: the rows of the interleaved workspace are HY_WS elements apart; state_row, ev_ref, par_off
    //  and one_off arrive pre-multiplied)
    constexpr uint32_t XS = NB > 0 ? (uint32_t)NBR_JS : (NB < 0 ? (uint32_t)CRB_XS : (uint32_t)HY_WS);
    constexpr uint32_t ES = NB == 0 ? (uint32_t)HY_WS : 1u; // order stride of an event jet
#define XJ(i, j) (s_ssp[i] >= 0 ? __ldcg(&gj[(uint32_t)s_ssp[i] * P1 + (j)]) : w[s_srow[i] + (j) * XS])
    // unit jet [1, 0, ..., 0] (never changes)
    if constexpr (NB == 0)
        for (uint32_t i = sub; i < P1; i += G) w[one_off + i * XS] = i == 0 ? (R)1 : (R)0;
    // register-resident N-body path: per-lane constants (pair, exchange slots, body)
    NbrLane<(NB > 0 ? NB : 2)> nl{};
    if constexpr (NB > 0) {
        // lane record (hy_nbody_match.hpp): bodies a, b of the lane's pair
        const uint2 lr = *reinterpret_cast<const uint2 *>(s_imm + NBR_LANE0 + sub);
        nl.xa = (int32_t)(NBR_BS * lr.x);
        nl.xb = (int32_t)(NBR_BS * lr.y);
        nl.ta = (int32_t)(NBR_TB0 + nbr_ts(WGX) * sub);
        // Body lanes: lanes of the FIRST half-warp serve the NB bodies of the warp's two trajectories
        // (a 64-bit shared access costs one wavefront per active half-warp).
        // (trajectory 0: lanes 0..NB-1, trajectory 1: lanes 8..8+NB-1 - one quarter-warp each, so the
        // 128-bit accesses of the two trajectories never meet in one wavefront)
        nl.body = G == 32 ? lane < (uint32_t)NB : (lane < 16u && (lane & 7u) < (uint32_t)NB);
        // (every lane gets valid body addresses - the loads of the body phase are not predicated:
        // lanes 16..31 mirror lanes 0..15, the spare lanes of a quarter-warp mirror its first bodies)
        // (G = 32: one trajectory per warp - its bodies on lanes 0..NB-1, the other lanes mirror them)
        const uint32_t bt = G == 32 ? 0u : (lane >> 3) & 1u, bd = (G == 32 ? lane : (lane & 7u)) % (uint32_t)NB;
        const int32_t col = (G != 32 && bt != (lane >> 4)) ? (bt ? (int32_t)RS : -(int32_t)RS) : 0; // the other trajectory's column
        nl.xbody = col + (int32_t)(NBR_BS * bd);
        nl.coef = (int32_t)(NBR_CS * bd);
#pragma unroll
        for (int q = 0; q < NB - 1; ++q)
#ifndef HY_NBR_SMEM_EXCHANGE
            // source lane of term q: the pair lanes of trajectory bt are lanes 16 bt .. 16 bt + 14
            nl.tin[q] = (int32_t)(16u * bt) + (int32_t) * reinterpret_cast<const uint32_t *>(s_imm + NBR_OFF0 + NBR_CS * bd + q);
#else
            nl.tin[q] = col + NBR_TB0 +
                        nbr_ts(WGX) * (int32_t) * reinterpret_cast<const uint32_t *>(s_imm + NBR_OFF0 + NBR_CS * bd + q);
#endif
    }

    // register-resident CR3BP path: per-lane constants
    CrbLane<R> cl{};
    if constexpr (NB < 0) {
        cl.sub = sub != 0u;
        cl.soff = 3 * (int32_t)sub;
        cl.c = (R)s_imm[sub];
        cl.m = (R)s_imm[2u + sub];
        cl.ga = (R)s_imm[4u + 2u * sub];
        cl.gb = (R)s_imm[5u + 2u * sub];
    }
    const uint32_t wg = threadIdx.x >> 7; // warpgroup (WGX)
    if constexpr (WGX) {
        // every warpgroup starts in the tail state; the first two to reach their jets get the registers
        wg_bar(wg);
        wg_reg_dec<NBR_WGX_TREG>();
    }
    // Persistent loop: every iteration is ONE step of the group's current trajectory (or the
    // fetch of a new one).  NB > 0: the two trajectories of a warp step in lockstep (the jets are a
    // warp-wide phase); a half-warp without a live trajectory idles through it on stale data.
#ifdef HY_WGX_PROF
    long long prof_wait = 0, prof_jets = 0, prof_rel = 0, prof_n = 0;
    const long long prof_t0 = clock64();
    long long prof_c = 0; // HY_TAIL_PROF: "wait" = norms + step size, "release" = state update
#endif
    bool have = false;
    unsigned int traj = 0;
    R hi = 0, lo = 0, mdt = 0, tf_hi = 0, tf_lo = 0;
    uint32_t gi = 0; // next grid point to emit (MODE_GRID)
    uint32_t cc = 0; // recorded continuous-output steps
    bool cd_live = true; // (register kernels with events) a terminal-event cooldown may be running
    // (recorder state other than the step count `cc` is NOT kept in registers across the jets: the id of
    //  the lane's last chunk is re-read from P.rec.tail once per recorded step; the per-launch step
    //  budget is checked against the step count at fetch time, kept in shared memory)
    long long oc = HY_OUTCOME_TIME_LIMIT;
    R mn = r_inf<R>(), mx = 0, h = 0;
    unsigned long long ns = 0;
    for (;;) {
#ifndef HY_INTERP_NO_LOCKSTEP
        // Interpreter: every group of the warp runs the SAME program, so the groups stay converged as
        // long as they start their steps together.  Ragged step counts would let a group that fetched
        // a new trajectory drift out of phase (each group then issues its own copy of the op stream:
        // 3x slower on config 5) - re-align the warp at the top of every iteration.
        if constexpr (NB == 0 && G < 32) __syncwarp(wmask);
#endif
        if (!have) {
            // ---- fetch the next trajectory for this group ----
            if (sub == 0) traj = (WGX && wg >= P.wgx_wgs) ? P.B : atomicAdd(P.counter, 1u);
            if (G > 1) traj = __shfl_sync(gmask, traj, 0, G);
            if (traj < P.B && (!FX || !P.active || P.active[traj])) {
                have = true;
                for (uint32_t i = sub; i < n; i += G) {
                    const R x0 = P.state[(size_t)i * P.B + traj];
                    w[s_srow[i]] = x0;
                    if (s_ssp[i] >= 0) gj[(uint32_t)s_ssp[i] * P1] = x0;
                }
                // (register-resident kernels keep the parameters as unit-stride rows after the state jets)
                for (uint32_t i = sub; i < d.n_par; i += G)
                    w[par_off + i * (NB == 0 ? XS : 1u)] = P.pars[(size_t)i * P.B + traj];
                hi = P.t_hi[traj];
                lo = P.t_lo[traj];
                mdt = P.mdt ? P.mdt[traj] : r_inf<R>();
                tf_hi = 0;
                tf_lo = 0;
                if (P.mode == MODE_FOR || P.mode == MODE_UNTIL) {
                    // (absolute final times: written by prep_tf_kernel before the first launch)
                    tf_hi = P.tf_hi[traj];
                    tf_lo = P.tf_lo[traj];
                } else if (P.mode == MODE_GRID) {
                    tf_hi = P.grid[(size_t)(P.grid_k - 1) * P.B + traj];
                }
                gi = (FX && P.resume) ? P.gidx[traj] : 0u;
                cc = 0;
                cd_live = true;
                if (P.mode == MODE_GRID && !(FX && P.resume)) {
                    // Grid points at (or before) the starting time take the current state.
                    const R dir = tf_hi - hi;
                    while (gi < P.grid_k) {
                        const R g = P.grid[(size_t)gi * P.B + traj];
                        const R dg = (g - hi) - lo;
                        if ((dir >= (R)0 && dg > (R)0) || (dir < (R)0 && dg < (R)0)) break;
                        for (uint32_t i = sub; i < n; i += G)
                            P.gout[((size_t)gi * n + i) * P.B + traj] = w[s_srow[i]];
                        ++gi;
                    }
                }
                if ((FX && P.rec.on)) {
                    if (P.rec.append) {
                        cc = P.rec.count[traj];
                    } else if (sub == 0) {
                        P.rec.t0_hi[traj] = hi;
                        P.rec.t0_lo[traj] = lo;
                        P.rec.head[traj] = HY_REC_NONE;
                        P.rec.tail[traj] = HY_REC_NONE;
                    }
                }
                if (P.mode != MODE_STEP) mdt = r_abs(mdt);
                oc = HY_OUTCOME_TIME_LIMIT;
                mn = r_inf<R>();
                mx = 0;
                h = 0;
                ns = 0;
                if ((FX && P.resume)) {
                    mn = P.min_h[traj];
                    mx = P.max_h[traj];
                    ns = P.n_steps[traj];
                    h = P.last_h[traj];
                }
                if ((FX && P.launch_steps) && sub == 0) s_ns0[slot] = (uint32_t)ns; // step count at the start of the launch
                if (G > 1) __syncwarp(gmask);
            }
        }
        if constexpr (WGX) {
            if (!wg_any(wg, have)) break; // (warpgroup-wide vote: the four warps trade registers together)
        } else if constexpr (NB != 0) {
            if (!__any_sync(0xffffffffu, have)) break;
        } else {
#ifndef HY_INTERP_NO_LOCKSTEP
            if (!__any_sync(wmask, have)) break; // (groups without work idle until the whole warp is done)
#else
            if (!have) break;
#endif
        }

        bool fin = false; // the trajectory ends with this iteration
        R rem = 0, lim = 0;
        if (P.mode == MODE_STEP) {
            lim = P.mdt ? mdt : (P.backward ? -r_inf<R>() : r_inf<R>());
        } else {
            rem = time_sub(tf_hi, tf_lo, hi, lo);
            if (rem == (R)0) {
                oc = HY_OUTCOME_TIME_LIMIT;
                fin = true; // already there: nothing to integrate
            }
            lim = r_abs(rem) < mdt ? rem : r_copysign(mdt, rem);
        }
        // ---- recorder: the step needs a slot; take a new chunk from the pool when the lane's
        // chunks are full.  Pool exhausted: the lane stops (PAUSED) BEFORE the step; the host adds
        // a segment and resumes it.
        if ((FX && P.rec.on) && have && !fin && (cc % HY_REC_CH) == 0u) {
            // (a chunk is taken exactly when the step that fills its first slot is about to run, so
            //  "cc is a multiple of the chunk size" == "the lane's chunks are full")
            uint32_t cid = 0;
            if (sub == 0) cid = atomicAdd(P.rec.next, 1u);
            if (G > 1) cid = __shfl_sync(gmask, cid, 0, G);
            if (cid >= P.rec.cap_chunks) {
                oc = HY_OUTCOME_PAUSED_POOL;
                fin = true;
            } else if (sub == 0) {
                *reinterpret_cast<uint32_t *>(P.rec.chunk(cid)) = HY_REC_NONE;
                const uint32_t rtail = P.rec.tail[traj];
                if (rtail == HY_REC_NONE)
                    P.rec.head[traj] = cid;
                else
                    *reinterpret_cast<uint32_t *>(P.rec.chunk(rtail)) = cid;
                P.rec.tail[traj] = cid;
            }
            if (G > 1) __syncwarp(gmask); // (the group reads P.rec.tail below)
        }
        const bool stepping = have && !fin;

        if constexpr (FX && NB < 0 && sizeof(R) == 8) {
            // (the bulk copy of the previous step's record must have READ the column before the jets rewrite it)
            if (P.rec.on && P.rec.sk != 1u) {
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
            }
        }
        if (NB != 0 || stepping) {
            // ---- jets: orders 0..p-1 of every op (the state recurrence is part of the program) ----
            if constexpr (NB > 0) {
                if constexpr (WGX) {
                    // acquire the jet registers (blocks until another warpgroup has left its jets)
#ifdef HY_WGX_PROF
                    const long long c0 = clock64();
#endif
                    wg_bar(wg);
                    wg_reg_inc<NBR_WGX_JREG>();
#ifdef HY_WGX_PROF
                    const long long c1 = clock64();
#endif
                    nbr_jets<R, NB, PM, true, true>(w, s_imm + nl.coef, nl, p, par_off);
#ifdef HY_WGX_PROF
                    const long long c2 = clock64();
#endif
                    wg_bar(wg);
                    wg_reg_dec<NBR_WGX_TREG>();
#ifdef HY_WGX_PROF
                    prof_wait += c1 - c0;
                    prof_jets += c2 - c1;
                    prof_rel += clock64() - c2;
                    ++prof_n;
#endif
                } else if (p == (uint32_t)PM) {
#ifdef HY_WGX_PROF
                    const long long c1 = clock64();
#endif
                    nbr_jets<R, NB, PM, true>(w, s_imm + nl.coef, nl, p, par_off);
#ifdef HY_WGX_PROF
                    prof_c = clock64();
                    prof_jets += prof_c - c1;
                    ++prof_n;
#endif
                } else
                    nbr_jets<R, NB, PM, false>(w, s_imm + nl.coef, nl, p, par_off);
            } else if constexpr (NB < 0) {
                // (PM = NBR_PMAX: the standard builds; PM = CRB_PMAX_HI: the FP64 order-22 build)
                constexpr int CPM = PM == NBR_PMAX ? CrbPmax<R>::value : PM;
                if (p == (uint32_t)CPM)
                    crb_jets<R, CPM, true>(w, cl, p);
                else
                    crb_jets<R, CPM, false>(w, cl, p);
            } else {
#ifdef HY_JIT
                hy_gen_jets<R>(w, s_rk, hi);
                if (d.n_events) hy_gen_ev_sweep<R>(w, s_rk, hi);
#else
                const DOp *lops = s_ops + sub;
                const DTerm *lterms = s_terms + sub;
                const uint32_t n_ph = P.pd.n_phases;
                // (with events, one extra sweep at order p over the event-function ops only)
                const uint32_t k_end = d.n_events ? p + 1 : p;
#pragma unroll 1
                for (uint32_t k = 0; k < k_end; ++k) {
                    const bool ev_sweep = k == p;
#pragma unroll 1
                    for (uint32_t ph = 0; ph < n_ph; ++ph) {
                        const uint32_t e = s_phase[ph + 1];
#pragma unroll 1
                        for (uint32_t i = s_phase[ph]; i < e; ++i) {
                            const DOp o = lops[i * G];
                            if (ev_sweep &&
                                (!(o.flags & HY_OPF_EVENT) || (o.flags & HY_OPF_SVD) || o.opcode == HY_OP_SVD))
                                continue;
                            exec_op<R, G>(o, lterms, w, s_rk, s_imm, k, hi, gj, P1);
                        }
                        if (G > 1) __syncwarp(gmask);
                    }
                }
#endif
            }
            // ---- event functions on the register-resident kernels (hy_evtape.cuh): lane e mod G
            // evaluates event e from the state jets.  Linear ops at every order, products / squares
            // whose history nobody reads at orders 0, p-1, p (all the step-size norms need).
            if constexpr (FX && NB != 0) {
                if (reg_events) {
                    const EvtCtx<R, (int)XS> ec{w, s_srow, ews, s_rk, s_eimm, hi};
#ifdef HY_JIT_EVT
                    // (every lane of the group: the products are spread over the lanes; the interval scratch
                    //  is free here and carries their results to lane 0)
                    hy_gen_evt_norms<R, (int)XS>(ec, s_eterms, ews + (P.evt.eiv_off - P.evt.ews_off), sub);
#else
                    for (uint32_t e = sub; e < P.evt.n_events; e += G) {
                        const uint32_t o0 = s_estart[e], o1 = s_estart[e + 1];
                        // pass A: the ops that are needed at every order, op by op (they never read a
                        // three-order op, and an op's order k needs its operands at orders <= k only)
#pragma unroll 1
                        for (uint32_t i = o0; i < o1; ++i) {
                            const EOp o = s_eops[i];
                            if (o.flags & EOF_ALL) evt_exec_all<R, (int)XS>(o, s_eterms, ec, p);
                        }
                        // pass B: the rest at orders 0, p-1, p, order by order
#pragma unroll 1
                        for (uint32_t q = 0; q < 3; ++q) {
                            const uint32_t k = q == 0 ? 0u : p - 2u + q;
#pragma unroll 1
                            for (uint32_t i = o0; i < o1; ++i) {
                                const EOp o = s_eops[i];
                                if (!(o.flags & EOF_ALL)) evt_exec<R, (int)XS>(o, s_eterms, ec, k);
                            }
                        }
                    }
#endif
                    __syncwarp();
                }
            }
        }

        // ---- the rest of the step ----
        // NB > 0: BOTH half-warps run it, converged, so that every warp-level primitive uses the
        // full mask (a partial mask costs a ~90-clock MATCH/VOTE check per use); a half-warp that is
        // not stepping computes on stale data and commits nothing (`stepping` guards every side effect).
        if (NB != 0 || stepping) {
            constexpr unsigned TM_FULL = 0xffffffffu;
            const unsigned tmask = NB != 0 ? TM_FULL : gmask;
            // (recorder: the id of the lane's last chunk, asked for now - an L2 round trip - and needed only
            //  after the step size is known; not kept in a register across the jets)
            uint32_t rec_tail = HY_REC_NONE;
            if constexpr (FX) {
                if (P.rec.on && stepping) rec_tail = __ldcg(&P.rec.tail[traj]);
            }
            // ---- step size ----
            R n0 = 0, n1 = 0, n2 = 0;
            // NB > 0: lane `sub` < 2 NB owns one 3-vector of the state (position or velocity of a body;
            // the assignment keeps the 128-bit accesses of a quarter-warp conflict-free)
            // (NB = 6: positions of bodies 0-5 on lanes 0-5, velocities of bodies 4, 5 on lanes 6, 7 and of
            // bodies 0-3 on lanes 8-11 - conflict-free quarter-warps; otherwise positions first, then velocities)
            const uint32_t vb = NB == 6 ? (sub < 6u ? sub : (sub < 8u ? sub - 2u : sub - 8u))
                                        : (sub < (uint32_t)NB ? sub : sub - (uint32_t)NB);     // body
            const uint32_t voff = vb * (uint32_t)NBR_BS + (sub < (uint32_t)NB ? 0u : 4u);       // block offset
            const bool vown = NB > 0 && sub < 2u * NB && vb < (uint32_t)NB;
            if constexpr (NB > 0) {
                if (vown) {
                    Vec3<R> a, b, c;
                    lds3(w + voff, a);
                    lds3(w + voff + (p - 1) * NBR_JS, b);
                    lds3(w + voff + p * NBR_JS, c);
                    n0 = nan_max(nan_max(nan_max(n0, r_abs(a.x)), r_abs(a.y)), r_abs(a.z));
                    n1 = nan_max(nan_max(nan_max(n1, r_abs(b.x)), r_abs(b.y)), r_abs(b.z));
                    n2 = nan_max(nan_max(nan_max(n2, r_abs(c.x)), r_abs(c.y)), r_abs(c.z));
                }
                if constexpr (FX) {
                    if (reg_events) // the event functions take part in the norms
                        for (uint32_t e = sub; e < P.evt.n_events; e += G) {
                            const R *x = &wev[s_ev[e]];
                            n0 = nan_max(n0, r_abs(x[0]));
                            n1 = nan_max(n1, r_abs(x[p - 1]));
                            n2 = nan_max(n2, r_abs(x[p]));
                        }
                }
            } else {
                for (uint32_t i = sub; i < n + d.n_events; i += G) {
                    R x0, x1, x2;
                    if (i < n) {
                        x0 = XJ(i, 0);
                        x1 = XJ(i, p - 1);
                        x2 = XJ(i, p);
                    } else {
                        const R *x = &wev[s_ev[i - n]];
                        x0 = x[0];
                        x1 = x[(p - 1) * ES];
                        x2 = x[p * ES];
                    }
                    n0 = nan_max(n0, r_abs(x0));
                    n1 = nan_max(n1, r_abs(x1));
                    n2 = nan_max(n2, r_abs(x2));
                }
            }
#pragma unroll
            for (int m = G >> 1; m > 0; m >>= 1) {
                n0 = nan_max(n0, shfl_xor<R>(tmask, n0, m));
                n1 = nan_max(n1, shfl_xor<R>(tmask, n1, m));
                n2 = nan_max(n2, shfl_xor<R>(tmask, n2, m));
            }
            R hn; // the new step size
            {
                const bool isnan_ = n0 != n0 || n1 != n1 || n2 != n2;
                const R num = n0 < (R)1 ? (R)1 : n0;
                R rho_p, rho_pm1;
                if (G > 1) {
                    // one root call for both radii: even lanes take order p, odd lanes order p-1
                    const bool odd = sub & 1u;
                    const R r = r_root(num / (odd ? n1 : n2), odd ? P.inv_pm1 : P.inv_p);
                    const R o = shfl_xor<R>(tmask, r, 1);
                    rho_p = odd ? o : r;
                    rho_pm1 = odd ? r : o;
                } else {
                    rho_p = r_root(num / n2, P.inv_p);
                    rho_pm1 = r_root(num / n1, P.inv_pm1);
                }
                hn = (rho_p < rho_pm1 ? rho_p : rho_pm1) * P.rhofac;
                if (isnan_) hn = n0 + n1 + n2; // NaN
            }
            if (signbit(lim)) hn = -hn;
            long long so = HY_OUTCOME_SUCCESS;
            if (r_abs(hn) > r_abs(lim)) {
                hn = lim;
                so = HY_OUTCOME_TIME_LIMIT;
            }
            if (stepping) h = hn;
#if defined(HY_WGX_PROF) && defined(HY_TAIL_PROF)
            {
                const long long t_ = clock64();
                prof_wait += t_ - prof_c;
                prof_c = t_;
            }
#endif

            // ---- event detection: may truncate the step at a terminal event ----
            int term_ev = -1;
            int nt_fired = 0; // non-terminal events logged in this step
            if (NB == 0 && d.n_events) {
                R h_eff = h;
                if (sub == 0)
                    detect_events<R, (int)ES>(w, s_ev, d.n_events, d.n_tevents, (int)p, h, hi, lo, traj, ns, P.ev, h_eff,
                                              term_ev, nt_fired);
                if (G > 1) {
                    h_eff = __shfl_sync(gmask, h_eff, 0, G);
                    term_ev = __shfl_sync(gmask, term_ev, 0, G);
                    nt_fired = __shfl_sync(gmask, nt_fired, 0, G);
                }
                if (term_ev >= 0) {
                    h = h_eff;
                    so = -(long long)term_ev - 1;
                }
                if (sub == 0 && d.n_tevents) advance_cooldowns<R>(traj, d.n_tevents, h, P.ev);
            }

            if constexpr (FX && NB != 0) {
                if (reg_events) {
                    // ---- can an event happen in [0, h] at all?  Interval Horner enclosures of the state
                    // polynomials over the step, pushed through the event tape in interval arithmetic.
                    R *iv = ews + (P.evt.eiv_off - P.evt.ews_off);
                    bool maybe = false;
#ifdef HY_JIT_EVT
                    maybe = hy_gen_evt_interval<R, (int)XS>(w, iv, s_eterms, s_eimm, hi, h, sub);
#else
                    for (uint32_t i = sub; i < n; i += G) {
                        if (!s_eused[i]) continue; // (no event reads this state variable)
                        const Ival<R> v = iv_horner<R>(w + s_srow[i], (int)XS, (int)p, h);
                        iv[2 * i] = v.lo;
                        iv[2 * i + 1] = v.hi;
                    }
                    __syncwarp();
                    for (uint32_t e = sub; e < P.evt.n_events; e += G) {
                        const uint32_t o0 = s_estart[e], o1 = s_estart[e + 1];
#pragma unroll 1
                        for (uint32_t i = o0; i < o1; ++i) evt_interval<R>(s_eops[i], s_eterms, iv, s_eimm, hi, h);
                        const R glo = iv[2 * s_eslot[e]], ghi = iv[2 * s_eslot[e] + 1];
                        if (!(glo > (R)0 || ghi < (R)0)) maybe = true; // 0 inside the enclosure (or NaN)
                    }
#endif
                    const unsigned mb = __ballot_sync(TM_FULL, maybe && stepping);
                    const unsigned gbits = G == 32 ? 0xffffffffu : ((1u << G) - 1u);
                    const bool gmaybe = ((mb >> (lane & ~(uint32_t)(G - 1))) & gbits) != 0u;
                    if (P.evt.stats && sub == 0 && stepping) {
                        atomicAdd(P.evt.stats, 1ULL);
                        if (gmaybe) atomicAdd(P.evt.stats + 1, 1ULL);
                    }
                    if (gmaybe) {
                        // ---- rare: the remaining orders of the event jets, then the root finder ----
                        const EvtCtx<R, (int)XS> ec{w, s_srow, ews, s_rk, s_eimm, hi};
#ifdef HY_JIT_EVT
                        if (sub == 0) {
#pragma unroll 1
                            for (uint32_t k = 1; k + 1 < p; ++k) hy_gen_evt_order<R, (int)XS>(ec, s_eterms, k);
                        }
#else
                        for (uint32_t e = sub; e < P.evt.n_events; e += G) {
                            const uint32_t o0 = s_estart[e], o1 = s_estart[e + 1];
#pragma unroll 1
                            for (uint32_t k = 1; k + 1 < p; ++k)
#pragma unroll 1
                                for (uint32_t i = o0; i < o1; ++i) {
                                    const EOp o = s_eops[i];
                                    if (!(o.flags & EOF_ALL)) evt_exec<R, (int)XS>(o, s_eterms, ec, k);
                                }
                        }
#endif
                        __syncwarp(gmask);
                        R h_eff = h;
                        if (sub == 0)
                            detect_events<R>(wev, s_ev, d.n_events, d.n_tevents, (int)p, h, hi, lo, traj, ns, P.ev, h_eff,
                                             term_ev, nt_fired);
                        h_eff = __shfl_sync(gmask, h_eff, 0, G);
                        term_ev = __shfl_sync(gmask, term_ev, 0, G);
                        nt_fired = __shfl_sync(gmask, nt_fired, 0, G);
                        if (term_ev >= 0) {
                            h = h_eff;
                            hn = h_eff;
                            so = -(long long)term_ev - 1;
                        }
                    }
                    // (cd_live: a cooldown may be running - after the fetch, and from a terminal event on)
                    if (term_ev >= 0) cd_live = true;
                    if (stepping && sub == 0 && d.n_tevents && cd_live) cd_live = advance_cooldowns<R>(traj, d.n_tevents, h, P.ev);
                    __syncwarp();
                }
            }

            // ---- continuous-output record on the register-resident kernels: the WARP copies the step
            // records of its trajectories one after the other, 32 consecutive elements per store (full
            // 256-byte segments; the per-group loop below writes 8 G bytes per trajectory and store)
            if constexpr (FX && NB < 0 && sizeof(R) == 8) {
                if (P.rec.on && P.rec.sk != 1u) {
                    // order-major records: orders 1..p of the trajectory's column ARE the record - one bulk copy
                    // shared -> global (TMA) issued by lane 0; order 0 (which the state update below rewrites) goes
                    // by ordinary stores.  The copy is waited for at the top of the next step, before the jets
                    // overwrite the column.
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (stepping) {
                        R *dst = P.rec.chunk(rec_tail) + 2u + (cc % HY_REC_CH) * P.rec.rec_len;
                        for (uint32_t i = sub; i < n; i += G) dst[i] = w[i];
                        if (sub == 0) {
                            const uint32_t sa = (uint32_t)__cvta_generic_to_shared(w + n);
                            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + n), "r"(sa),
                                         "r"((uint32_t)(n * p * 8u))
                                         : "memory");
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                    }
                }
            }
            if constexpr (FX && NB != 0) {
                if (P.rec.on && !(NB < 0 && sizeof(R) == 8 && P.rec.sk != 1u)) {
                    R *my_dst = nullptr;
                    if (stepping) my_dst = P.rec.chunk(rec_tail) + 2u + (cc % HY_REC_CH) * P.rec.rec_len;
                    const uint32_t nP = n * P1;
                    // (P.rec_off: the [variable][order] -> column offset table; records of up to 128 elements -
                    //  the CR3BP's 126 - keep this lane's four offsets in registers: per trajectory four
                    //  independent load / store pairs instead of a chain of dependent table look-ups)
                    const bool small_rec = nP <= 128u;
                    uint32_t ro[4];
#pragma unroll
                    for (int m_ = 0; m_ < 4; ++m_)
                        ro[m_] = (small_rec && lane + 32u * m_ < nP) ? __ldg(&P.rec_off[lane + 32u * m_]) : 0u;
#pragma unroll 1
                    for (uint32_t g = 0; g < 32u / G; ++g) {
                        R *dst = reinterpret_cast<R *>(__shfl_sync(TM_FULL, (unsigned long long)my_dst, (int)(g * G)));
                        if (!dst) continue; // (warp-uniform)
                        const R *wg = w + ((int)g - (int)(lane / G)) * (int)RS; // column of group g of this warp
                        if (small_rec) {
                            R v_[4];
#pragma unroll
                            for (int m_ = 0; m_ < 4; ++m_) v_[m_] = wg[ro[m_]];
#pragma unroll
                            for (int m_ = 0; m_ < 4; ++m_)
                                if (lane + 32u * m_ < nP) dst[lane + 32u * m_] = v_[m_];
                        } else {
                            for (uint32_t e = lane; e < nP; e += 32u) dst[e] = wg[__ldg(&P.rec_off[e])];
                        }
                    }
                    __syncwarp();
                }
            }
            // ---- optional tc write, then the state update ----
            if (stepping && (P.write_tc || (FX && P.rec.on) || P.mode == MODE_GRID)) {
                // tc: only the coefficients of the lane's LAST step of this launch can be observed (every step
                // overwrites them; callbacks that read them run between launches) - time limit reached, terminal
                // event, step / launch budget, pause for a callback, or a single step.  (A step that ends in a
                // non-finite state is found out after the update: its coefficients are not written.)
                const bool last_step =
                    P.mode == MODE_STEP || term_ev >= 0 || (so == HY_OUTCOME_TIME_LIMIT && h == rem) ||
                    (P.max_steps && ns + 1 >= P.max_steps) ||
                    ((FX && P.launch_steps) && (uint32_t)(ns + 1) - s_ns0[slot] >= (uint32_t)P.launch_steps) ||
                    (FX && P.pause_on_nt && nt_fired);
                if (P.write_tc && P.tc && last_step) {
                    // (variable-major loops: no division per element)
                    for (uint32_t v_ = 0; v_ < n; ++v_)
                        for (uint32_t k_ = sub; k_ < P1; k_ += G) P.tc[(size_t)(v_ * P1 + k_) * P.B + traj] = XJ(v_, k_);
                    if (G > 1) __syncwarp(gmask);
                }
                if ((FX && P.rec.on) && NB == 0) {
                    // step record: [n][p+1] coefficients (the end time follows after the time update)
                    R *dstc = P.rec.chunk(rec_tail) + 2u + (cc % HY_REC_CH) * P.rec.rec_len;
                    for (uint32_t v_ = 0; v_ < n; ++v_) {
                        R *dv = dstc + v_ * P1;
                        for (uint32_t k_ = sub; k_ < P1; k_ += G) dv[k_] = XJ(v_, k_);
                    }
                }
                if (P.mode == MODE_GRID) {
                    // Dense output at every grid point inside this step.
                    while (gi < P.grid_k) {
                        const R g = P.grid[(size_t)gi * P.B + traj];
                        const R tau = (g - hi) - lo; // time since the start of the step
                        if (r_abs(tau) > r_abs(h)) break;
                        for (uint32_t i = sub; i < n; i += G) {
                            R acc = XJ(i, p);
                            for (uint32_t k = p; k-- > 0;) acc = r_fma(acc, tau, XJ(i, k));
                            P.gout[((size_t)gi * n + i) * P.B + traj] = acc;
                        }
                        ++gi;
                    }
                }
                if (G > 1) __syncwarp(gmask);
            }
            if constexpr (NB != 0) __syncwarp(); // reconverge the trajectories of the warp
            bool finite = true;
            if (NB > 0 && !P.high_accuracy) {
                // Horner on the lane's 3-vector: three independent chains, 128-bit + 64-bit loads
                if (vown) {
                    R *x = w + voff;
                    Vec3<R> acc, t;
                    lds3(x + p * NBR_JS, acc);
                    if (!WGX && p == (uint32_t)PM) {
                        // all loads first (independent), then the three FMA chains
                        Vec3<R> c[PM];
#pragma unroll
                        for (int k = 0; k < PM; ++k) lds3(x + k * NBR_JS, c[k]);
#pragma unroll
                        for (int k = PM - 1; k >= 0; --k) {
                            acc.x = r_fma(acc.x, hn, c[k].x);
                            acc.y = r_fma(acc.y, hn, c[k].y);
                            acc.z = r_fma(acc.z, hn, c[k].z);
                        }
                    } else {
                        for (uint32_t k = p; k-- > 0;) {
                            lds3(x + k * NBR_JS, t);
                            acc.x = r_fma(acc.x, hn, t.x);
                            acc.y = r_fma(acc.y, hn, t.y);
                            acc.z = r_fma(acc.z, hn, t.z);
                        }
                    }
                    finite = r_abs(acc.x) < r_inf<R>() && r_abs(acc.y) < r_inf<R>() && r_abs(acc.z) < r_inf<R>();
                    if (stepping) sts3(x, acc.x, acc.y, acc.z);
                }
            } else if (stepping) {
                for (uint32_t i = sub; i < n; i += G) {
                    R acc;
                    const int sp = s_ssp[i];
                    if (sp >= 0) {
                        // spilled jet: stream it back from the global scratch (independent loads)
                        const R *x = gj + (uint32_t)sp * P1;
                        if (!P.high_accuracy) {
                            acc = __ldcg(&x[p]);
                            for (uint32_t k = p; k-- > 0;) acc = r_fma(acc, h, __ldcg(&x[k]));
                        } else {
                            R sum = __ldcg(&x[0]), comp = 0, hk = h;
                            for (uint32_t k = 1; k <= p; ++k) {
                                const R term = __ldcg(&x[k]) * hk;
                                const R y = ef_sub(term, comp);
                                const R tt = ef_add(sum, y);
                                comp = ef_sub(ef_sub(tt, sum), y);
                                sum = tt;
                                hk = hk * h;
                            }
                            acc = sum;
                        }
                        gj[(uint32_t)sp * P1] = acc;
                    } else {
                        const R *x = &w[s_srow[i]];
                        if (!P.high_accuracy) {
                            acc = x[p * XS];
                            for (uint32_t k = p; k-- > 0;) acc = r_fma(acc, h, x[k * XS]);
                        } else {
                            R sum = x[0], comp = 0, hk = h;
                            for (uint32_t k = 1; k <= p; ++k) {
                                const R term = x[k * XS] * hk;
                                const R y = ef_sub(term, comp);
                                const R tt = ef_add(sum, y);
                                comp = ef_sub(ef_sub(tt, sum), y);
                                sum = tt;
                                hk = hk * h;
                            }
                            acc = sum;
                        }
                    }
                    finite = finite && (r_abs(acc) < r_inf<R>());
                    w[s_srow[i]] = acc;
                }
            }
            if constexpr (NB > 0) {
                // one full-mask ballot, each half-warp looks at its own 16 bits
                const unsigned bad = __ballot_sync(TM_FULL, !finite);
                finite = G == 32 ? bad == 0u : ((bad >> (lane & 16u)) & 0xffffu) == 0u;
            } else if constexpr (NB < 0) {
                __syncwarp();
                const unsigned bad = __ballot_sync(TM_FULL, !finite);
                finite = ((bad >> (lane & ~1u)) & 3u) == 0u;
            } else {
                if (G > 1) finite = !__any_sync(gmask, !finite);
            }
#if defined(HY_WGX_PROF) && defined(HY_TAIL_PROF)
            prof_rel += clock64() - prof_c;
#endif
            // ---- angle reduction (the reference's callback.angle_reducer, expose_callbacks.cpp:67-72,
            // as a device-side post-step op): x <- x - 2 pi floor(x / (2 pi)) for the selected variables.
            // Runs in the propagate modes only (step callbacks are not part of step()).
            if (FX && P.n_red && P.mode != MODE_STEP) {
                if constexpr (NB != 0) {
                    __syncwarp();
                } else {
                    if (G > 1) __syncwarp(gmask);
                }
                if (stepping) {
                    const R two_pi = (R)6.283185307179586476925286766559;
                    for (uint32_t q = sub; q < P.n_red; q += G) {
                        const uint32_t i = P.red_idx[q];
                        const R x = w[s_srow[i]];
                        const R y = x - two_pi * floor(x / two_pi);
                        w[s_srow[i]] = y;
                        if (s_ssp[i] >= 0) gj[(uint32_t)s_ssp[i] * P1] = y;
                    }
                }
                if constexpr (NB != 0) {
                    __syncwarp();
                } else {
                    if (G > 1) __syncwarp(gmask);
                }
            }
            if (stepping) {
                time_add(hi, lo, h);
                ++ns;
                if (!finite) so = HY_OUTCOME_ERR_NF_STATE;
                if ((FX && P.rec.on)) {
                    if (sub == 0) {
                        const bool fin_ = (P.mode != MODE_STEP) && so == HY_OUTCOME_TIME_LIMIT && h == rem;
                        R *dstt = P.rec.chunk(rec_tail) + 2u + (cc % HY_REC_CH) * P.rec.rec_len + n * P1;
                        dstt[0] = fin_ ? tf_hi : hi;
                        dstt[1] = fin_ ? tf_lo : lo;
                    }
                    ++cc;
                }

                // ---- does the trajectory end here? ----
                if (P.mode == MODE_STEP || so == HY_OUTCOME_ERR_NF_STATE || term_ev >= 0) {
                    // single step / non-finite state / terminal event (the host may resume the lane)
                    oc = so;
                    fin = true;
                } else {
                    if (so == HY_OUTCOME_SUCCESS) {
                        const R ah = r_abs(h);
                        if (ah < mn) mn = ah;
                        if (ah > mx) mx = ah;
                    }
                    if (so == HY_OUTCOME_TIME_LIMIT && h == rem) {
                        hi = tf_hi;
                        lo = tf_lo;
                        oc = HY_OUTCOME_TIME_LIMIT;
                        fin = true;
                    } else if (P.max_steps && ns >= P.max_steps) {
                        oc = HY_OUTCOME_STEP_LIMIT;
                        fin = true;
                    } else if (((FX && P.launch_steps) && (uint32_t)ns - s_ns0[slot] >= (uint32_t)P.launch_steps) ||
                               (FX && P.pause_on_nt && nt_fired)) {
                        // the host wants the lane back (step callback / non-terminal event callback)
                        oc = HY_OUTCOME_PAUSED;
                        fin = true;
                    }
                }
            }
            if constexpr (NB != 0) {
                __syncwarp();
            } else {
                if (G > 1) __syncwarp(gmask);
            }
        }

        // ---- retire the trajectory ----
        if (have && fin) {
            for (uint32_t i = sub; i < n; i += G) P.state[(size_t)i * P.B + traj] = w[s_srow[i]];
            if (sub == 0) {
                P.t_hi[traj] = hi;
                P.t_lo[traj] = lo;
                P.last_h[traj] = h;
                P.outcome[traj] = oc;
                P.min_h[traj] = mn;
                P.max_h[traj] = mx;
                P.n_steps[traj] = ns;
                if (FX && P.gidx) P.gidx[traj] = gi;
                if ((FX && P.rec.on)) P.rec.count[traj] = cc;
            }
            have = false;
            if (G > 1) __syncwarp(gmask);
        }
    }
    // (bulk copies of continuous-output records still in flight: complete before the thread exits)
    if constexpr (FX && NB < 0 && sizeof(R) == 8) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#ifdef HY_WGX_PROF
    if constexpr (NB > 0) {
        if (blockIdx.x == 0 && lane == 0 && prof_n)
            printf("warp %u: iters %lld  per iter: wait %lld jets %lld release %lld tail %lld\n", threadIdx.x >> 5, prof_n,
                   prof_wait / prof_n, prof_jets / prof_n, prof_rel / prof_n,
                   (clock64() - prof_t0 - prof_wait - prof_jets - prof_rel) / prof_n);
    }
#endif
}

// ---- dense output of the last step (reference update_d_output,
// expose_batch_integrators.cpp:519-541) ----
// tc is [n][p+1][B]; one thread per lane, coalesced over lanes.
template <typename R>
__global__ void dense_eval_kernel(const R *__restrict__ tc, const R *__restrict__ t_hi, const R *__restrict__ t_lo,
                                  const R *__restrict__ last_h, const R *__restrict__ t, int rel_time,
                                  R *__restrict__ out, uint32_t n, uint32_t p, uint32_t B)
{
    const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= B) return;
    // tau = time since the START of the last step
    const R tau = rel_time ? last_h[l] + t[l] : ((t[l] - t_hi[l]) - t_lo[l]) + last_h[l];
    const uint32_t P1 = p + 1;
    for (uint32_t i = 0; i < n; ++i) {
        const R *x = tc + (size_t)i * P1 * B + l;
        R acc = x[(size_t)p * B];
        for (uint32_t k = p; k-- > 0;) acc = r_fma(acc, tau, x[(size_t)k * B]);
        out[(size_t)i * B + l] = acc;
    }
}

// ---- absolute final times of propagate_until / propagate_for (double-length), one thread per lane ----
template <typename R>
__global__ void prep_tf_kernel(const R *__restrict__ t, int is_delta, const R *__restrict__ t_hi,
                               const R *__restrict__ t_lo, R *__restrict__ tf_hi, R *__restrict__ tf_lo, uint32_t B)
{
    const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= B) return;
    R hi = t[l], lo = 0;
    if (is_delta) {
        hi = t_hi[l];
        lo = t_lo[l];
        time_add(hi, lo, t[l]);
    }
    tf_hi[l] = hi;
    tf_lo[l] = lo;
}

// ---- continuous output (reference continuous_output_batch, taylor_expose_c_output.cpp:260-526) ----
// The recorder leaves per-lane linked lists of chunks; rec_index_kernel flattens them into a
// directory (dir[dir_off[l] + j] = id of the lane's j-th chunk) so that a step is found in O(1).
template <typename R>
__global__ void rec_index_kernel(const RecDev<R> rec, uint32_t *__restrict__ dir_off, uint32_t *__restrict__ dir,
                                 unsigned int *__restrict__ dir_next, uint32_t B)
{
    const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= B) return;
    const uint32_t nch = (rec.count[l] + HY_REC_CH - 1u) / HY_REC_CH;
    const uint32_t off = nch ? atomicAdd(dir_next, nch) : 0u;
    dir_off[l] = off;
    uint32_t cid = rec.head[l];
    for (uint32_t j = 0; j < nch; ++j) {
        dir[off + j] = cid;
        cid = *reinterpret_cast<const uint32_t *>(rec.chunk(cid));
    }
}

template <typename R>
__device__ __forceinline__ const R *rec_step(const RecDev<R> &rec, const uint32_t *dir, uint32_t off, uint32_t s)
{
    return rec.chunk(dir[off + s / HY_REC_CH]) + 2u + (s % HY_REC_CH) * rec.rec_len;
}

// Evaluation: per-lane bisection over the recorded end times + Horner.
// t [K][B], out [K][n][B]; a lane without a record evaluates to NaN.
template <typename R>
__global__ void cout_eval_kernel(const RecDev<R> rec, const uint32_t *__restrict__ dir_off,
                                 const uint32_t *__restrict__ dir, const R *__restrict__ t, R *__restrict__ out,
                                 uint32_t n, uint32_t p, uint32_t B, uint32_t K)
{
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)K * B) return;
    const uint32_t l = (uint32_t)(gid % B), q = (uint32_t)(gid / B);
    const uint32_t S = rec.count[l], P1 = p + 1, nP = n * P1;
    const R tq = t[(size_t)q * B + l];
    if (S == 0) {
        for (uint32_t i = 0; i < n; ++i) out[((size_t)q * n + i) * B + l] = tq - tq + (R)NAN;
        return;
    }
    const uint32_t off = dir_off[l];
    const bool fwd = rec_step(rec, dir, off, S - 1)[nP] >= rec.t0_hi[l];
    // first step s whose end time is beyond tq (clamped to the recorded range)
    uint32_t lo = 0, hi = S - 1;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        const R te = rec_step(rec, dir, off, mid)[nP];
        const bool before = fwd ? (tq < te) : (tq > te);
        if (before)
            hi = mid;
        else
            lo = mid + 1;
    }
    const uint32_t sidx = lo;
    R t0h = rec.t0_hi[l], t0l = rec.t0_lo[l];
    if (sidx > 0) {
        const R *pr = rec_step(rec, dir, off, sidx - 1);
        t0h = pr[nP];
        t0l = pr[nP + 1];
    }
    const R tau = (tq - t0h) - t0l;
    const R *base = rec_step(rec, dir, off, sidx);
    const uint32_t si = rec.si, sk = rec.sk;
    if (sk != 1u && n <= 8u) {
        // order-major record: all variables advance together, every order one contiguous read
        R acc[8];
#pragma unroll
        for (uint32_t i = 0; i < 8u; ++i) acc[i] = i < n ? base[p * sk + i] : (R)0;
        for (uint32_t k = p; k-- > 0;) {
            const R *x = base + k * sk;
#pragma unroll
            for (uint32_t i = 0; i < 8u; ++i)
                if (i < n) acc[i] = r_fma(acc[i], tau, x[i]);
        }
#pragma unroll
        for (uint32_t i = 0; i < 8u; ++i)
            if (i < n) out[((size_t)q * n + i) * B + l] = acc[i];
        return;
    }
    for (uint32_t i = 0; i < n; ++i) {
        const R *x = base + i * si;
        R acc = x[p * sk];
        for (uint32_t k = p; k-- > 0;) acc = r_fma(acc, tau, x[k * sk]);
        out[((size_t)q * n + i) * B + l] = acc;
    }
}

// Dense copies in the reference's layouts: tcs [S][n][p+1][B] and times (hi, lo) [S+1][B], NaN past a
// lane's own count (taylor_expose_c_output.cpp:449-451).  One thread per (step, lane).
template <typename R>
__global__ void rec_gather_kernel(const RecDev<R> rec, const uint32_t *__restrict__ dir_off,
                                  const uint32_t *__restrict__ dir, R *__restrict__ tcs, R *__restrict__ thi,
                                  R *__restrict__ tlo, uint32_t n, uint32_t p, uint32_t B, uint32_t S)
{
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)(S + 1) * B) return;
    const uint32_t l = (uint32_t)(gid % B), s = (uint32_t)(gid / B);
    const uint32_t cnt = rec.count[l], nP = n * (p + 1);
    const R nan = (R)NAN;
    if (s == 0) {
        thi[l] = rec.t0_hi[l];
        tlo[l] = rec.t0_lo[l];
    }
    if (s < S) {
        const R *src = s < cnt ? rec_step(rec, dir, dir_off[l], s) : nullptr;
        if (tcs)
            for (uint32_t v_ = 0, i = 0; v_ < n; ++v_)
                for (uint32_t k_ = 0; k_ <= p; ++k_, ++i)
                    tcs[((size_t)s * nP + i) * B + l] = src ? src[v_ * rec.si + k_ * rec.sk] : nan;
        thi[(size_t)(s + 1) * B + l] = src ? src[nP] : nan;
        tlo[(size_t)(s + 1) * B + l] = src ? src[nP + 1] : nan;
    }
}

// ---- FMA peak microbenchmark (compute roof) ----
template <typename R> __global__ void fma_peak_kernel(R *out, int iters)
{
    R a0 = threadIdx.x * (R)1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
      a7 = a0 + 7;
    const R b = (R)1.0000001, c = (R)1e-7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            a0 = r_fma(a0, b, c);
            a1 = r_fma(a1, b, c);
            a2 = r_fma(a2, b, c);
            a3 = r_fma(a3, b, c);
            a4 = r_fma(a4, b, c);
            a5 = r_fma(a5, b, c);
            a6 = r_fma(a6, b, c);
            a7 = r_fma(a7, b, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

} // namespace hy
