/*
 * hy_cuda.h - C ABI of libhy_cuda, the B200 batch Taylor integrator.
 *
 * This is the drop-in boundary for ONE hot path of bluescarni/heyoka.py: the
 * `taylor_adaptive_batch` step / propagate loop and the ensemble driver.  Each
 * entry point below names the reference interface (file:line under
 * /root/reference) that a binding layer would route to it.  No C++ or torch
 * types cross this boundary: plain pointers, sizes and POD structs only.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on failure; the message is
 *    available from hy_last_error() (thread-local).
 *  - `fp_bits` selects float (32) or double (64); every `void*` array argument
 *    below holds elements of that type unless stated otherwise.
 *  - host arrays are row-major with the batch (lane) index fastest, exactly like
 *    the reference's `state[n, B]`, `pars[m, B]`, `time[B]`
 *    (expose_batch_integrators.cpp:115-157).
 *  - a hy_ctx is single-threaded; distinct contexts may be driven from distinct
 *    host threads (reference: one integrator object is not thread-safe,
 *    _ensemble_impl.py:52).
 */
#ifndef HY_CUDA_H
#define HY_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------
 * The opcode tape: the host-lowered Taylor decomposition (u-variables plus
 * elementary operations) of an ODE system.  It replaces the LLVM IR the
 * reference JIT-compiles in the taylor_adaptive_batch constructor
 * (expose_batch_integrators.cpp:166-208).
 *
 * Every u-variable owns rows of a per-trajectory workspace.  A "jet" u-variable
 * owns order+1 consecutive rows (row base+k = normalised derivative of order
 * k); a "cur" u-variable owns one row holding only the order being computed.
 * State variable i is always the jet at rows [i*(order+1), (i+1)*(order+1)).
 * A row reference is `base | HY_REF_JET` for jets, `base` for cur rows.
 * ------------------------------------------------------------------------ */
#define HY_REF_JET 0x80000000u
#define HY_REF_ONE 0x7fffffffu /* pseudo operand: the constant jet [1,0,0,...] */

enum hy_opcode {
    HY_OP_LINCOMB = 0, /* dst[k] = sum_i coef_i * src_i[k]           (terms)   */
    HY_OP_MUL = 1,     /* dst[k] = sum_j a[j] b[k-j]                           */
    HY_OP_SQUARE = 2,  /* dst[k] = sum_j a[j] a[k-j]   (symmetric evaluation)  */
    HY_OP_DIV = 3,     /* dst[k] = (a[k] - sum_{j>=1} b[j] dst[k-j]) / b[0]    */
    HY_OP_POW = 4,     /* dst = a^imm                                          */
    HY_OP_SQRT = 5,    /* dst = a^(1/2), order 0 by sqrt()                     */
    HY_OP_EXP = 6,     /* dst = exp(a)                                         */
    HY_OP_LOG = 7,     /* dst = log(a)                                         */
    HY_OP_SINCOS = 8,  /* dst = sin(a), dst2 = cos(a) (always as a pair)       */
    HY_OP_TIME = 9,    /* dst = [t, 1, 0, ...]                                 */
    HY_OP_SVD = 10,    /* state jet: dst[k+1] = a[k] / (k+1)                   */
    HY_OP_SUMSQ = 11,  /* dst[k] = sum_i sum_j a_i[j] a_i[k-j]       (terms)   */
    HY_OP_MULSH = 12,  /* dst_i[k] = sum_j a_i[j] b[k-j], i<n (terms=a_i,dst_i)*/
    HY_OP_ADDSUB = 13, /* dst[k] = (+-)a[k] (+-)b[k]   (signs: HY_OPF_NEGA/NEGB)  */
    HY_OP_INTG = 14,   /* dst = F(a) for a function with dF/da = b (both jets):
                          dst[0] = F(a[0]), dst[k] = (1/k) sum_{j=1..k} j a[j] b[k-j];
                          imm selects F: 0 asin, 1 acos, 2 atan, 3 erf - b is built
                          by earlier ops: +-(1-a^2)^(-1/2), 1/(1+a^2), 2/sqrt(pi) exp(-a^2) */
    HY_OP_COUNT
};

/* flags */
#define HY_OPF_EVENT 0x1u /* needed at order p for the event polynomials */
#define HY_OPF_NEGA 0x2u  /* ADDSUB: first operand enters with a minus sign  */
#define HY_OPF_NEGB 0x4u  /* ADDSUB: second operand enters with a minus sign */
#define HY_OPF_SVD 0x8u   /* LINCOMB/ADDSUB fused with the state recurrence:
                             dst is a state jet and the op stores
                             dst[k+1] = value[k] / (k+1) instead of dst[k]   */

typedef struct hy_op {
    uint16_t opcode;
    uint16_t flags;
    uint32_t dst;  /* row reference of the (first) output                     */
    uint32_t dst2; /* second output (SINCOS) or unused                        */
    uint32_t a;    /* first operand row reference                             */
    uint32_t b;    /* second operand row reference, or first term index       */
    uint32_t n;    /* number of terms (LINCOMB / SUMSQ / MULSH)               */
    double imm;    /* POW exponent; INTG: code of the function                */
} hy_op; /* 32 bytes */

typedef struct hy_term {
    uint32_t src; /* row reference (or HY_REF_ONE)                            */
    int32_t par;  /* >= 0: multiply coef by pars[par]; -1: none               */
    double coef;  /* numeric coefficient (MULSH: unused)                      */
    uint32_t dst; /* MULSH: output row reference of this term; else unused    */
    uint32_t pad;
} hy_term; /* 24 bytes */

typedef struct hy_dims {
    uint32_t n_state;   /* number of state variables n                        */
    uint32_t n_par;     /* number of runtime parameters m                     */
    uint32_t order;     /* Taylor order p                                     */
    uint32_t n_rows;    /* workspace rows per trajectory                      */
    uint32_t n_ops;     /* length of the op array                             */
    uint32_t n_terms;   /* length of the term array                           */
    uint32_t n_levels;  /* number of dependency levels                        */
    uint32_t n_events;  /* number of event functions (terminal first)         */
    uint32_t n_tevents; /* how many of those are terminal                     */
} hy_dims;

/* Outcome codes: values of the reference's `taylor_outcome` enum
 * (core.cpp:324-336).  Non-negative / small negative values are terminal-event
 * indices: idx (continuing) and -idx-1 (stopping). */
#define HY_OUTCOME_SUCCESS (-4294967297LL)
#define HY_OUTCOME_STEP_LIMIT (-4294967298LL)
#define HY_OUTCOME_TIME_LIMIT (-4294967299LL)
#define HY_OUTCOME_ERR_NF_STATE (-4294967300LL)
#define HY_OUTCOME_CB_STOP (-4294967301LL)
/* Not a reference value: returned by hy_propagate_ex for a lane that stopped inside a launch
 * and can be resumed by the next one (per-launch step budget reached, or a non-terminal event
 * fired and the caller asked to get the lane back to run its callback). */
#define HY_OUTCOME_PAUSED (-4294967400LL)

typedef struct hy_ctx hy_ctx;
typedef struct hy_cout hy_cout; /* a recorded continuous output (device memory) */

/* One record of the device event log (hy_events_drain). */
typedef struct hy_event_rec {
    uint32_t lane;   /* batch index                                           */
    uint32_t ev_idx; /* event index (terminal events first)                   */
    int32_t d_sgn;   /* sign of the event function's time derivative at root  */
    uint32_t step;   /* lane-local step counter at which it fired             */
    double t;        /* absolute trigger time                                 */
} hy_event_rec;

const char *hy_last_error(void);
int hy_device_count(int *count);

/* Replaces the reference constructor (expose_batch_integrators.cpp:91-212):
 * uploads the tape, allocates device state for `batch` lanes on `device`.
 * `level_start` has n_levels+1 entries delimiting the ops of each dependency
 * level; `ev_ref` has n_events jet references; `ev_dir` their directions
 * (-1, 0 = any, +1); `ev_cooldown` the terminal events' cooldowns (<0: auto).
 * `tol` fixes the step-size safety factor; the order is dims->order. */
int hy_create(hy_ctx **out, int device, int fp_bits, const hy_dims *dims, const hy_op *ops,
              const hy_term *terms, const uint32_t *level_start, const uint32_t *ev_ref,
              const int32_t *ev_dir, const double *ev_cooldown, double tol, int high_accuracy,
              uint32_t batch);
int hy_destroy(hy_ctx *ctx);

/* hy_create with the two extra tapes that let EVENT-carrying systems run on the
 * register-resident kernels (those keep the jets of the ODE's sub-expressions in registers, so
 * event functions cannot share u-variables with the ODE):
 *  - `ode`: the tape of the ODE alone (no event functions; dims->n_events = 0) - what the
 *    N-body / CR3BP matchers look at;
 *  - `evt`: the event functions alone, as functions of the state jets, one event after the
 *    other without shared sub-expressions (ops [op_start[e], op_start[e+1]) belong to event e).
 *    Row references: state variable i = (i * (order + 1)) | HY_REF_JET as in the main tape,
 *    everything else is numbered from n_state * (order + 1) ("event workspace", n_rows rows).
 * `full` is the combined tape hy_create takes (used by the tape interpreter whenever the ODE
 * tape is not matched).  `ode` and `evt` may be NULL (then this is hy_create).
 * Reference: events are part of the integrator's construction,
 * expose_batch_integrators.cpp:166-208 (t_events / nt_events keyword arguments). */
typedef struct hy_tape {
    const hy_dims *dims;
    const hy_op *ops;
    const hy_term *terms;
    const uint32_t *level_start;
    const uint32_t *ev_ref;
} hy_tape;
typedef struct hy_event_tape {
    uint32_t n_ops, n_terms, n_rows, n_events;
    const hy_op *ops;
    const hy_term *terms;
    const uint32_t *ev_ref;   /* [n_events] jet of every event function                 */
    const uint32_t *op_start; /* [n_events + 1]                                         */
} hy_event_tape;
int hy_create2(hy_ctx **out, int device, int fp_bits, const hy_tape *full, const hy_tape *ode,
               const hy_event_tape *evt, const int32_t *ev_dir, const double *ev_cooldown, double tol,
               int high_accuracy, uint32_t batch);

/* Deep copy of a context onto `device` (< 0: the source's device).  Replaces what
 * copy.deepcopy(ta) does once per ensemble iteration in the reference
 * (_ensemble_impl.py:47; copy_wrapper/deepcopy_wrapper, common_utils.hpp): the scheduled
 * program is reused - no tape matching, no scheduling - and the lane data (state, pars,
 * time, last_h, results, tc, cooldowns) is copied device to device (peer copy across GPUs). */
int hy_clone(const hy_ctx *src, hy_ctx **out, int device);
int hy_get_device(hy_ctx *ctx, int *device);
/* Debug counters of the event path on the register-resident kernels (enabled by
 * HY_CUDA_EVENT_STATS=1 when the context is created): steps taken, and steps in which the
 * interval enclosure over the step could not exclude an event (full evaluation + root finder). */
int hy_get_event_stats(hy_ctx *ctx, uint64_t *steps, uint64_t *full_evals);
/* Wait for everything queued on the context's stream. */
int hy_sync(hy_ctx *ctx);

/* Page-locked host buffers for the numpy mirrors of state/pars/time (the
 * reference's views alias integrator memory, expose_batch_integrators.cpp:
 * 394-518; here they alias pinned memory the DMA engines can reach). */
int hy_host_alloc(void **ptr, size_t bytes);
int hy_host_free(void *ptr);

/* Use an externally owned CUDA stream (cudaStream_t) for all work of ctx. */
int hy_set_stream(hy_ctx *ctx, void *cuda_stream);

/* Host <-> device state transfer.  Replaces the writable numpy views of the
 * reference (expose_batch_integrators.cpp:394-472): the host mirrors are
 * pushed before and pulled after each step/propagate call.  NULL = skip. */
int hy_upload(hy_ctx *ctx, const void *state, const void *pars, const void *t_hi,
              const void *t_lo);
int hy_download(hy_ctx *ctx, void *state, void *t_hi, void *t_lo, void *last_h);

/* Device-side Taylor coefficients [n, order+1, B] / last step sizes [B] of a copied or
 * unpickled integrator (the reference's copies carry tc and last_h:
 * expose_batch_integrators.cpp:665-669, pickle_wrappers.hpp:35-73). */
int hy_set_tc(hy_ctx *ctx, const void *tc);
int hy_set_last_h(hy_ctx *ctx, const void *last_h);

/* Device-resident variants (pointers are device addresses of the same
 * layouts); used when the ensemble already lives in HBM. */
int hy_upload_dev(hy_ctx *ctx, const void *d_state, const void *d_pars, const void *d_t_hi,
                  const void *d_t_lo);
int hy_state_dev(hy_ctx *ctx, void **d_state, void **d_t_hi, void **d_t_lo);

/* One adaptive step for every lane: reference `step(write_tc)` /
 * `step(max_delta_t, write_tc)` / `step_backward`
 * (expose_batch_integrators.cpp:233-241).  `max_delta_t` is [B] or NULL
 * (= +inf, or -inf when backward != 0).  Outputs (host, [B]): outcome, h. */
int hy_step(hy_ctx *ctx, const void *max_delta_t, int backward, int write_tc, int64_t *outcome,
            void *h);

/* propagate_for / propagate_until (expose_batch_integrators.cpp:243-314).
 * `t` is [B]: the final times (is_delta = 0) or the time intervals
 * (is_delta = 1).  `max_delta_t` is [B] or NULL.  max_steps = 0: unlimited.
 * Outputs (host, [B], any may be NULL): outcome, min_h, max_h, n_steps -
 * the reference's `propagate_res` tuples (:393). */
int hy_propagate(hy_ctx *ctx, const void *t, int is_delta, uint64_t max_steps,
                 const void *max_delta_t, int write_tc, int c_output, int64_t *outcome,
                 void *min_h, void *max_h, uint64_t *n_steps);

/* The general form behind hy_propagate / hy_propagate_grid.  It adds what the front end
 * needs to run Python between launches without re-integrating or re-uploading anything
 * (step callbacks: step_cb_utils.cpp:70-98; event callbacks: taylor_expose_events.cpp:109-138):
 *  - `active` [B] bytes (NULL: all): only these lanes take part; the others keep their state,
 *    time and results;
 *  - `resume`: the lanes continue the previous call (final times, step counters, min/max h,
 *    grid position are kept on the device; `t` / `grid` are not read again);
 *  - `launch_steps` > 0: a lane gives control back after that many steps, `pause_on_nt`: after
 *    a step that logged a non-terminal event; its outcome is then HY_OUTCOME_PAUSED;
 *  - `c_output`: 0 off, 1 record into a fresh continuous output, 2 append to the current one;
 *  - `grid` != NULL selects propagate_grid (t/is_delta ignored). */
typedef struct hy_prop_args {
    const void *t;           /* [B] final times or time intervals                       */
    int is_delta;
    uint64_t max_steps;      /* per lane over the whole (resumed) call, 0 = unlimited    */
    const void *max_delta_t; /* [B] or NULL                                             */
    int write_tc;
    int c_output;
    const uint8_t *active;
    int resume;
    uint64_t launch_steps;
    int pause_on_nt;
    const void *grid;        /* [grid_k, B] or NULL                                     */
    size_t grid_k;
    void *grid_out;          /* [grid_k, n, B]                                          */
} hy_prop_args;
int hy_propagate_ex(hy_ctx *ctx, const hy_prop_args *args, int64_t *outcome, void *min_h, void *max_h,
                    uint64_t *n_steps);

/* The reference's callback.angle_reducer (expose_callbacks.cpp:67-72), a C++ builtin there,
 * as a device-side post-step op here: the listed state variables are reduced to [0, 2 pi)
 * after every step of the propagate_* calls.  n = 0 switches it off. */
int hy_set_angle_reducer(hy_ctx *ctx, const uint32_t *state_idx, uint32_t n);

/* propagate_grid (expose_batch_integrators.cpp:315-392): `grid` is [k, B]
 * host; `out` is [k, n, B] host, NaN-filled past an early exit. */
int hy_propagate_grid(hy_ctx *ctx, const void *grid, size_t k, uint64_t max_steps,
                      const void *max_delta_t, void *out, int64_t *outcome, void *min_h,
                      void *max_h, uint64_t *n_steps);

/* Timing of the last hy_step/hy_propagate* call, measured with CUDA events
 * on the context's stream: milliseconds spent in the propagate kernel(s)
 * and the number of kernel launches. */
int hy_last_timing(hy_ctx *ctx, double *kernel_ms, uint64_t *launches);

/* Taylor coefficients of the last step, [n, order+1, B]
 * (reference `tc`, expose_batch_integrators.cpp:473-489). */
int hy_get_tc(hy_ctx *ctx, void *tc);

/* Dense output of the last step (reference `update_d_output`,
 * expose_batch_integrators.cpp:519-541).  `t` is [B]; out is [n, B]. */
int hy_dense_eval(hy_ctx *ctx, const void *t, int rel_time, void *out);

/* Continuous output recorded by hy_propagate(c_output=1)
 * (reference continuous_output_batch, taylor_expose_c_output.cpp:260-526).  The propagate
 * kernel records in ONE pass: every lane appends its steps (Taylor coefficients + end time)
 * to a per-lane list of fixed-size chunks taken from a device pool (ragged storage: a lane
 * costs what it recorded).  The record is an object of its own, like the reference's:
 * hy_cout_detach hands it to the caller (NULL if nothing was recorded), and it stays valid
 * after the integrator moves on.  hy_cout_free(rec, ctx) gives the pool back to `ctx` for its
 * next recording (no cudaMalloc in steady state), hy_cout_free(rec, NULL) frees it.
 * hy_cout_info: per-lane number of recorded steps and max over lanes.
 * hy_cout_get:  tcs [S, n, order+1, B] and times (hi, lo) [S+1, B], padded
 *               with NaN past each lane's own count (the +1 row: :449-451).
 * hy_cout_eval: out[i, :, :] = x(t[i, :]) for i < k; t is [k, B], out is [k, n, B]. */
int hy_cout_detach(hy_ctx *ctx, hy_cout **out);
int hy_cout_free(hy_cout *rec, hy_ctx *recycle_into);
int hy_cout_info(hy_cout *rec, uint64_t *n_steps, uint64_t *max_steps);
int hy_cout_get(hy_cout *rec, void *tcs, void *times_hi, void *times_lo, uint64_t S);
int hy_cout_eval(hy_cout *rec, const void *t, size_t k, void *out);
/* the same with device pointers (t [k, B] and out [k, n, B] in device memory) */
int hy_cout_eval_dev(hy_cout *rec, const void *d_t, size_t k, void *d_out);

/* Events (taylor_expose_events.cpp:185-317; integrator side
 * expose_batch_integrators.cpp:651-656).  The device appends one record per
 * detected event; the host drains them after each call and dispatches the
 * Python callbacks in chronological order per lane. */
int hy_events_count(hy_ctx *ctx, uint64_t *n);
int hy_events_drain(hy_ctx *ctx, hy_event_rec *recs, uint64_t cap, uint64_t *n);
int hy_get_cooldowns(hy_ctx *ctx, void *elapsed, void *total); /* [B, n_tevents], total<0: none */
int hy_set_cooldowns(hy_ctx *ctx, const void *elapsed, const void *total); /* restore (pickle/copy) */
int hy_reset_cooldowns(hy_ctx *ctx, int64_t lane);             /* -1 = all lanes */

/* Introspection for benches/tests: launch geometry chosen for the tape. */
/* kernel_variant of a kernel generated from the tape and compiled at hy_create time with NVRTC
 * (csrc/hy_jit.hpp): one thread per trajectory, the order sweep as straight-line code.  This is
 * the counterpart of the reference's LLVM JIT (expose_batch_integrators.cpp:166-208). */
#define HY_VARIANT_JIT 1000u
/* Bits of hy_create's `high_accuracy` argument.  HY_CREATE_COMPACT is the reference's
 * `compact_mode=True` (expose_batch_integrators.cpp:198): no per-system code generation - the
 * tape interpreter runs the system. */
#define HY_CREATE_HIGH_ACCURACY 1
#define HY_CREATE_COMPACT 2

typedef struct hy_launch_info {
    uint32_t group;          /* threads cooperating on one trajectory        */
    uint32_t traj_per_cta;   /* trajectories resident per CTA                */
    uint32_t threads;        /* threads per CTA                              */
    uint32_t ctas;           /* persistent CTAs launched                     */
    uint32_t smem_bytes;     /* dynamic shared memory per CTA                */
    uint32_t ws_in_smem;     /* 1: jets in shared memory, 0: global fallback */
    uint32_t n_sm;
    uint32_t regs_per_thread;
    uint32_t kernel_variant; /* 0: tape interpreter; 3..6: register-resident
                                N-body kernel for N bodies (hy_nbody_reg.cuh);
                                7, 8: the same kernel on 32-lane groups, built at
                                hy_create time (NVRTC), as are the builds for
                                parametric masses; 1000: HY_VARIANT_JIT;
                                226: the 6-body FP64 build unrolled to order 22
                                (tol = 1e-18, orders 21..22);
                                222: the FP64 CR3BP build unrolled to order 22;
                                203: register-resident CR3BP kernel
                                (hy_cr3bp_reg.cuh); 106: experimental
                                warpgroup-rotation N-body kernel (HY_CUDA_WGX=1) */
} hy_launch_info;
int hy_get_launch_info(hy_ctx *ctx, hy_launch_info *info);

/* Which kernel hy_create would pick for a tape (no device needed): 0 = the tape
 * interpreter, 3..6 = the register-resident N-body kernel for that many bodies,
 * 203 = the register-resident CR3BP kernel.  The N-body kernel needs the tape of
 * a Newtonian N-body system in Cartesian coordinates (what the reference's
 * model.nbody builds, expose_models.cpp:237-272), no events or parameters, every
 * pair present, order <= 20 (226: 6 bodies, FP64 assumed, orders 21..22)
 * (hy_nbody_match.hpp); the CR3BP kernel exactly the
 * tape of model.cr3bp (expose_models.cpp:395-400) at orders up to 20 (FP64; 222:
 * orders 21..22) or 9 (FP32), no events or parameters (hy_cr3bp_match.hpp). */
int hy_tape_kernel_variant(const hy_dims *dims, const hy_op *ops, const hy_term *terms,
                           uint32_t *variant);

/* Generate and compile the run-time kernel of a tape ahead of time, without a device: fills the
 * on-disk kernel cache (csrc/jit_cache, or $HY_CUDA_JIT_CACHE) that hy_create looks up first.
 * *from_cache: 1 the kernel was already cached, 0 it was compiled now (*compile_s seconds), -1 the
 * tape is served by the interpreter (no kernel is generated for it). */
int hy_jit_precompile(int fp_bits, const hy_tape *full, uint32_t batch, int *from_cache, double *compile_s);
/* The same for the generated event functions of a system that a register-resident kernel serves
 * (`ode`: the ODE-only tape, `evt`: the event tape, as for hy_create2).  *from_cache = -1: no such
 * kernel exists for the system (its events run on the tape interpreter). */
int hy_jit_precompile_events(int fp_bits, const hy_tape *full, const hy_tape *ode, const hy_event_tape *evt,
                             int *from_cache, double *compile_s);

/* DFMA/FFMA peak microbenchmark used as the compute roof (no peak for
 * FP64/FP32 FMA is in MEASURED_PEAKS.json): returns TFLOP/s. */
int hy_measure_fma_peak(int device, int fp_bits, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* HY_CUDA_H */
