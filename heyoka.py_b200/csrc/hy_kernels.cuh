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
//    group and levels are separated by __syncwarp(group mask).
//
// This replaces the reference's JIT-compiled taylor_step + C++ propagate loop
// (/root/reference/heyoka/expose_batch_integrators.cpp:233-314 -> [UPSTREAM]).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hy_cuda.h"

namespace hy {

// ---- precision-generic math wrappers ----
__device__ __forceinline__ double r_fma(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float r_fma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double r_sqrt(double a) { return sqrt(a); }
__device__ __forceinline__ float r_sqrt(float a) { return sqrtf(a); }
__device__ __forceinline__ double r_pow(double a, double b) { return pow(a, b); }
__device__ __forceinline__ float r_pow(float a, float b) { return powf(a, b); }
__device__ __forceinline__ double r_exp(double a) { return exp(a); }
__device__ __forceinline__ float r_exp(float a) { return expf(a); }
__device__ __forceinline__ double r_log(double a) { return log(a); }
__device__ __forceinline__ float r_log(float a) { return logf(a); }
__device__ __forceinline__ void r_sincos(double a, double *s, double *c) { sincos(a, s, c); }
__device__ __forceinline__ void r_sincos(float a, float *s, float *c) { sincosf(a, s, c); }
__device__ __forceinline__ double r_abs(double a) { return fabs(a); }
__device__ __forceinline__ float r_abs(float a) { return fabsf(a); }
__device__ __forceinline__ double r_copysign(double a, double b) { return copysign(a, b); }
__device__ __forceinline__ float r_copysign(float a, float b) { return copysignf(a, b); }
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

enum { MODE_UNTIL = 0, MODE_FOR = 1, MODE_STEP = 2 };

template <typename R> struct KParams {
    hy_dims d;
    const hy_op *ops;
    const hy_term *terms;
    const uint32_t *level_start;
    const uint32_t *ev_ref;
    R *state;      // [n][B]
    const R *pars; // [m][B]
    R *t_hi, *t_lo, *last_h;
    const R *tf;  // [B] final time / delta (unused in MODE_STEP)
    const R *mdt; // [B] or nullptr
    long long *outcome;
    R *min_h, *max_h;
    unsigned long long *n_steps;
    R *tc; // [n][p+1][B] or nullptr
    unsigned int *counter;
    R *gws; // global workspace fallback (ws_in_smem == 0)
    uint32_t B, T, TS;
    unsigned long long max_steps;
    int mode, backward, write_tc, high_accuracy, ws_in_smem;
    R rhofac, inv_p, inv_pm1;
};

// Row reference -> row index at order k.
__device__ __forceinline__ uint32_t ref_row(uint32_t ref, uint32_t k) { return (ref & 0x7fffffffu) + (ref >> 31) * k; }

template <typename R> __device__ __forceinline__ R pow0(R x, double alpha)
{
    if (alpha == -1.5) return (R)1 / (x * r_sqrt(x));
    if (alpha == -0.5) return (R)1 / r_sqrt(x);
    if (alpha == 1.5) return x * r_sqrt(x);
    if (alpha == -1.0) return (R)1 / x;
    if (alpha == -2.0) return (R)1 / (x * x);
    return r_pow(x, (R)alpha);
}

// One op of the tape at order k for the trajectory whose column is `w`
// (w = ws + t; element of row r is w[r * TS]).
template <typename R>
__device__ __forceinline__ void exec_op(const hy_op &o, const hy_term *__restrict__ terms, R *w, const uint32_t TS,
                                        const R *__restrict__ rk, const uint32_t k, const R tm, const uint32_t par_row)
{
#define ROW(r) w[(size_t)(r) * TS]
    switch (o.opcode) {
    case HY_OP_LINCOMB: {
        R acc = 0;
        const hy_term *t = terms + o.b;
        for (uint32_t i = 0; i < o.n; ++i) {
            R c = (R)t[i].coef;
            if (t[i].par >= 0) c = c * ROW(par_row + t[i].par);
            const uint32_t src = t[i].src;
            R v = (src == HY_REF_ONE) ? (k == 0 ? (R)1 : (R)0) : ROW(ref_row(src, k));
            acc = r_fma(c, v, acc);
        }
        ROW(ref_row(o.dst, k)) = acc;
    } break;
    case HY_OP_MUL: {
        const R *a = &ROW(o.a & 0x7fffffffu), *b = &ROW((o.b & 0x7fffffffu) + k);
        R acc = 0;
        for (uint32_t j = 0; j <= k; ++j) acc = r_fma(a[(size_t)j * TS], b[-(ptrdiff_t)((size_t)j * TS)], acc);
        ROW(ref_row(o.dst, k)) = acc;
    } break;
    case HY_OP_SQUARE: {
        const R *a = &ROW(o.a & 0x7fffffffu), *b = a + (size_t)k * TS;
        R acc = 0;
        const uint32_t half = (k + 1) >> 1;
        for (uint32_t j = 0; j < half; ++j) acc = r_fma(a[(size_t)j * TS], b[-(ptrdiff_t)((size_t)j * TS)], acc);
        acc = acc + acc;
        if ((k & 1u) == 0) {
            R m = a[(size_t)(k >> 1) * TS];
            acc = r_fma(m, m, acc);
        }
        ROW(ref_row(o.dst, k)) = acc;
    } break;
    case HY_OP_SUMSQ: {
        R acc = 0;
        const uint32_t half = (k + 1) >> 1;
        const hy_term *t = terms + o.b;
        for (uint32_t i = 0; i < o.n; ++i) {
            const R *a = &ROW(t[i].src & 0x7fffffffu), *b = a + (size_t)k * TS;
            for (uint32_t j = 0; j < half; ++j)
                acc = r_fma(a[(size_t)j * TS], b[-(ptrdiff_t)((size_t)j * TS)], acc);
        }
        acc = acc + acc;
        if ((k & 1u) == 0)
            for (uint32_t i = 0; i < o.n; ++i) {
                R m = ROW((t[i].src & 0x7fffffffu) + (k >> 1));
                acc = r_fma(m, m, acc);
            }
        ROW(ref_row(o.dst, k)) = acc;
    } break;
    case HY_OP_MULSH: {
        const R *b = &ROW((o.a & 0x7fffffffu) + k);
        const hy_term *t = terms + o.b;
        for (uint32_t i = 0; i < o.n; ++i) {
            const R *a = &ROW(t[i].src & 0x7fffffffu);
            R acc = 0;
            for (uint32_t j = 0; j <= k; ++j)
                acc = r_fma(a[(size_t)j * TS], b[-(ptrdiff_t)((size_t)j * TS)], acc);
            ROW(ref_row(t[i].dst, k)) = acc;
        }
    } break;
    case HY_OP_DIV: {
        const R *b = &ROW(o.b & 0x7fffffffu);
        R *c = &ROW(o.dst & 0x7fffffffu);
        if (k == 0) ROW(o.dst2) = (R)1 / b[0];
        R acc = ROW(ref_row(o.a, k));
        for (uint32_t j = 1; j <= k; ++j) acc = r_fma(-b[(size_t)j * TS], c[(size_t)(k - j) * TS], acc);
        c[(size_t)k * TS] = acc * ROW(o.dst2);
    } break;
    case HY_OP_POW:
    case HY_OP_SQRT: {
        const R *a = &ROW(o.a & 0x7fffffffu);
        R *c = &ROW(o.dst & 0x7fffffffu);
        const double alpha = o.opcode == HY_OP_SQRT ? 0.5 : o.imm;
        if (k == 0) {
            ROW(o.dst2) = (R)1 / a[0];
            c[0] = o.opcode == HY_OP_SQRT ? r_sqrt(a[0]) : pow0<R>(a[0], alpha);
        } else {
            const R al = (R)alpha, al1 = (R)(alpha + 1.0), kal = (R)k * al;
            R acc = 0;
            const R *ak = a + (size_t)k * TS;
            for (uint32_t j = 0; j < k; ++j) {
                const R wgt = r_fma(-(R)j, al1, kal);
                acc = r_fma(wgt * ak[-(ptrdiff_t)((size_t)j * TS)], c[(size_t)j * TS], acc);
            }
            c[(size_t)k * TS] = (acc * rk[k]) * ROW(o.dst2);
        }
    } break;
    case HY_OP_EXP: {
        const R *a = &ROW(o.a & 0x7fffffffu);
        R *c = &ROW(o.dst & 0x7fffffffu);
        if (k == 0) {
            c[0] = r_exp(a[0]);
        } else {
            R acc = 0;
            for (uint32_t j = 1; j <= k; ++j) acc = r_fma((R)j * a[(size_t)j * TS], c[(size_t)(k - j) * TS], acc);
            c[(size_t)k * TS] = acc * rk[k];
        }
    } break;
    case HY_OP_LOG: {
        const R *a = &ROW(o.a & 0x7fffffffu);
        R *c = &ROW(o.dst & 0x7fffffffu);
        if (k == 0) {
            ROW(o.dst2) = (R)1 / a[0];
            c[0] = r_log(a[0]);
        } else {
            R acc = 0;
            for (uint32_t j = 1; j < k; ++j) acc = r_fma((R)j * c[(size_t)j * TS], a[(size_t)(k - j) * TS], acc);
            c[(size_t)k * TS] = r_fma(-acc, rk[k], a[(size_t)k * TS]) * ROW(o.dst2);
        }
    } break;
    case HY_OP_SINCOS: {
        const R *a = &ROW(o.a & 0x7fffffffu);
        R *s = &ROW(o.dst & 0x7fffffffu), *c = &ROW(o.dst2 & 0x7fffffffu);
        if (k == 0) {
            R sv, cv;
            r_sincos(a[0], &sv, &cv);
            s[0] = sv;
            c[0] = cv;
        } else {
            R sa = 0, ca = 0;
            for (uint32_t j = 1; j <= k; ++j) {
                const R ja = (R)j * a[(size_t)j * TS];
                sa = r_fma(ja, c[(size_t)(k - j) * TS], sa);
                ca = r_fma(ja, s[(size_t)(k - j) * TS], ca);
            }
            s[(size_t)k * TS] = sa * rk[k];
            c[(size_t)k * TS] = -(ca * rk[k]);
        }
    } break;
    case HY_OP_TIME: {
        ROW((o.dst & 0x7fffffffu) + k) = k == 0 ? tm : (k == 1 ? (R)1 : (R)0);
    } break;
    case HY_OP_SVD: {
        ROW((o.dst & 0x7fffffffu) + k + 1) = ROW(ref_row(o.a, k)) * rk[k + 1];
    } break;
    default: break;
    }
#undef ROW
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
//   [ops | terms | level_start | ev_ref | rk | ws]
struct SmemLayout {
    uint32_t off_ops, off_terms, off_levels, off_ev, off_rk, off_ws, total;
};

__host__ __device__ inline uint32_t align_up(uint32_t x, uint32_t a) { return (x + a - 1) / a * a; }

__host__ __device__ inline SmemLayout make_layout(const hy_dims &d, uint32_t TS, uint32_t real_bytes, int ws_in_smem)
{
    SmemLayout L;
    uint32_t o = 0;
    L.off_ops = o;
    o += d.n_ops * (uint32_t)sizeof(hy_op);
    L.off_terms = o;
    o += d.n_terms * (uint32_t)sizeof(hy_term);
    o = align_up(o, 8);
    L.off_levels = o;
    o += (d.n_levels + 1) * 4;
    L.off_ev = o;
    o += d.n_events * 4;
    o = align_up(o, 8);
    L.off_rk = o;
    o += (d.order + 2) * real_bytes;
    o = align_up(o, 16);
    L.off_ws = o;
    if (ws_in_smem) o += (d.n_rows + d.n_par) * TS * real_bytes;
    L.total = o;
    return L;
}

template <typename R, int G> __global__ void __launch_bounds__(512, 1) propagate_kernel(const KParams<R> P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const hy_dims &d = P.d;
    const SmemLayout L = make_layout(d, P.TS, sizeof(R), P.ws_in_smem);
    hy_op *s_ops = reinterpret_cast<hy_op *>(smem_raw + L.off_ops);
    hy_term *s_terms = reinterpret_cast<hy_term *>(smem_raw + L.off_terms);
    uint32_t *s_levels = reinterpret_cast<uint32_t *>(smem_raw + L.off_levels);
    uint32_t *s_ev = reinterpret_cast<uint32_t *>(smem_raw + L.off_ev);
    R *s_rk = reinterpret_cast<R *>(smem_raw + L.off_rk);
    R *ws = P.ws_in_smem ? reinterpret_cast<R *>(smem_raw + L.off_ws)
                         : P.gws + (size_t)blockIdx.x * (d.n_rows + d.n_par) * P.TS;

    // ---- stage the tape ----
    {
        const uint32_t nw_ops = d.n_ops * (uint32_t)sizeof(hy_op) / 4;
        const uint32_t *src = reinterpret_cast<const uint32_t *>(P.ops);
        uint32_t *dst = reinterpret_cast<uint32_t *>(s_ops);
        for (uint32_t i = threadIdx.x; i < nw_ops; i += blockDim.x) dst[i] = src[i];
        const uint32_t nw_t = d.n_terms * (uint32_t)sizeof(hy_term) / 4;
        src = reinterpret_cast<const uint32_t *>(P.terms);
        dst = reinterpret_cast<uint32_t *>(s_terms);
        for (uint32_t i = threadIdx.x; i < nw_t; i += blockDim.x) dst[i] = src[i];
        for (uint32_t i = threadIdx.x; i <= d.n_levels; i += blockDim.x) s_levels[i] = P.level_start[i];
        for (uint32_t i = threadIdx.x; i < d.n_events; i += blockDim.x) s_ev[i] = P.ev_ref[i];
        for (uint32_t i = threadIdx.x; i < d.order + 2; i += blockDim.x)
            s_rk[i] = i == 0 ? (R)0 : (R)(1.0 / (double)i);
    }
    __syncthreads();

    const uint32_t TS = P.TS;
    const uint32_t p = d.order, P1 = p + 1, n = d.n_state;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t sub = threadIdx.x & (G - 1);
    const uint32_t slot = threadIdx.x / G; // trajectory slot in this CTA
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(uint32_t)(G - 1)));
    const uint32_t par_row = d.n_rows;
    if (slot >= P.T) return; // whole groups only: safe w.r.t. group-mask syncs
    R *w = ws + slot;

    for (;;) {
        // ---- fetch the next trajectory for this group ----
        unsigned int traj = 0;
        if (sub == 0) traj = atomicAdd(P.counter, 1u);
        if (G > 1) traj = __shfl_sync(gmask, traj, 0, G);
        if (traj >= P.B) break;

        for (uint32_t i = sub; i < n; i += G) w[(size_t)(i * P1) * TS] = P.state[(size_t)i * P.B + traj];
        for (uint32_t i = sub; i < d.n_par; i += G) w[(size_t)(par_row + i) * TS] = P.pars[(size_t)i * P.B + traj];
        R hi = P.t_hi[traj], lo = P.t_lo[traj];
        R mdt = P.mdt ? P.mdt[traj] : r_inf<R>();
        R tf_hi = 0, tf_lo = 0;
        if (P.mode == MODE_FOR) {
            tf_hi = hi;
            tf_lo = lo;
            time_add(tf_hi, tf_lo, P.tf[traj]);
        } else if (P.mode == MODE_UNTIL) {
            tf_hi = P.tf[traj];
        }
        if (P.mode != MODE_STEP) mdt = r_abs(mdt);
        long long oc = HY_OUTCOME_TIME_LIMIT;
        R mn = r_inf<R>(), mx = 0, h = 0;
        unsigned long long ns = 0;
        if (G > 1) __syncwarp(gmask);

        for (;;) {
            R rem = 0, lim;
            if (P.mode == MODE_STEP) {
                lim = P.mdt ? mdt : (P.backward ? -r_inf<R>() : r_inf<R>());
            } else {
                rem = time_sub(tf_hi, tf_lo, hi, lo);
                if (rem == (R)0) {
                    oc = HY_OUTCOME_TIME_LIMIT;
                    break;
                }
                lim = r_abs(rem) < mdt ? rem : r_copysign(mdt, rem);
            }

            // ---- jets: orders 0..p-1 of every op, then x[k+1] ----
            for (uint32_t k = 0; k < p; ++k) {
                for (uint32_t lv = 0; lv < d.n_levels; ++lv) {
                    const uint32_t e = s_levels[lv + 1];
                    for (uint32_t i = s_levels[lv] + sub; i < e; i += G)
                        exec_op<R>(s_ops[i], s_terms, w, TS, s_rk, k, hi, par_row);
                    if (G > 1) __syncwarp(gmask);
                }
            }
            if (d.n_events) {
                for (uint32_t lv = 0; lv < d.n_levels; ++lv) {
                    const uint32_t e = s_levels[lv + 1];
                    for (uint32_t i = s_levels[lv] + sub; i < e; i += G)
                        if ((s_ops[i].flags & HY_OPF_EVENT) && s_ops[i].opcode != HY_OP_SVD)
                            exec_op<R>(s_ops[i], s_terms, w, TS, s_rk, p, hi, par_row);
                    if (G > 1) __syncwarp(gmask);
                }
            }

            // ---- step size ----
            R n0 = 0, n1 = 0, n2 = 0;
            for (uint32_t i = sub; i < n + d.n_events; i += G) {
                const R *x = i < n ? &w[(size_t)(i * P1) * TS] : &w[(size_t)(s_ev[i - n] & 0x7fffffffu) * TS];
                n0 = nan_max(n0, r_abs(x[0]));
                n1 = nan_max(n1, r_abs(x[(size_t)(p - 1) * TS]));
                n2 = nan_max(n2, r_abs(x[(size_t)p * TS]));
            }
#pragma unroll
            for (int m = G >> 1; m > 0; m >>= 1) {
                n0 = nan_max(n0, shfl_xor<R>(gmask, n0, m));
                n1 = nan_max(n1, shfl_xor<R>(gmask, n1, m));
                n2 = nan_max(n2, shfl_xor<R>(gmask, n2, m));
            }
            if (n0 != n0 || n1 != n1 || n2 != n2) {
                h = n0 + n1 + n2; // NaN
            } else {
                const R num = n0 < (R)1 ? (R)1 : n0;
                const R rho_p = r_pow(num / n2, P.inv_p);
                const R rho_pm1 = r_pow(num / n1, P.inv_pm1);
                h = (rho_p < rho_pm1 ? rho_p : rho_pm1) * P.rhofac;
            }
            if (signbit(lim)) h = -h;
            long long so = HY_OUTCOME_SUCCESS;
            if (r_abs(h) > r_abs(lim)) {
                h = lim;
                so = HY_OUTCOME_TIME_LIMIT;
            }

            // ---- optional tc write, then the state update ----
            if (P.write_tc && P.tc) {
                for (uint32_t i = sub; i < n * P1; i += G) P.tc[(size_t)i * P.B + traj] = w[(size_t)i * TS];
                if (G > 1) __syncwarp(gmask);
            }
            bool finite = true;
            for (uint32_t i = sub; i < n; i += G) {
                R *x = &w[(size_t)(i * P1) * TS];
                R acc;
                if (!P.high_accuracy) {
                    acc = x[(size_t)p * TS];
                    for (uint32_t k = p; k-- > 0;) acc = r_fma(acc, h, x[(size_t)k * TS]);
                } else {
                    R sum = x[0], comp = 0, hk = h;
                    for (uint32_t k = 1; k <= p; ++k) {
                        const R term = x[(size_t)k * TS] * hk; // single rounding (no fma partner)
                        const R y = ef_sub(term, comp);
                        const R tt = ef_add(sum, y);
                        comp = ef_sub(ef_sub(tt, sum), y);
                        sum = tt;
                        hk = hk * h;
                    }
                    acc = sum;
                }
                finite = finite && (r_abs(acc) < r_inf<R>());
                x[0] = acc;
            }
            if (G > 1) finite = !__any_sync(gmask, !finite);
            time_add(hi, lo, h);
            ++ns;
            if (!finite) so = HY_OUTCOME_ERR_NF_STATE;

            if (P.mode == MODE_STEP) {
                oc = so;
                break;
            }
            if (so == HY_OUTCOME_ERR_NF_STATE) {
                oc = so;
                break;
            }
            if (so == HY_OUTCOME_SUCCESS) {
                const R ah = r_abs(h);
                if (ah < mn) mn = ah;
                if (ah > mx) mx = ah;
            }
            if (so == HY_OUTCOME_TIME_LIMIT && h == rem) {
                hi = tf_hi;
                lo = tf_lo;
                oc = HY_OUTCOME_TIME_LIMIT;
                break;
            }
            if (P.max_steps && ns >= P.max_steps) {
                oc = HY_OUTCOME_STEP_LIMIT;
                break;
            }
            if (G > 1) __syncwarp(gmask);
        }

        // ---- retire the trajectory ----
        if (G > 1) __syncwarp(gmask);
        for (uint32_t i = sub; i < n; i += G) P.state[(size_t)i * P.B + traj] = w[(size_t)(i * P1) * TS];
        if (sub == 0) {
            P.t_hi[traj] = hi;
            P.t_lo[traj] = lo;
            P.last_h[traj] = h;
            P.outcome[traj] = oc;
            P.min_h[traj] = mn;
            P.max_h[traj] = mx;
            P.n_steps[traj] = ns;
        }
        if (G > 1) __syncwarp(gmask);
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
