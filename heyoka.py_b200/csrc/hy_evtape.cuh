// hy_evtape.cuh - event functions on the register-resident kernels.
//
// The register-resident kernels (hy_nbody_reg.cuh, hy_cr3bp_reg.cuh) keep the jets of the ODE's
// sub-expressions in registers; only the STATE jets are in shared memory.  Event functions
// (reference: t_event_batch / nt_event_batch, /root/reference/heyoka/taylor_expose_events.cpp:185-317)
// are therefore lowered by the host to a SECOND tape - the event functions as functions of the
// state variables alone (hy_b200/decompose.py: decompose_event_tape) - which this file evaluates
// from the state jets after the ODE's orders are complete:
//
//  * event e is served by lane e mod G of the trajectory's group (the per-event tapes share
//    nothing, so no synchronisation is needed inside the evaluation);
//  * every step needs only orders 0, p-1 and p of an event function (they enter the step-size
//    norms).  Products / squares of jets are explicit convolutions of their
//    operands, so an op whose history nobody reads is evaluated at those three orders only
//    (EOF_ALL marks the ops that must run at every order);
//  * after the step size is known, a cheap conservative test decides whether an event can
//    happen in [0, h] at all: interval Horner enclosures of the state polynomials over the
//    step, pushed through the event tape in interval arithmetic (widened by a few ulps per
//    operation).  Only if 0 is inside the enclosure are the remaining orders computed and the
//    root finder of hy_events.cuh run - on the same polynomial the tape interpreter would see.
//
// Jets written by this file are unit-stride rows of the trajectory column (the "event workspace");
// state jets are read with the layout of the kernel that owns them (offset s_srow[i], stride XS).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hy_cuda.h"

namespace hy {

// reference encoding (16 bits): kind in the top two bits
enum : uint16_t { ER_CUR = 0x0000, ER_JET = 0x4000, ER_STATE = 0x8000, ER_ONE = 0xc000, ER_KIND = 0xc000 };
enum : uint8_t { EOF_NEGA = 1, EOF_NEGB = 2, EOF_ALL = 4 };

struct EOp { // 32 bytes
    uint8_t opcode, flags;
    uint16_t n;             // terms
    uint16_t dst, dst2;     // outputs (dst2: SINCOS cosine jet, or the scratch row holding 1/a[0])
    uint16_t a, b;          // operands; for term ops b = first term
    uint16_t imm;           // index into the immediate table (POW exponent)
    uint16_t sd, sd2, sa, sb; // interval slots of dst, dst2, a, b
    uint16_t pad[5];
};
static_assert(sizeof(EOp) == 32, "EOp must be 32 bytes");
struct ETerm { // 16 bytes
    double coef;
    uint16_t src, dst;   // operand / MULSH output
    uint16_t ssrc, sdst; // their interval slots
};
static_assert(sizeof(ETerm) == 16, "ETerm must be 16 bytes");

// Device view of the event tape (the blob is staged in shared memory by the kernel).
struct EvtDev {
    const void *blob;      // [ops | terms | imm (double) | op_start (uint32, n_events + 1) | ev_ref (uint16 -> uint32)]
    uint32_t n_ops, n_terms, n_imm, n_events;
    uint32_t ews_off;      // offset of the event workspace in the trajectory column
    uint32_t eiv_off;      // offset of the interval scratch: 2 x (n_state + n_slots) elements
    uint32_t n_slots;
    uint32_t bytes;        // size of the blob
    unsigned long long *stats; // optional counters (HY_CUDA_EVENT_STATS=1): [0] steps, [1] steps whose
                               // enclosure contained 0 (remaining orders + root finder run)
    // The event workspace + interval scratch outside the trajectory column: a global slab, gstride elements
    // per resident trajectory (ews_off = 0 then).  A step touches a few dozen of its elements (L1 / L2
    // hits), and the shared memory it frees holds more trajectories - the register-resident kernels are
    // latency bound, their rate follows the resident warps.  gws = null: inside the column.
    void *gws;
    uint32_t gstride;
};

template <typename R, int XS> struct EvtCtx {
    const R *w;            // trajectory column
    const uint32_t *srow;  // offset of every state variable (order 0)
    R *ews;                // event workspace
    const R *rk;           // 1/k table
    const double *imm;
    R tm;                  // time at the start of the step (TIME op)
    // pointer to order 0 and stride between orders of a reference
    __device__ __forceinline__ const R *base(uint16_t ref, int &stride) const
    {
        const uint16_t kind = ref & ER_KIND, off = ref & 0x3fff;
        if (kind == ER_STATE) {
            stride = XS;
            return w + srow[off];
        }
        stride = kind == ER_JET ? 1 : 0;
        return ews + off;
    }
    __device__ __forceinline__ R ld(uint16_t ref, uint32_t k) const
    {
        if ((ref & ER_KIND) == ER_ONE) return k == 0 ? (R)1 : (R)0;
        int st;
        const R *p = base(ref, st);
        return p[(int)k * st];
    }
    __device__ __forceinline__ void st(uint16_t ref, uint32_t k, R v) const
    {
        ews[(ref & 0x3fff) + ((ref & ER_KIND) == ER_JET ? k : 0u)] = v;
    }
};

__device__ __forceinline__ double evt_fma(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float evt_fma(float a, float b, float c) { return fmaf(a, b, c); }

// sum_{j < n} pa[j sa] * pb[-j sb], chained like conv<R> (hy_kernels.cuh): four accumulators over
// blocks of eight terms, folded as (s0 + s1) + (s2 + s3).
template <typename R>
static __device__ __noinline__ R evt_conv(const R *__restrict__ pa, int sa, const R *__restrict__ pb, int sb, int n)
{
    R s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll 1
    for (int j0 = 0; j0 < n; j0 += 8) {
        R a[8], b[8];
        const int m = n - j0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const bool v = u < m;
            a[u] = v ? pa[(j0 + u) * sa] : (R)0;
            b[u] = v ? pb[-(j0 + u) * sb] : (R)0;
        }
        s0 = evt_fma(a[0], b[0], s0);
        s1 = evt_fma(a[1], b[1], s1);
        s2 = evt_fma(a[2], b[2], s2);
        s3 = evt_fma(a[3], b[3], s3);
        s0 = evt_fma(a[4], b[4], s0);
        s1 = evt_fma(a[5], b[5], s1);
        s2 = evt_fma(a[6], b[6], s2);
        s3 = evt_fma(a[7], b[7], s3);
    }
    return (s0 + s1) + (s2 + s3);
}

// The same sum with compile-time strides and length (generated event code, hy_jit.hpp: every load has
// an immediate offset, nothing loops or branches).  Term u goes to chain u mod 4, as above.
template <typename R, int SA, int SB, int N>
__device__ __forceinline__ R evt_conv_ct(const R *__restrict__ pa, const R *__restrict__ pb)
{
    R s[4] = {0, 0, 0, 0};
#pragma unroll
    for (int u = 0; u < N; ++u) s[u & 3] = evt_fma(pa[u * SA], pb[-u * SB], s[u & 3]);
    return (s[0] + s[1]) + (s[2] + s[3]);
}

// ---- product units of the generated event code (hy_jit.hpp, EvtGen): ONE lane computes orders 0, p-1
// and p of a square / product from a single pass over its operands' coefficients (registers), and the
// lanes of a trajectory's group work on different products at the same time - same code, operand
// addresses selected by lane.  `al`: the operand is (jet + c), i.e. a LINCOMB "1 * x + c" that nobody
// else reads (it differs from x at order 0 only, so its jet is never materialised).  The sums are
// chained exactly like evt_exec's SQUARE / MUL (term u on chain u mod 4, folded (s0 + s1) + (s2 + s3)).
template <typename R, int K> __device__ __forceinline__ R evt_sq_at(const R *v)
{
    constexpr int half = (K + 1) >> 1;
    R s[4] = {0, 0, 0, 0};
#pragma unroll
    for (int u = 0; u < half; ++u) s[u & 3] = evt_fma(v[u], v[K - u], s[u & 3]);
    R acc = (s[0] + s[1]) + (s[2] + s[3]);
    acc = acc + acc;
    if ((K & 1) == 0) acc = evt_fma(v[K >> 1], v[K >> 1], acc);
    return acc;
}
template <typename R, int K> __device__ __forceinline__ R evt_mul_at(const R *a, const R *b)
{
    R s[4] = {0, 0, 0, 0};
#pragma unroll
    for (int u = 0; u <= K; ++u) s[u & 3] = evt_fma(a[u], b[K - u], s[u & 3]);
    return (s[0] + s[1]) + (s[2] + s[3]);
}
template <typename R, int S, int P>
__device__ __forceinline__ void evt_unit_sq(const R *__restrict__ a, bool al, R c, R *__restrict__ out, bool on)
{
    R v[P + 1];
#pragma unroll
    for (int u = 0; u <= P; ++u) v[u] = a[u * S];
    if (al) v[0] = v[0] + c;
    const R o0 = evt_sq_at<R, 0>(v), o1 = evt_sq_at<R, (P > 0 ? P - 1 : 0)>(v), o2 = evt_sq_at<R, P>(v);
    if (on) {
        out[0] = o0;
        out[1] = o1;
        out[2] = o2;
    }
}
template <typename R, int SA, int SB, int P>
__device__ __forceinline__ void evt_unit_mul(const R *__restrict__ a, bool ala, R ca, const R *__restrict__ b, bool alb,
                                             R cb, R *__restrict__ out, bool on)
{
    R va[P + 1], vb[P + 1];
#pragma unroll
    for (int u = 0; u <= P; ++u) {
        va[u] = a[u * SA];
        vb[u] = b[u * SB];
    }
    if (ala) va[0] = va[0] + ca;
    if (alb) vb[0] = vb[0] + cb;
    const R o0 = evt_mul_at<R, 0>(va, vb), o1 = evt_mul_at<R, (P > 0 ? P - 1 : 0)>(va, vb), o2 = evt_mul_at<R, P>(va, vb);
    if (on) {
        out[0] = o0;
        out[1] = o1;
        out[2] = o2;
    }
}

// sum_{j = j0}^{j1} (wk + j wj) * pa[j sa] * pb[(k - j) sb]  (the recurrences of pow / exp / log / sincos)
template <typename R>
static __device__ __noinline__ R evt_wconv(const R *__restrict__ pa, int sa, const R *__restrict__ pb, int sb, int k,
                                           int j0, int j1, R wk, R wj)
{
    R s0 = 0, s1 = 0;
    R jr = (R)j0;
#pragma unroll 1
    for (int j = j0; j <= j1; j += 2) {
        const bool v1 = j + 1 <= j1;
        const R a0 = pa[j * sa], b0 = pb[(k - j) * sb];
        const R a1 = v1 ? pa[(j + 1) * sa] : (R)0, b1 = v1 ? pb[(k - j - 1) * sb] : (R)0;
        s0 = evt_fma(evt_fma(jr, wj, wk) * a0, b0, s0);
        s1 = evt_fma(evt_fma(jr + (R)1, wj, wk) * a1, b1, s1);
        jr += (R)2;
    }
    return s0 + s1;
}

// math wrappers (defined in hy_kernels.cuh)
template <typename R> __device__ __noinline__ R pow0(R x, double alpha);

// One op of the event tape at order k (same recurrences as exec_op in hy_kernels.cuh).
template <typename R, int XS>
__device__ __forceinline__ void evt_exec(const EOp &o, const ETerm *__restrict__ terms, const EvtCtx<R, XS> &C,
                                         const uint32_t k)
{
    switch (o.opcode) {
    case HY_OP_LINCOMB: {
        R acc = 0;
        for (uint32_t i = 0; i < o.n; ++i) {
            const ETerm t = terms[o.b + i];
            acc = evt_fma((R)t.coef, C.ld(t.src, k), acc);
        }
        C.st(o.dst, k, acc);
    } break;
    case HY_OP_ADDSUB: {
        R a = C.ld(o.a, k), b = C.ld(o.b, k);
        if (o.flags & EOF_NEGA) a = -a;
        if (o.flags & EOF_NEGB) b = -b;
        C.st(o.dst, k, a + b);
    } break;
    case HY_OP_MUL: {
        int sa, sb;
        const R *a = C.base(o.a, sa), *b = C.base(o.b, sb);
        C.st(o.dst, k, evt_conv<R>(a, sa, b + (int)k * sb, sb, (int)k + 1));
    } break;
    case HY_OP_SQUARE: {
        int sa;
        const R *a = C.base(o.a, sa);
        R acc = evt_conv<R>(a, sa, a + (int)k * sa, sa, (int)((k + 1) >> 1));
        acc = acc + acc;
        if ((k & 1u) == 0) {
            const R m = a[(int)(k >> 1) * sa];
            acc = evt_fma(m, m, acc);
        }
        C.st(o.dst, k, acc);
    } break;
    case HY_OP_SUMSQ: {
        const int half = (int)((k + 1) >> 1);
        R acc = 0, acc2 = 0;
        for (uint32_t i = 0; i < o.n; ++i) {
            int sa;
            const R *a = C.base(terms[o.b + i].src, sa);
            acc += evt_conv<R>(a, sa, a + (int)k * sa, sa, half);
            if ((k & 1u) == 0) {
                const R m = a[(int)(k >> 1) * sa];
                acc2 = evt_fma(m, m, acc2);
            }
        }
        C.st(o.dst, k, (acc + acc) + acc2);
    } break;
    case HY_OP_MULSH: {
        int sb;
        const R *b = C.base(o.a, sb);
        for (uint32_t i = 0; i < o.n; ++i) {
            const ETerm t = terms[o.b + i];
            int sa;
            const R *a = C.base(t.src, sa);
            C.st(t.dst, k, evt_conv<R>(a, sa, b + (int)k * sb, sb, (int)k + 1));
        }
    } break;
    case HY_OP_DIV: {
        int sb;
        const R *b = C.base(o.b, sb);
        R *c = C.ews + (o.dst & 0x3fff);
        R *inv = C.ews + (o.dst2 & 0x3fff);
        if (k == 0) *inv = (R)1 / b[0];
        R acc = C.ld(o.a, k);
        if (k > 0) acc -= evt_conv<R>(b + sb, sb, c + k - 1, 1, (int)k);
        c[k] = acc * *inv;
    } break;
    case HY_OP_POW:
    case HY_OP_SQRT: {
        int sa;
        const R *a = C.base(o.a, sa);
        R *c = C.ews + (o.dst & 0x3fff);
        R *inv = C.ews + (o.dst2 & 0x3fff);
        const double alpha = o.opcode == HY_OP_SQRT ? 0.5 : C.imm[o.imm];
        if (k == 0) {
            *inv = (R)1 / a[0];
            c[0] = o.opcode == HY_OP_SQRT ? (R)sqrt((double)a[0]) : pow0<R>(a[0], alpha);
        } else {
            // sum_{j < k} (k alpha - j (alpha + 1)) a[k - j] c[j]
            const R s = evt_wconv<R>(c, 1, a, sa, (int)k, 0, (int)k - 1, (R)k * (R)alpha, (R)(-(alpha + 1.0)));
            c[k] = (s * C.rk[k]) * *inv;
        }
    } break;
    case HY_OP_EXP: {
        int sa;
        const R *a = C.base(o.a, sa);
        R *c = C.ews + (o.dst & 0x3fff);
        if (k == 0)
            c[0] = (R)exp((double)a[0]);
        else
            c[k] = evt_wconv<R>(a, sa, c, 1, (int)k, 1, (int)k, (R)0, (R)1) * C.rk[k];
    } break;
    case HY_OP_LOG: {
        int sa;
        const R *a = C.base(o.a, sa);
        R *c = C.ews + (o.dst & 0x3fff);
        R *inv = C.ews + (o.dst2 & 0x3fff);
        if (k == 0) {
            *inv = (R)1 / a[0];
            c[0] = (R)log((double)a[0]);
        } else {
            const R s = k > 1 ? evt_wconv<R>(c, 1, a, sa, (int)k, 1, (int)k - 1, (R)0, (R)1) : (R)0;
            c[k] = evt_fma(-s, C.rk[k], a[(int)k * sa]) * *inv;
        }
    } break;
    case HY_OP_SINCOS: {
        int sa;
        const R *a = C.base(o.a, sa);
        R *s = C.ews + (o.dst & 0x3fff), *c = C.ews + (o.dst2 & 0x3fff);
        if (k == 0) {
            double sv, cv;
            sincos((double)a[0], &sv, &cv);
            s[0] = (R)sv;
            c[0] = (R)cv;
        } else {
            const R ss = evt_wconv<R>(a, sa, c, 1, (int)k, 1, (int)k, (R)0, (R)1);
            const R cs = evt_wconv<R>(a, sa, s, 1, (int)k, 1, (int)k, (R)0, (R)1);
            s[k] = ss * C.rk[k];
            c[k] = -(cs * C.rk[k]);
        }
    } break;
    case HY_OP_TIME: {
        C.ews[(o.dst & 0x3fff) + k] = k == 0 ? C.tm : (k == 1 ? (R)1 : (R)0);
    } break;
    default: break;
    }
}

// One op at EVERY order 0..p (pass A of the event evaluation).  Linear combinations - the common
// case: shifted coordinates such as x - mu - keep their operand pointers in registers and run one
// tight loop over the orders; everything else goes through evt_exec order by order.
template <typename R, int XS>
__device__ __forceinline__ void evt_exec_all(const EOp &o, const ETerm *__restrict__ terms, const EvtCtx<R, XS> &C,
                                             const uint32_t p)
{
    if (o.opcode == HY_OP_LINCOMB && o.n <= 4 && (o.dst & ER_KIND) == ER_JET) {
        const R *src[4];
        int st[4];
        R cf[4], c0 = 0; // c0: the constant (terms on the unit jet) - enters at order 0 only
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            src[i] = C.rk + 1; // (a finite value: unused slots contribute 0 * 1)
            st[i] = 0;
            cf[i] = 0;
        }
        bool ok = true;
        for (uint32_t i = 0; i < o.n; ++i) {
            const ETerm t = terms[o.b + i];
            if ((t.src & ER_KIND) == ER_ONE) {
                // exec_op adds the terms in order with an FMA; a constant in any position other than the
                // last would change the rounding of the order-0 sum: fall back to the generic path then
                if (i + 1 != o.n) ok = false;
                c0 = (R)t.coef;
            } else {
                src[i] = C.base(t.src, st[i]);
                cf[i] = (R)t.coef;
                if (st[i] == 0) ok = false; // (a single-row operand: generic path)
            }
        }
        if (ok) {
            R *dst = C.ews + (o.dst & 0x3fff);
#pragma unroll 1
            for (uint32_t k = 0; k <= p; ++k) {
                R acc = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) acc = evt_fma(cf[i], src[i][(int)k * st[i]], acc);
                dst[k] = k == 0 ? evt_fma(c0, (R)1, acc) : acc;
            }
            return;
        }
    }
#pragma unroll 1
    for (uint32_t k = 0; k <= p; ++k) evt_exec<R, XS>(o, terms, C, k);
}

// ---- interval arithmetic (round-to-nearest + widening: every result is pushed outwards by 4 ulps
// of its magnitude, which covers the rounding of the few operations that formed it) ----
template <typename R> struct Ival {
    R lo, hi;
};
template <typename R> __device__ __forceinline__ Ival<R> iv_widen(R lo, R hi)
{
    const R e = (R)4 * (sizeof(R) == 8 ? (R)2.220446049250313e-16 : (R)1.1920929e-07f);
    const R m = (fabs(lo) > fabs(hi) ? fabs(lo) : fabs(hi)) * e + (sizeof(R) == 8 ? (R)1e-300 : (R)1e-37f);
    return Ival<R>{lo - m, hi + m};
}
template <typename R> __device__ __forceinline__ Ival<R> iv_mul(Ival<R> a, Ival<R> b)
{
    const R p0 = a.lo * b.lo, p1 = a.lo * b.hi, p2 = a.hi * b.lo, p3 = a.hi * b.hi;
    const R lo = fmin(fmin(p0, p1), fmin(p2, p3)), hi = fmax(fmax(p0, p1), fmax(p2, p3));
    return iv_widen<R>(lo, hi);
}
template <typename R> __device__ __forceinline__ Ival<R> iv_sqr(Ival<R> a)
{
    const R l2 = a.lo * a.lo, h2 = a.hi * a.hi;
    if (a.lo >= (R)0) return iv_widen<R>(l2, h2);
    if (a.hi <= (R)0) return iv_widen<R>(h2, l2);
    return iv_widen<R>((R)0, fmax(l2, h2));
}
template <typename R> __device__ __forceinline__ Ival<R> iv_all()
{
    const R inf = sizeof(R) == 8 ? (R)__longlong_as_double(0x7ff0000000000000LL) : (R)__int_as_float(0x7f800000);
    return Ival<R>{-inf, inf};
}

// Enclosure of sum_k c[k stride] tau^k over tau in [0, h] (h of either sign) by interval Horner.
template <typename R> __device__ __forceinline__ Ival<R> iv_horner(const R *c, int stride, int p, R h)
{
    R lo = c[p * stride], hi = lo;
    for (int k = p - 1; k >= 0; --k) {
        const R a = lo * h, b = hi * h; // (h < 0 swaps the ends)
        const R mn = fmin(a, b), mx = fmax(a, b);
        const R ck = c[k * stride];
        lo = fmin(mn, (R)0) + ck;
        hi = fmax(mx, (R)0) + ck;
    }
    return iv_widen<R>(lo, hi);
}

// A cheaper enclosure of the same polynomial (generated event code, hy_jit.hpp): the first-order term
// exactly, the rest by absolute values - c0 + [min(0, c1 h), max(0, c1 h)] -+ sum_{k>=2} |c_k| |h|^k.
// One FMA per coefficient instead of the ~12 instructions of an interval Horner step (FP64 min / max
// are compare + select pairs); looser, but only the rare full evaluation pays for that.
template <typename R, int S, int P> __device__ __forceinline__ Ival<R> iv_taylor_abs(const R *c, R h)
{
    const R ah = fabs(h);
    R r = fabs(c[P * S]);
#pragma unroll
    for (int k = P - 1; k >= 2; --k) r = evt_fma(r, ah, fabs(c[k * S]));
    r = r * ah * ah; // sum_{k>=2} |c_k| |h|^k
    const R d1 = c[S] * h, c0 = c[0];
    const R lo = c0 + fmin(d1, (R)0) - r, hi = c0 + fmax(d1, (R)0) + r;
    return iv_widen<R>(lo, hi);
}

// Pieces of evt_interval for the generated straight-line form (intervals in registers, literal coefficients).
template <typename R> __device__ __forceinline__ Ival<R> iv_fix(Ival<R> v)
{
    if (!(v.lo == v.lo) || !(v.hi == v.hi)) v = iv_all<R>();
    return v;
}
template <typename R> __device__ __forceinline__ void iv_lin_term(R &lo, R &hi, R c, Ival<R> x)
{
    lo += c >= (R)0 ? c * x.lo : c * x.hi;
    hi += c >= (R)0 ? c * x.hi : c * x.lo;
}
template <typename R> __device__ __forceinline__ Ival<R> iv_lin_finish(R lo, R hi, int n)
{
    const R e = (R)(4 + 2 * n);
    Ival<R> r = iv_widen<R>(lo, hi);
    const R m = (fabs(lo) + fabs(hi)) * e * (sizeof(R) == 8 ? (R)2.3e-16 : (R)1.2e-07f);
    r.lo -= m;
    r.hi += m;
    return iv_fix<R>(r);
}

// Interval image of one op given the intervals of its operands in iv[] (slot-indexed: slots
// 0..n_state-1 hold the state enclosures, the ops' outputs follow).  NaNs propagate to "unknown".
template <typename R>
__device__ __forceinline__ void evt_interval(const EOp &o, const ETerm *__restrict__ terms, R *iv, const double *imm,
                                             R t0, R h)
{
    auto get = [&](uint16_t ref, uint16_t slot) -> Ival<R> {
        if ((ref & ER_KIND) == ER_ONE) return Ival<R>{(R)1, (R)1};
        return Ival<R>{iv[2 * slot], iv[2 * slot + 1]};
    };
    auto put = [&](uint16_t slot, Ival<R> v) {
        if (!(v.lo == v.lo) || !(v.hi == v.hi)) v = iv_all<R>();
        iv[2 * slot] = v.lo;
        iv[2 * slot + 1] = v.hi;
    };
    switch (o.opcode) {
    case HY_OP_LINCOMB: {
        R lo = 0, hi = 0;
        for (uint32_t i = 0; i < o.n; ++i) {
            const ETerm t = terms[o.b + i];
            const Ival<R> x = get(t.src, t.ssrc);
            const R c = (R)t.coef;
            lo += c >= (R)0 ? c * x.lo : c * x.hi;
            hi += c >= (R)0 ? c * x.hi : c * x.lo;
        }
        const R e = (R)(4 + 2 * (int)o.n);
        Ival<R> r = iv_widen<R>(lo, hi);
        r.lo -= (fabs(lo) + fabs(hi)) * e * (sizeof(R) == 8 ? (R)2.3e-16 : (R)1.2e-07f);
        r.hi += (fabs(lo) + fabs(hi)) * e * (sizeof(R) == 8 ? (R)2.3e-16 : (R)1.2e-07f);
        put(o.sd, r);
    } break;
    case HY_OP_ADDSUB: {
        Ival<R> a = get(o.a, o.sa), b = get(o.b, o.sb);
        if (o.flags & EOF_NEGA) a = Ival<R>{-a.hi, -a.lo};
        if (o.flags & EOF_NEGB) b = Ival<R>{-b.hi, -b.lo};
        put(o.sd, iv_widen<R>(a.lo + b.lo, a.hi + b.hi));
    } break;
    case HY_OP_MUL: put(o.sd, iv_mul<R>(get(o.a, o.sa), get(o.b, o.sb))); break;
    case HY_OP_SQUARE: put(o.sd, iv_sqr<R>(get(o.a, o.sa))); break;
    case HY_OP_SUMSQ: {
        R lo = 0, hi = 0;
        for (uint32_t i = 0; i < o.n; ++i) {
            const ETerm t = terms[o.b + i];
            const Ival<R> s = iv_sqr<R>(get(t.src, t.ssrc));
            lo += s.lo;
            hi += s.hi;
        }
        put(o.sd, iv_widen<R>(lo, hi));
    } break;
    case HY_OP_MULSH: {
        const Ival<R> b = get(o.a, o.sa);
        for (uint32_t i = 0; i < o.n; ++i) {
            const ETerm t = terms[o.b + i];
            put(t.sdst, iv_mul<R>(get(t.src, t.ssrc), b));
        }
    } break;
    case HY_OP_DIV: {
        const Ival<R> a = get(o.a, o.sa), b = get(o.b, o.sb);
        if (b.lo > (R)0 || b.hi < (R)0)
            put(o.sd, iv_mul<R>(a, iv_widen<R>((R)1 / b.hi, (R)1 / b.lo)));
        else
            put(o.sd, iv_all<R>());
    } break;
    case HY_OP_SQRT: {
        const Ival<R> a = get(o.a, o.sa);
        if (a.lo >= (R)0)
            put(o.sd, iv_widen<R>((R)sqrt((double)a.lo), (R)sqrt((double)a.hi)));
        else
            put(o.sd, iv_all<R>());
    } break;
    case HY_OP_POW: {
        const Ival<R> a = get(o.a, o.sa);
        const double al = imm[o.imm];
        if (a.lo > (R)0) {
            const R x = (R)pow((double)a.lo, al), y = (R)pow((double)a.hi, al);
            Ival<R> r = iv_widen<R>(fmin(x, y), fmax(x, y));
            r = iv_widen<R>(r.lo, r.hi); // (libm pow: a second widening)
            put(o.sd, r);
        } else
            put(o.sd, iv_all<R>());
    } break;
    case HY_OP_EXP: {
        const Ival<R> a = get(o.a, o.sa);
        Ival<R> r = iv_widen<R>((R)exp((double)a.lo), (R)exp((double)a.hi));
        put(o.sd, iv_widen<R>(r.lo, r.hi));
    } break;
    case HY_OP_LOG: {
        const Ival<R> a = get(o.a, o.sa);
        if (a.lo > (R)0) {
            Ival<R> r = iv_widen<R>((R)log((double)a.lo), (R)log((double)a.hi));
            put(o.sd, iv_widen<R>(r.lo, r.hi));
        } else
            put(o.sd, iv_all<R>());
    } break;
    case HY_OP_SINCOS: {
        put(o.sd, Ival<R>{(R)-1, (R)1});
        put(o.sd2, Ival<R>{(R)-1, (R)1});
    } break;
    case HY_OP_TIME: {
        const R t1 = t0 + h;
        put(o.sd, iv_widen<R>(fmin(t0, t1), fmax(t0, t1)));
    } break;
    default: break;
    }
}

} // namespace hy
