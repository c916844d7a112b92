// hy_nbody_reg.cuh - REGISTER-RESIDENT jets for pairwise central-force tapes
// (the Newtonian N-body problem, BASELINE config 2).
//
// The tape interpreter in hy_kernels.cuh reads every convolution operand from
// shared memory (1.3 loads per DFMA), which caps it at ~1/8 of the FP64 pipe.
// When hy_create recognises the tape of an N-body system (hy_nbody_match.hpp)
// the same persistent kernel computes the jets with this file instead:
//
//  * lane s of the G-lane group owns body pair s = (A, B):  d = x_A - x_B,
//    r2 = d.d, w = r2^(-3/2), t = d*w.  The jets of d (3 components), r2 and w
//    live in REGISTERS for the whole step: the order loop is fully unrolled
//    (compile-time K), so every operand is a register and every weight of the
//    power recurrence an immediate.  5(p-1) values per lane = 190 registers in
//    FP64 at order 20.
//  * the two trajectories of a warp step in lockstep.  Lanes 0..2NB-1 of the warp
//    double as BODY lanes (all in one half-warp: a 64-bit shared access costs one
//    wavefront per active half-warp): after each order they gather the NB-1 pair
//    products of their body from the pair lanes' registers by warp shuffle (the
//    shared-memory exchange buffer it replaced is kept under
//    -DHY_NBR_SMEM_EXCHANGE for A/B measurements), form the acceleration (the
//    tape's LINCOMB, same term order), and write v[k+1] and x[k+2] into the state
//    jets, which live in shared memory (the pair lanes read x[k+1] from there, the
//    Horner update the whole history).  Vectors move as one 128-bit + one 64-bit
//    access.
//  * ONE __syncwarp per order (for the positions): t[k] feeds d[k+2], not d[k+1],
//    so the gather of order k overlaps the convolutions of order k+1; the body
//    lanes' work is predicated, not branched, so it shares a basic block with the
//    next order's convolutions and ptxas interleaves the two.
//
// Arithmetic: the same recurrences, term order and roundings as the tape
// interpreter's fused pair op (pair3_k) + LINCOMB + SVD - the two paths agree
// bit for bit (tests/test_gpu_nbody_reg.py).
//
// Reference path replaced: the JIT-compiled taylor_step of
// hey::taylor_adaptive_batch<T> ([UPSTREAM], called from
// /root/reference/heyoka/expose_batch_integrators.cpp:233-314).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hy {

constexpr int NBR_PMAX = 20;   // Taylor order of the fully unrolled register-resident kernels (tol = eps64)
constexpr int NBR_LMAX = 22;   // rows of the column layout; also the order of the 6-body FP64 high-accuracy
                               // build (tol = 1e-18: the reference's own benchmark, ensemble_batch_perf.ipynb)
constexpr int NBR_VARIANT_P22 = 226; // hy_launch_info.kernel_variant of that build
// bodies / lanes of a group: 6 bodies (15 pairs) on 16-lane groups, two trajectories per warp - the precompiled
// kernels; with HY_NBR_G32 (a build made at hy_create time, hy_jit.hpp) 7 or 8 bodies (21 / 28 pairs) on the 32 lanes
// of a warp, one trajectory per warp.  The host mirrors these numbers in NbrHostLayout (hy_nbody_match.hpp).
#ifdef HY_NBR_G32
constexpr int NBR_MAXB = 8;
constexpr int NBR_NLANES = 32;
#else
constexpr int NBR_MAXB = 6;    // bodies (pairs <= 15 fit one 16-lane group)
constexpr int NBR_NLANES = 16;
#endif

// Trajectory column (shared memory, elements):
//   body b, order k:  [x, y, z, -, vx, vy, vz, -]  at  b * NBR_BS + k * NBR_JS   (16-byte aligned
//                     triples: one 128-bit + one 64-bit access moves a vector)
//   exchange buffer:  2 x 16 pair slots of NBR_TS elements (t0, t1, t2 used) at NBR_TB0
//                     (used only with -DHY_NBR_SMEM_EXCHANGE; the default moves the pair
//                     products by shuffle and leaves it idle)
// NBR_BS = 8 * 23 + 2 and NBR_TS = 6 keep the 128-bit accesses of a quarter-warp conflict-free.
constexpr int NBR_JS = 8;                            // stride between orders of a state variable
constexpr int NBR_BS = NBR_JS * (NBR_LMAX + 1) + 2;  // stride between bodies
constexpr int NBR_TB0 = NBR_MAXB * NBR_BS;           // offset of the exchange buffer
// stride between pair slots of the exchange buffer: 6 (conflict-free 128-bit stores); the
// warpgroup-rotation variant (WGX, 24 trajectories per SM) packs them at 4 to fit shared memory
__host__ __device__ constexpr int nbr_ts(bool wgx) { return wgx ? 4 : 6; }
#ifdef HY_NBR_SMEM_EXCHANGE
__host__ __device__ constexpr int nbr_tbuf(bool wgx) { return 16 * nbr_ts(wgx); }            // one exchange buffer
#else
__host__ __device__ constexpr int nbr_tbuf(bool) { return 0; } // (pair products travel by shuffle: no buffer)
#endif
__host__ __device__ constexpr int nbr_ws(bool wgx) { return NBR_TB0 + 2 * nbr_tbuf(wgx); }    // column length
// register budgets of the warpgroup-rotation variant (3 warpgroups x 168 = 2 x JREG + TREG)
constexpr int NBR_WGX_JREG = 224, NBR_WGX_TREG = 56;
// element offset of state variable i (order 0) in the column
__host__ __device__ constexpr int nbr_state_off(int i) { return (i / 6) * NBR_BS + (i % 6) + ((i % 6) >= 3 ? 1 : 0); }

// Host-built description of a matched N-body tape (hy_nbody_match.hpp).  It travels in
// the program's immediate table (shared memory):
//   imm[body * NBR_CS + q]             coefficient of term q of the body's acceleration sums
//   imm[NBR_LANE0 + s]                 2 x uint32 of lane s: bodies a, b of its pair (d = x[a] - x[b])
//   imm[NBR_OFF0 + body * NBR_CS + q]  uint32: pair slot index feeding term q of the body's sums
constexpr int NBR_CS = 9; // odd stride: the body lanes read their rows conflict-free
constexpr int NBR_LANE0 = NBR_MAXB * NBR_CS;
constexpr int NBR_OFF0 = NBR_LANE0 + NBR_NLANES;
constexpr int NBR_NIMM = NBR_OFF0 + NBR_MAXB * NBR_CS;

// ---- one order of one pair, everything in registers ----
// d*, r2, c: jets (orders 0..PMAX-2 are re-read later); dk*: d[K] (just formed).
// Term order (shared with pair3_k in hy_kernels.cuh): every chain adds the term
// that involves the newest value LAST, so the chains start before it arrives.
template <typename R, int K, int PMAX>
__device__ __forceinline__ void nbr_pair_order(R (&d0)[PMAX], R (&d1)[PMAX], R (&d2)[PMAX], R (&r2)[PMAX],
                                               R (&c)[PMAX], R &inv, const R dk0, const R dk1, const R dk2, R &t0,
                                               R &t1, R &t2)
{
    constexpr int half = (K + 1) / 2;
    // r2[K] = 2 * sum_i sum_{j<half} d_i[j] d_i[K-j]  (+ sum_i d_i[K/2]^2 for even K)
    R q0 = 0, q1 = 0, q2 = 0;
    if constexpr (half > 1) {
        q0 = d0[1] * d0[K - 1];
        q1 = d1[1] * d1[K - 1];
        q2 = d2[1] * d2[K - 1];
#pragma unroll
        for (int j = 2; j < half; ++j) {
            q0 = fma(d0[j], d0[K - j], q0);
            q1 = fma(d1[j], d1[K - j], q1);
            q2 = fma(d2[j], d2[K - j], q2);
        }
    }
    R e = 0;
    if constexpr ((K & 1) == 0 && K > 0) e = fma(d2[K / 2], d2[K / 2], fma(d1[K / 2], d1[K / 2], d0[K / 2] * d0[K / 2]));
    if constexpr (half > 0) {
        q0 = fma(d0[0], dk0, q0);
        q1 = fma(d1[0], dk1, q1);
        q2 = fma(d2[0], dk2, q2);
    }
    R acc = (q0 + q1) + q2;
    acc = acc + acc;
    if constexpr (K == 0) e = fma(dk2, dk2, fma(dk1, dk1, dk0 * dk0));
    if constexpr ((K & 1) == 0) acc += e;
    // w[K] by the power recurrence, alpha = -3/2: weights K*alpha - j*(alpha+1) are immediates
    R ck;
    if constexpr (K == 0) {
        inv = (R)1 / acc;
        ck = (R)1 / (acc * sqrt(acc));
    } else {
        constexpr double alpha = -1.5, al1 = alpha + 1.0, kal = (double)K * alpha;
        R s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
        for (int j = 1; j < K; ++j) {
            const R pr = (R)(kal - (double)j * al1) * r2[K - j];
            if ((j & 3) == 0) s0 = fma(pr, c[j], s0);
            if ((j & 3) == 1) s1 = fma(pr, c[j], s1);
            if ((j & 3) == 2) s2 = fma(pr, c[j], s2);
            if ((j & 3) == 3) s3 = fma(pr, c[j], s3);
        }
        const R tot = fma((R)kal * acc, c[0], (s0 + s1) + (s2 + s3));
        ck = (tot * (R)(1.0 / (double)K)) * inv;
    }
    // t_i[K] = sum_{j<=K} d_i[j] w[K-j]; the j = 0 term (newest w) goes last
    R a0 = 0, a1 = 0, a2 = 0, b0 = 0, b1 = 0, b2 = 0;
    if constexpr (K >= 1) {
        // j = K uses d[K] from the register
        if constexpr (K & 1) {
            b0 = dk0 * c[0];
            b1 = dk1 * c[0];
            b2 = dk2 * c[0];
        } else {
            a0 = dk0 * c[0];
            a1 = dk1 * c[0];
            a2 = dk2 * c[0];
        }
#pragma unroll
        for (int j = K - 1; j >= 1; --j) {
            const R cj = c[K - j];
            if (j & 1) {
                b0 = fma(d0[j], cj, b0);
                b1 = fma(d1[j], cj, b1);
                b2 = fma(d2[j], cj, b2);
            } else {
                a0 = fma(d0[j], cj, a0);
                a1 = fma(d1[j], cj, a1);
                a2 = fma(d2[j], cj, a2);
            }
        }
    }
    if constexpr (K == 0) {
        t0 = dk0 * ck;
        t1 = dk1 * ck;
        t2 = dk2 * ck;
    } else {
        t0 = fma(d0[0], ck, a0 + b0);
        t1 = fma(d1[0], ck, a1 + b1);
        t2 = fma(d2[0], ck, a2 + b2);
    }
    if constexpr (K < PMAX - 1) {
        d0[K] = dk0;
        d1[K] = dk1;
        d2[K] = dk2;
        r2[K] = acc;
        c[K] = ck;
    }
}

// Per-lane constants of the register-resident path: element offsets into the trajectory
// column `w` (shared memory), so that every access is `LDS/STS [base + immediate]`.
template <int NB> struct NbrLane {
    int32_t xa, xb;      // blocks of the pair's bodies
    int32_t ta;          // exchange slot the pair writes (buffer 0)
    int32_t xbody;       // block of this lane's body (body lanes; may point into the
    int32_t tin[NB - 1]; // pair lanes feeding the body's terms  neighbouring column)
                         // (HY_NBR_SMEM_EXCHANGE: their exchange slots)
    int32_t coef;        // offset of the body's coefficient row in the immediate table
    bool body;
};

// 128-bit + 64-bit shared-memory moves of a 3-vector (16-byte aligned).  The body lanes' stores
// are predicated inline PTX, not a branch: their work stays in the same basic block as the pair
// convolutions of the next order, so ptxas interleaves the two.
template <typename R> struct Vec3 {
    R x, y, z;
};
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
// (plain typed accesses: the compiler knows the alignment from the vector type and is free to
// schedule them; __syncwarp orders them like any other memory access)
__device__ __forceinline__ void lds3(const double *p, Vec3<double> &v)
{
    const double2 a = *reinterpret_cast<const double2 *>(p);
    v.x = a.x;
    v.y = a.y;
    v.z = p[2];
}
__device__ __forceinline__ void lds3(const float *p, Vec3<float> &v)
{
    const float2 a = *reinterpret_cast<const float2 *>(p);
    v.x = a.x;
    v.y = a.y;
    v.z = p[2];
}
__device__ __forceinline__ void sts3(double *p, double x, double y, double z)
{
    *reinterpret_cast<double2 *>(p) = make_double2(x, y);
    p[2] = z;
}
__device__ __forceinline__ void sts3(float *p, float x, float y, float z)
{
    *reinterpret_cast<float2 *>(p) = make_float2(x, y);
    p[2] = z;
}
// Predicated vector store (no branch: the code stays in one basic block).
__device__ __forceinline__ void sts3_if(double *p, double x, double y, double z, bool on)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %4, 0;\n\t"
                 "@q st.shared.v2.f64 [%0], {%1, %2};\n\t@q st.shared.f64 [%0+16], %3;\n\t}"
                 :
                 : "r"(smem_u32(p)), "d"(x), "d"(y), "d"(z), "r"((int)on)
                 : "memory");
}
__device__ __forceinline__ void sts3_if(float *p, float x, float y, float z, bool on)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %4, 0;\n\t"
                 "@q st.shared.v2.f32 [%0], {%1, %2};\n\t@q st.shared.f32 [%0+8], %3;\n\t}"
                 :
                 : "r"(smem_u32(p)), "f"(x), "f"(y), "f"(z), "r"((int)on)
                 : "memory");
}

// FULL: the Taylor order equals PMAX (no run-time order checks, one basic block per order).
template <typename R, int NB, int PMAX, bool FULL, bool WGX, int K> struct NbrOrders {
    static __device__ __forceinline__ void run(R *__restrict__ w, const R (&cf)[NB - 1], const NbrLane<NB> &L,
                                               const uint32_t p, R (&d0)[PMAX], R (&d1)[PMAX], R (&d2)[PMAX],
                                               R (&r2)[PMAX], R (&c)[PMAX], R &inv, R dk0, R dk1, R dk2)
    {
        if constexpr (!FULL) {
            if (K >= p) return;
        }
        constexpr int JS = NBR_JS, NQ = NB - 1;
        constexpr int buf = (K & 1) * nbr_tbuf(WGX);
        R t0, t1, t2;
        nbr_pair_order<R, K, PMAX>(d0, d1, d2, r2, c, inv, dk0, dk1, dk2, t0, t1, t2);
#ifdef HY_NBR_SMEM_EXCHANGE
        sts3(&w[L.ta + buf], t0, t1, t2);
#endif
        __syncwarp();
        // d[K+1] = x_a[K+1] - x_b[K+1]  (x[K+1] was written one order ago; rows up to
        // NBR_PMAX exist whatever p is: no run-time guard on K + 1 < p)
        if constexpr (K + 1 < PMAX) {
            Vec3<R> xa, xb;
            lds3(&w[L.xa + (K + 1) * JS], xa);
            lds3(&w[L.xb + (K + 1) * JS], xb);
            dk0 = xa.x - xb.x;
            dk1 = xa.y - xb.y;
            dk2 = xa.z - xb.z;
        }
        {
            // BODY LANES: acceleration of the body at order K - the tape's LINCOMB, term order
            // kept - then v[K+1] = a[K]/(K+1) and x[K+2] = v[K+1]/(K+2).  No branch: every lane runs
            // the loads and the arithmetic (the other lanes mirror a body lane: a zeroing move and a
            // predicate per load would cost more issue slots than the duplicate wavefronts), only
            // the stores are predicated.
            const bool on = L.body;
            R a0 = 0, a1 = 0, a2 = 0;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                Vec3<R> t;
#ifndef HY_NBR_SMEM_EXCHANGE
                // the pair products come straight from the pair lanes' registers (tin = source lane):
                // no exchange buffer, no store -> load round trip (+5 % over the shared-memory exchange)
                t.x = __shfl_sync(0xffffffffu, t0, L.tin[q]);
                t.y = __shfl_sync(0xffffffffu, t1, L.tin[q]);
                t.z = __shfl_sync(0xffffffffu, t2, L.tin[q]);
#else
                lds3(&w[L.tin[q] + buf], t); // (unpredicated: the other lanes mirror a body lane's address)
#endif
                a0 = fma(cf[q], t.x, a0);
                a1 = fma(cf[q], t.y, a1);
                a2 = fma(cf[q], t.z, a2);
            }
            constexpr R rk1 = (R)(1.0 / (double)(K + 1)), rk2 = (R)(1.0 / (double)(K + 2));
            const R v0 = a0 * rk1, v1 = a1 * rk1, v2 = a2 * rk1;
            sts3_if(&w[L.xbody + (K + 1) * JS + 4], v0, v1, v2, on);
            if constexpr (K + 2 <= PMAX) sts3_if(&w[L.xbody + (K + 2) * JS], v0 * rk2, v1 * rk2, v2 * rk2, on);
        }
        if constexpr (K + 1 < PMAX)
            NbrOrders<R, NB, PMAX, FULL, WGX, K + 1>::run(w, cf, L, p, d0, d1, d2, r2, c, inv, dk0, dk1, dk2);
    }
};

// All orders 0..p-1 of one step.  On entry the order-0 rows of the state jets
// hold the state (visible to the whole group); on exit rows 0..p are complete.
template <typename R, int NB, int PMAX, bool FULL, bool WGX = false>
__device__ __forceinline__ void nbr_jets(R *__restrict__ w, const double *__restrict__ coef, const NbrLane<NB> &L,
                                         const uint32_t p, const uint32_t par_off = 0)
{
    constexpr int JS = NBR_JS;
    R d0[PMAX], d1[PMAX], d2[PMAX], r2[PMAX], c[PMAX], inv = 0;
    // coefficients of this lane's body (registers for the whole step)
    R cf[NB - 1];
#pragma unroll
    for (int q = 0; q < NB - 1; ++q) cf[q] = (R)coef[q];
#ifdef HY_NBR_PAR
    // Masses as runtime parameters (a build made at hy_create time, hy_jit.hpp): term q of the body's sums is
    // coef[q] * par[pidx - 1] - the tape's LINCOMB multiplies the two before the FMA, and so does this.  The
    // parameter rows sit at par_off in the column of the trajectory this body lane serves.
    {
        const int col = L.xbody - (L.coef / NBR_CS) * NBR_BS; // 0, or the neighbouring column
#pragma unroll
        for (int q = 0; q < NB - 1; ++q) {
            const uint32_t pi = reinterpret_cast<const uint2 *>(coef + NBR_OFF0 + q)->y;
            if (pi) cf[q] = cf[q] * w[col + (int)par_off + (int)pi - 1];
        }
    }
#else
    (void)par_off;
#endif
    // the body lanes read the other trajectory's state: make the whole warp's updates visible
    __syncwarp();
    // x[1] = v[0]
    {
        Vec3<R> v;
        lds3(&w[L.xbody + 4], v);
        sts3_if(&w[L.xbody + JS], v.x, v.y, v.z, L.body);
    }
    Vec3<R> xa, xb;
    lds3(&w[L.xa], xa);
    lds3(&w[L.xb], xb);
    NbrOrders<R, NB, PMAX, FULL, WGX, 0>::run(w, cf, L, p, d0, d1, d2, r2, c, inv, xa.x - xb.x, xa.y - xb.y, xa.z - xb.z);
    __syncwarp();
}

// ---- warpgroup rotation (WGX): register hand-over between the warpgroups of a CTA ----
// The 190 jet registers are live only during the jets.  Three warpgroups share the register file:
// two hold NBR_WGX_JREG registers (jets), one NBR_WGX_TREG (tail of the step / waiting); a warpgroup
// releases its registers when it leaves the jets and re-acquires them (blocking) before the next ones.
template <int N> __device__ __forceinline__ void wg_reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void wg_reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
// named barrier of one warpgroup (ids 1..3; 0 is __syncthreads)
__device__ __forceinline__ void wg_bar(uint32_t wg) { asm volatile("bar.sync %0, 128;" ::"r"(wg + 1u) : "memory"); }
__device__ __forceinline__ bool wg_any(uint32_t wg, bool v)
{
    uint32_t r;
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\tbar.red.or.pred p, %2, 128, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(r)
                 : "r"((uint32_t)v), "r"(wg + 1u)
                 : "memory");
    return r != 0;
}

} // namespace hy
