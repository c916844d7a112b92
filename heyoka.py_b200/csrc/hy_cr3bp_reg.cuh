// hy_cr3bp_reg.cuh - register-resident jets for the circular restricted three-body problem.
//
// The tape of the reference's model.cr3bp (/root/reference/heyoka/expose_models.cpp:395-400; equations in
// doc/notebooks/The restricted three-body problem.ipynb:44-62) is 24 ops on 6 state variables
// (hy_cr3bp_match.hpp).  TWO lanes serve one trajectory, one per primary:
//
//   lane 0 (primary A):  X = x + cA   J = y   R = X^2 + y^2 + z^2   C = R^(-3/2)   GG =  gA C1 + gB C2
//   lane 1 (primary B):  X = x + cB   J = z   R = X^2 + y^2 + z^2   C = R^(-3/2)   GG = -(gA C1 + gB C2)
//   products             T = X * (m C)        U = J * GG
//
// The five jets X, J, R, C, GG (orders 0..p-1) stay in registers for the whole step (5 x 20 doubles);
// the scaled jet m C is recomputed on the fly (one multiplication per term) instead of being held.
// The two lanes run the SAME instruction stream (squares, power recurrence, two products) on their
// own data: 16 trajectories per warp step in lockstep, nothing diverges.  Per order they swap
// J^2, C, T and U with one `shfl.bfly` each and both form the six state recurrences; lane 0 stores
// (x, y, z)[k+1], lane 1 (px, py, pz)[k+1] to the trajectory's column in shared memory (layout
// [order][variable]), which the shared tail of the step (norms, step size, Horner, dense output)
// reads like any other state jet.
//
// Arithmetic: the interpreter's (hy_kernels.cuh exec_op) - the same four-chain convolutions,
// symmetric squares, power recurrence, LINCOMB term order and roundings - so the two paths agree
// bit for bit (tests/test_gpu_cr3bp_reg.py).
#pragma once
#include <cstdint>

namespace hy {

constexpr int CRB_XS = 6;       // stride between orders of a state variable in the column
constexpr int CRB_NIMM = 8;     // immediates: cA cB | mA mB | gA gB | nA nB
constexpr int CRB_VARIANT = 203; // hy_launch_info.kernel_variant of this kernel
constexpr int CRB_VARIANT_P22 = 222; // FP64 build unrolled to order 22 (tol = 1e-18, the setting of the
                                     // reference's "restricted three-body problem" notebook): orders 21..22
constexpr int CRB_PMAX_HI = 22;
template <typename R> struct CrbPmax;
template <> struct CrbPmax<double> {
    static constexpr int value = 20; // tol = eps64 (lower orders take the order-checked path)
};
template <> struct CrbPmax<float> {
    static constexpr int value = 9; // tol = eps32 (lower orders take the order-checked path)
};

template <typename R> struct CrbLane {
    R c, m, ga, gb; // this lane's constants
    bool sub;       // 0: primary A (x y z), 1: primary B (px py pz)
    int32_t soff;   // 3 * sub: the three state variables this lane stores
};

__device__ __forceinline__ double crb_fma(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float crb_fma(float a, float b, float c) { return fmaf(a, b, c); }
// first term of a chain: a product that the compiler must not contract into a later addition
__device__ __forceinline__ double crb_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float crb_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double crb_sqrt(double a) { return sqrt(a); }
__device__ __forceinline__ float crb_sqrt(float a) { return sqrtf(a); }
__device__ __forceinline__ double crb_swap(double v)
{
    return __hiloint2double(__shfl_xor_sync(0xffffffffu, __double2hiint(v), 1),
                            __shfl_xor_sync(0xffffffffu, __double2loint(v), 1));
}
__device__ __forceinline__ float crb_swap(float v) { return __shfl_xor_sync(0xffffffffu, v, 1); }

// (s0 + s1) + (s2 + s3) over the chains that exist (absent chains are +0)
template <typename R, int N> __device__ __forceinline__ R crb_fold(const R (&s)[4])
{
    if constexpr (N <= 0)
        return (R)0;
    else if constexpr (N == 1)
        return s[0];
    else if constexpr (N == 2)
        return s[0] + s[1];
    else if constexpr (N == 3)
        return (s[0] + s[1]) + s[2];
    else
        return (s[0] + s[1]) + (s[2] + s[3]);
}

// SQUARE at order K (exec_op HY_OP_SQUARE): 2 * sum_{j < half} a[j] a[K-j]  (+ a[K/2]^2)
template <typename R, int PMAX, int K> __device__ __forceinline__ R crb_square(const R (&a)[PMAX])
{
    constexpr int half = (K + 1) >> 1;
    R s[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < half; ++j) s[j & 3] = j < 4 ? crb_mul(a[j], a[K - j]) : crb_fma(a[j], a[K - j], s[j & 3]);
    R acc = crb_fold<R, half>(s);
    acc = acc + acc;
    if constexpr ((K & 1) == 0) acc = crb_fma(a[K >> 1], a[K >> 1], acc);
    return acc;
}

// FULL: the Taylor order equals PMAX (no run-time order checks).
template <typename R, int PMAX, bool FULL, int K> struct CrbOrders {
    static __device__ __forceinline__ void run(R *__restrict__ w, const CrbLane<R> &L, const uint32_t p, R (&X)[PMAX],
                                               R (&J)[PMAX], R (&Q)[PMAX], R (&C)[PMAX], R (&GG)[PMAX], R &inv,
                                               const R xk, const R yk, const R zk, const R pxk, const R pyk,
                                               const R pzk)
    {
        if constexpr (!FULL) {
            if (K >= p) return; // (warp-uniform)
        }
        const bool sub = L.sub;
        // ---- this order of the lane's coordinates (LINCOMB x + c: the constant enters at order 0 only) ----
        if constexpr (K == 0)
            X[K] = xk + L.c;
        else
            X[K] = xk;
        J[K] = sub ? zk : yk;
        // ---- squares, r^2 = (X^2 + y^2) + z^2 ----
        const R sqx = crb_square<R, PMAX, K>(X), sqj = crb_square<R, PMAX, K>(J);
        const R sqo = crb_swap(sqj);
        const R r2 = (sqx + (sub ? sqo : sqj)) + (sub ? sqj : sqo);
        Q[K] = r2;
        // ---- C = r2^(-3/2) (exec_op HY_OP_POW, alpha = -3/2) ----
        R ck;
        if constexpr (K == 0) {
            inv = (R)1 / r2;
            ck = (R)1 / (r2 * crb_sqrt(r2));
        } else {
            R s[4] = {0, 0, 0, 0};
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const R wj = (R)(-1.5 * (double)K + 0.5 * (double)j); // K alpha - j (alpha + 1)
                const R pr = crb_mul(wj, Q[K - j]);
                s[j & 3] = j < 4 ? crb_mul(pr, C[j]) : crb_fma(pr, C[j], s[j & 3]);
            }
            ck = (crb_fold<R, K>(s) * (R)(1.0 / (double)K)) * inv;
        }
        C[K] = ck;
        // ---- GG = ga C1 + gb C2 (LINCOMB, term order) ----
        const R co = crb_swap(ck);
        GG[K] = crb_fma(L.gb, sub ? ck : co, crb_mul(L.ga, sub ? co : ck));
        // ---- products: T = X * (m C), U = J * GG (exec_op HY_OP_MUL: sum_j a[j] b[K-j]) ----
        R st[4] = {0, 0, 0, 0}, su[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j <= K; ++j) {
            const R g = crb_mul(L.m, C[K - j]);
            st[j & 3] = j < 4 ? crb_mul(X[j], g) : crb_fma(X[j], g, st[j & 3]);
            su[j & 3] = j < 4 ? crb_mul(J[j], GG[K - j]) : crb_fma(J[j], GG[K - j], su[j & 3]);
        }
        const R t = crb_fold<R, K + 1>(st), u = crb_fold<R, K + 1>(su);
        const R to = crb_swap(t), uo = crb_swap(u);
        const R t1 = sub ? to : t, t2 = sub ? t : to, t3 = sub ? uo : u, t4 = sub ? u : uo;
        // ---- state recurrences (both lanes form all six) ----
        constexpr R rk1 = (R)(1.0 / (double)(K + 1));
        const R xn = (pxk + yk) * rk1;
        const R yn = (pyk - xk) * rk1;
        const R zn = pzk * rk1;
        const R pxn = ((pyk - t1) - t2) * rk1;
        const R pyn = (-pxk - t3) * rk1;
        const R pzn = t4 * rk1;
        R *o = w + L.soff + (K + 1) * CRB_XS;
        o[0] = sub ? pxn : xn;
        o[1] = sub ? pyn : yn;
        o[2] = sub ? pzn : zn;
        if constexpr (K + 1 < PMAX)
            CrbOrders<R, PMAX, FULL, K + 1>::run(w, L, p, X, J, Q, C, GG, inv, xn, yn, zn, pxn, pyn, pzn);
    }
};

// All orders 0..p-1 of one step (p <= PMAX).  On entry row 0 of the column holds the state; on exit
// rows 0..p are complete (visible to both lanes after the closing __syncwarp).
template <typename R, int PMAX, bool FULL>
__device__ __forceinline__ void crb_jets(R *__restrict__ w, const CrbLane<R> &L, const uint32_t p)
{
    R X[PMAX], J[PMAX], Q[PMAX], C[PMAX], GG[PMAX], inv = 0;
    __syncwarp();
    const R x0 = w[0], y0 = w[1], z0 = w[2], px0 = w[3], py0 = w[4], pz0 = w[5];
    CrbOrders<R, PMAX, FULL, 0>::run(w, L, p, X, J, Q, C, GG, inv, x0, y0, z0, px0, py0, pz0);
    __syncwarp();
}

} // namespace hy
