// hy_events.cuh - per-trajectory event detection on the device.
//
// For every event function g the order sweep leaves the Taylor polynomial
// g(tau) = sum_k g[k] tau^k of the current step in the workspace.  Detection
// (reference semantics: /root/reference/heyoka/taylor_expose_events.cpp:185-317,
// doc/notebooks/Event detection.ipynb; algorithm SURVEY.md A.9, arXiv:2204.09948):
//   1. rescale to q(s) = g(h s), s in [0, 1);
//   2. fast exclusion: interval-Horner enclosure of q over [0, 1]; if it does
//      not contain 0 there is no event in this step (the common case: O(p));
//   3. otherwise isolate the real roots with Descartes' rule of signs on the
//      reversed + translated polynomial, bisecting (Taylor shift by 1/2) until
//      every interval holds 0 or 1 sign change; exact roots at s = 0 and at the
//      bisection points are reported;
//   4. refine each isolated root by safeguarded Newton/bisection on q;
//   5. direction filter from the sign of dq/dtau at the root; terminal events
//      skip roots inside their cooldown window; the earliest terminal root
//      truncates the step; non-terminal roots before it are logged.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hy_cuda.h"

namespace hy {

constexpr int EV_MAXP1 = 32; // events need order <= 31
constexpr int EV_STACK = 14; // bisection depth
constexpr int EV_MAXROOTS = 8;

template <typename R> struct EvParams {
    const int32_t *dir;       // [n_events] -1, 0, +1
    const double *cooldown;   // [n_tevents] user cooldowns (< 0: automatic)
    R *cd_elapsed, *cd_total; // [B][n_tevents]; total < 0: not in cooldown
    hy_event_rec *log;
    unsigned long long *log_count;
    unsigned long long log_cap;
    R tol;
};

template <typename R> __device__ __forceinline__ R ev_horner(const R *c, int p, R s)
{
    R acc = c[p];
    for (int k = p - 1; k >= 0; --k) acc = fma(acc, s, c[k]);
    return acc;
}
template <typename R> __device__ __forceinline__ void ev_horner_d(const R *c, int p, R s, R &f, R &df)
{
    f = c[p];
    df = 0;
    for (int k = p - 1; k >= 0; --k) {
        df = fma(df, s, f);
        f = fma(f, s, c[k]);
    }
}

// Number of sign changes of (1+x)^p c(1/(1+x)): an upper bound (exact when 0
// or 1) on the number of roots of c in (0, 1).  `tmp` is scratch of p+1.
template <typename R> __device__ inline int ev_descartes(const R *c, int p, R *tmp)
{
    for (int k = 0; k <= p; ++k) tmp[k] = c[p - k]; // reverse
    for (int i = 0; i < p; ++i)                      // Taylor shift by 1
        for (int j = p - 1; j >= i; --j) tmp[j] += tmp[j + 1];
    int v = 0, last = 0;
    for (int k = 0; k <= p; ++k) {
        const int s = tmp[k] > (R)0 ? 1 : (tmp[k] < (R)0 ? -1 : 0);
        if (s != 0) {
            if (last != 0 && s != last) ++v;
            last = s;
        }
    }
    return v;
}

// Roots of q(s) = sum c[k] s^k in [0, 1).  Returns the number found (sorted by
// discovery, not by value); roots[] in s units.
template <typename R> __device__ inline int ev_find_roots(const R *c_in, int p, R *roots)
{
    R stk[EV_STACK][EV_MAXP1];
    R slo[EV_STACK], shi[EV_STACK];
    R tmp[EV_MAXP1];
    int nroots = 0;
    // Strip exact roots at s = 0 (a root at tau = 0 IS an event: Event detection.ipynb).
    int shift = 0;
    while (shift <= p && c_in[shift] == (R)0) ++shift;
    if (shift > p) return 0; // identically zero: no isolated roots
    if (shift > 0) roots[nroots++] = 0;
    const int pp = p - shift;
    if (pp == 0) return nroots;
    int sp = 0;
    for (int k = 0; k <= pp; ++k) stk[0][k] = c_in[k + shift];
    slo[0] = 0;
    shi[0] = 1;
    sp = 1;
    while (sp > 0 && nroots < EV_MAXROOTS) {
        --sp;
        R *c = stk[sp];
        const R lo = slo[sp], hi = shi[sp];
        const int v = ev_descartes<R>(c, pp, tmp);
        if (v == 0) continue;
        if (v == 1 || sp + 2 > EV_STACK || (hi - lo) < (R)1e-9) {
            // One root in (lo, hi) (or we cannot split further): refine on the
            // ORIGINAL polynomial.  f(lo) and f(hi) normally differ in sign.
            const R *q = c_in + shift;
            R a = lo, b = hi;
            // sign of q just right of lo: first non-zero coefficient of the interval polynomial
            int sa = 0;
            for (int k = 0; k <= pp && sa == 0; ++k) sa = c[k] > (R)0 ? 1 : (c[k] < (R)0 ? -1 : 0);
            R fb = ev_horner<R>(q, pp, b);
            if (fb == (R)0) {
                b = hi - (hi - lo) * (R)1e-6;
                fb = ev_horner<R>(q, pp, b);
            }
            if (sa == 0 || fb == (R)0 || (sa > 0) == (fb > (R)0)) continue; // no sign change: nothing to refine
            R x = (R)0.5 * (a + b);
            for (int it = 0; it < 100; ++it) {
                R f, df;
                ev_horner_d<R>(q, pp, x, f, df);
                if (f == (R)0) break;
                if ((f > (R)0) == (sa > 0))
                    a = x;
                else
                    b = x;
                R xn = x - f / df;
                if (!(xn > a && xn < b)) xn = (R)0.5 * (a + b);
                if (xn == x || (b - a) <= (R)0) break;
                const R dx = xn - x;
                x = xn;
                const R ax = x < (R)1e-30 ? (R)1e-30 : x;
                if ((dx < 0 ? -dx : dx) <= (R)2 * (sizeof(R) == 8 ? (R)1.1e-16 : (R)6e-8) * ax) break;
            }
            if (x < (R)1) roots[nroots++] = x;
            continue;
        }
        // Split at the midpoint: left(s) = c(s/2), right(s) = c((s+1)/2) = left(s+1).
        const R mid = (R)0.5 * (lo + hi);
        R *L = stk[sp]; // reuse the popped slot for the left half
        R sc = 1;
        for (int k = 0; k <= pp; ++k) {
            L[k] = c[k] * sc;
            sc *= (R)0.5;
        }
        R *Rr = stk[sp + 1];
        for (int k = 0; k <= pp; ++k) Rr[k] = L[k];
        for (int i = 0; i < pp; ++i)
            for (int j = pp - 1; j >= i; --j) Rr[j] += Rr[j + 1];
        // Exact root at the midpoint: right half has a zero constant term.
        if (Rr[0] == (R)0) {
            if (nroots < EV_MAXROOTS && mid < (R)1) roots[nroots++] = mid;
            // deflate the right half by s
            for (int k = 0; k < pp; ++k) Rr[k] = Rr[k + 1];
            Rr[pp] = 0;
        }
        // process the left half first (earlier roots first): push right, then left
        // (left currently sits at stk[sp]; swap so that left is on top)
        for (int k = 0; k <= pp; ++k) {
            const R t = L[k];
            L[k] = Rr[k];
            Rr[k] = t;
        }
        slo[sp] = mid;
        shi[sp] = hi;
        slo[sp + 1] = lo;
        shi[sp + 1] = mid;
        sp += 2;
    }
    return nroots;
}

// Detect the events of one step of one trajectory (run by ONE lane).
//   w        : trajectory workspace column
//   ev_ref   : jet references of the event functions (terminal first)
//   h        : the step (already clamped); may be negative
// Returns the (possibly truncated) step in h_out and the index of the terminal
// event that truncated it (-1: none).  Non-terminal events before the
// truncation point and the terminal event itself are appended to the log.
// GS: element stride between the orders of an event jet (1 for a trajectory column).
template <typename R, int GS = 1>
static __device__ __noinline__ void detect_events(const R *w, const uint32_t *ev_ref, uint32_t n_events, uint32_t n_tevents,
                                     int p, R h, R t_hi, R t_lo, uint32_t traj, unsigned long long step_idx,
                                     const EvParams<R> E, R &h_out, int &term_out, int &nt_out)
{
    h_out = h;
    term_out = -1;
    nt_out = 0;
    if (h == (R)0 || !(h == h)) return;
    R q[EV_MAXP1];
    R roots[EV_MAXROOTS];
    // candidate list (few entries): event, tau, d_sgn
    int cand_ev[2 * EV_MAXROOTS];
    R cand_tau[2 * EV_MAXROOTS];
    int cand_sg[2 * EV_MAXROOTS];
    int nc = 0;
    R best_tau_abs = (h < 0 ? -h : h) * (R)2;
    int best_ev = -1, best_sg = 0;
    R best_tau = 0;
    for (uint32_t e = 0; e < n_events; ++e) {
        const R *g = w + (ev_ref[e] & 0x7fffffffu);
        // Fast exclusion, straight from the jet in shared memory (no per-thread array: with the
        // shared-memory carve-out at its maximum, local memory lives in L2): interval Horner of
        // g over tau in [0, h] (or [h, 0]).
        {
            R lo = g[p * GS], hi = g[p * GS];
            for (int k = p - 1; k >= 0; --k) {
                const R a = lo * h, b = hi * h; // (h < 0 swaps the ends)
                const R mn = a < b ? a : b, mx = a < b ? b : a;
                const R gk = g[k * GS];
                lo = (mn < 0 ? mn : (R)0) + gk;
                hi = (mx > 0 ? mx : (R)0) + gk;
            }
            if (lo > (R)0 || hi < (R)0) continue;
        }
        // q(s) = g(h s)
        R hk = 1;
        for (int k = 0; k <= p; ++k) {
            q[k] = g[k * GS] * hk;
            hk *= h;
        }
        // second exclusion on the scaled polynomial: enclosure of q over [0, 1]
        R lo = q[p], hi = q[p];
        for (int k = p - 1; k >= 0; --k) {
            lo = (lo < 0 ? lo : (R)0) + q[k];
            hi = (hi > 0 ? hi : (R)0) + q[k];
        }
        if (lo > (R)0 || hi < (R)0) continue;
        const int nr = ev_find_roots<R>(q, p, roots);
        for (int r = 0; r < nr; ++r) {
            const R s = roots[r];
            R f, df;
            ev_horner_d<R>(q, p, s, f, df);
            // sign of dg/dtau = sign(dq/ds) * sign(h)
            int sg = df > 0 ? 1 : (df < 0 ? -1 : 0);
            if (h < 0) sg = -sg;
            const int dir = E.dir[e];
            if (dir != 0 && dir != sg) continue;
            const R tau = s * h;
            const R atau = tau < 0 ? -tau : tau;
            if (e < n_tevents) {
                const size_t ci = (size_t)traj * n_tevents + e;
                const R tot = E.cd_total[ci];
                if (tot >= (R)0 && atau < tot - E.cd_elapsed[ci]) continue; // inside the cooldown
                if (atau < best_tau_abs || (atau == best_tau_abs && (int)e < best_ev)) {
                    best_tau_abs = atau;
                    best_tau = tau;
                    best_ev = (int)e;
                    best_sg = sg;
                }
            } else if (nc < 2 * EV_MAXROOTS) {
                cand_ev[nc] = (int)e;
                cand_tau[nc] = tau;
                cand_sg[nc] = sg;
                ++nc;
            }
        }
    }
    auto push = [&](int e, R tau, int sg) {
        const unsigned long long idx = atomicAdd(E.log_count, 1ULL);
        if (idx < E.log_cap) {
            hy_event_rec r;
            r.lane = traj;
            r.ev_idx = (uint32_t)e;
            r.d_sgn = sg;
            r.step = (uint32_t)step_idx;
            r.t = (double)((t_hi + tau) + t_lo);
            E.log[idx] = r;
        }
    };
    for (int i = 0; i < nc; ++i) {
        const R at = cand_tau[i] < 0 ? -cand_tau[i] : cand_tau[i];
        if (best_ev < 0 || at < best_tau_abs) {
            push(cand_ev[i], cand_tau[i], cand_sg[i]);
            ++nt_out;
        }
    }
    if (best_ev >= 0) {
        push(best_ev, best_tau, best_sg);
        h_out = best_tau;
        term_out = best_ev;
        // start the cooldown of the event that fired
        const R *g = w + (ev_ref[best_ev] & 0x7fffffffu);
        R f = g[p * GS], df = 0, gm = 0, hk = 1;
        for (int k = p - 1; k >= 0; --k) {
            df = fma(df, best_tau, f);
            f = fma(f, best_tau, g[k * GS]);
        }
        for (int k = 0; k <= p; ++k) {
            const R a = g[k * GS] * hk;
            gm = (a < 0 ? -a : a) > gm ? (a < 0 ? -a : a) : gm;
            hk *= h;
        }
        const double user = E.cooldown[best_ev];
        R cd;
        if (user >= 0) {
            cd = (R)user;
        } else {
            // automatic: 10 * g_eps / |dg/dtau|  (heuristic, SURVEY.md A.9 "not verified")
            const R adf = df < 0 ? -df : df;
            const R g_eps = E.tol * (gm > (R)1 ? gm : (R)1);
            cd = adf > (R)0 ? (R)10 * g_eps / adf : (R)0;
        }
        const size_t ci = (size_t)traj * n_tevents + best_ev;
        E.cd_total[ci] = cd;
        E.cd_elapsed[ci] = -(best_tau < 0 ? -best_tau : best_tau); // the advance below adds |h_out| back
    }
}

// Advance the cooldown clocks of a trajectory by |h| (run by one lane).  Returns whether any cooldown
// is still running (a caller that tracks this skips the call - three global loads per step - until
// the next terminal event starts a cooldown).
template <typename R>
__device__ inline bool advance_cooldowns(uint32_t traj, uint32_t n_tevents, R h, const EvParams<R> E)
{
    const R ah = h < 0 ? -h : h;
    bool live = false;
    for (uint32_t e = 0; e < n_tevents; ++e) {
        const size_t ci = (size_t)traj * n_tevents + e;
        const R tot = E.cd_total[ci];
        if (tot >= (R)0) {
            const R el = E.cd_elapsed[ci] + ah;
            if (el >= tot) {
                E.cd_total[ci] = (R)-1;
                E.cd_elapsed[ci] = 0;
            } else {
                E.cd_elapsed[ci] = el;
                live = true;
            }
        }
    }
    return live;
}

} // namespace hy
