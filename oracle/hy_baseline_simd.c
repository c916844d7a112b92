/*
 * hy_baseline_simd.c - SIMD-batched, multithreaded CPU baseline for bench.py.
 *
 * TEST/BENCH INFRASTRUCTURE ONLY (never loaded by the product).  Same algorithm
 * and tape as oracle/hy_oracle.c, restructured the way the reference runs on a
 * CPU: W lanes advance in lock-step inside one thread (the reference's SIMD
 * batch mode, doc/notebooks/Batch mode overview.ipynb; lanes that are done take
 * zero-length steps until the whole batch is done) and batches are spread over
 * the host threads (the reference's thread-pool ensemble,
 * heyoka/_ensemble_impl.py:23-68).  FP64, propagate_until only, no events.
 * Validated against the scalar oracle by tests/test_oracle_golden.py.
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

#define W 8
#define MAXP 64
typedef double vd[W];

#define FORL for (int l = 0; l < W; ++l)
#define ROW(r) (ws + (size_t)(r) * W)
static inline double *rowk(double *ws, uint32_t ref, uint32_t k)
{
    return ws + (size_t)((ref & 0x7fffffffu) + ((ref & HY_REF_JET) ? k : 0)) * W;
}

static void exec_op(const hy_op *o, const hy_term *T, double *ws, const double *pars, const double *rk, const double *tm,
                    uint32_t k)
{
    switch (o->opcode) {
    case HY_OP_LINCOMB: {
        vd acc;
        FORL acc[l] = 0;
        for (uint32_t i = 0; i < o->n; ++i) {
            const hy_term *t = &T[o->b + i];
            if (t->src == HY_REF_ONE) {
                if (k == 0) FORL acc[l] += t->coef * (t->par >= 0 ? pars[(size_t)t->par * W + l] : 1.0);
            } else {
                const double *v = rowk(ws, t->src, k);
                if (t->par >= 0) {
                    const double *p = pars + (size_t)t->par * W;
#pragma omp simd
                    FORL acc[l] = fma(t->coef * p[l], v[l], acc[l]);
                } else {
#pragma omp simd
                    FORL acc[l] = fma(t->coef, v[l], acc[l]);
                }
            }
        }
        double *d = (o->flags & HY_OPF_SVD) ? ROW((o->dst & 0x7fffffffu) + k + 1) : rowk(ws, o->dst, k);
        const double s = (o->flags & HY_OPF_SVD) ? rk[k + 1] : 1.0;
        FORL d[l] = acc[l] * s;
    } break;
    case HY_OP_ADDSUB: {
        const double *a = rowk(ws, o->a, k), *b = rowk(ws, o->b, k);
        const double sa = (o->flags & HY_OPF_NEGA) ? -1.0 : 1.0, sb = (o->flags & HY_OPF_NEGB) ? -1.0 : 1.0;
        double *d = (o->flags & HY_OPF_SVD) ? ROW((o->dst & 0x7fffffffu) + k + 1) : rowk(ws, o->dst, k);
        const double s = (o->flags & HY_OPF_SVD) ? rk[k + 1] : 1.0;
#pragma omp simd
        FORL d[l] = (sa * a[l] + sb * b[l]) * s;
    } break;
    case HY_OP_MUL: {
        const double *a = ROW(o->a & 0x7fffffffu), *b = ROW(o->b & 0x7fffffffu);
        vd acc;
        FORL acc[l] = 0;
        for (uint32_t j = 0; j <= k; ++j) {
            const double *x = a + (size_t)j * W, *y = b + (size_t)(k - j) * W;
#pragma omp simd
            FORL acc[l] = fma(x[l], y[l], acc[l]);
        }
        double *d = rowk(ws, o->dst, k);
        FORL d[l] = acc[l];
    } break;
    case HY_OP_SQUARE:
    case HY_OP_SUMSQ: {
        vd acc, mid;
        FORL acc[l] = mid[l] = 0;
        const uint32_t half = (k + 1) / 2, nt = o->opcode == HY_OP_SQUARE ? 1 : o->n;
        for (uint32_t i = 0; i < nt; ++i) {
            const double *a = ROW((o->opcode == HY_OP_SQUARE ? o->a : T[o->b + i].src) & 0x7fffffffu);
            for (uint32_t j = 0; j < half; ++j) {
                const double *x = a + (size_t)j * W, *y = a + (size_t)(k - j) * W;
#pragma omp simd
                FORL acc[l] = fma(x[l], y[l], acc[l]);
            }
            if ((k & 1u) == 0) {
                const double *m = a + (size_t)(k / 2) * W;
#pragma omp simd
                FORL mid[l] = fma(m[l], m[l], mid[l]);
            }
        }
        double *d = rowk(ws, o->dst, k);
        FORL d[l] = 2 * acc[l] + mid[l];
    } break;
    case HY_OP_MULSH: {
        const double *b = ROW(o->a & 0x7fffffffu);
        for (uint32_t i = 0; i < o->n; ++i) {
            const double *a = ROW(T[o->b + i].src & 0x7fffffffu);
            vd acc;
            FORL acc[l] = 0;
            for (uint32_t j = 0; j <= k; ++j) {
                const double *x = a + (size_t)j * W, *y = b + (size_t)(k - j) * W;
#pragma omp simd
                FORL acc[l] = fma(x[l], y[l], acc[l]);
            }
            double *d = rowk(ws, T[o->b + i].dst, k);
            FORL d[l] = acc[l];
        }
    } break;
    case HY_OP_DIV: {
        const double *b = ROW(o->b & 0x7fffffffu);
        double *c = ROW(o->dst & 0x7fffffffu), *inv = ROW(o->dst2);
        if (k == 0) FORL inv[l] = 1.0 / b[l];
        vd acc;
        const double *a = rowk(ws, o->a, k);
        FORL acc[l] = a[l];
        for (uint32_t j = 1; j <= k; ++j) {
            const double *x = b + (size_t)j * W, *y = c + (size_t)(k - j) * W;
#pragma omp simd
            FORL acc[l] = fma(-x[l], y[l], acc[l]);
        }
        FORL c[(size_t)k * W + l] = acc[l] * inv[l];
    } break;
    case HY_OP_POW:
    case HY_OP_SQRT: {
        const double *a = ROW(o->a & 0x7fffffffu);
        double *c = ROW(o->dst & 0x7fffffffu), *inv = ROW(o->dst2);
        const double alpha = o->opcode == HY_OP_SQRT ? 0.5 : o->imm;
        if (k == 0) {
            FORL
            {
                const double x = a[l];
                inv[l] = 1.0 / x;
                c[l] = alpha == -1.5 ? 1.0 / (x * sqrt(x))
                                     : (alpha == 0.5 ? sqrt(x)
                                                     : (alpha == -0.5 ? 1.0 / sqrt(x)
                                                                      : (alpha == -1.0 ? 1.0 / x : pow(x, alpha))));
            }
        } else {
            vd acc;
            FORL acc[l] = 0;
            for (uint32_t j = 0; j < k; ++j) {
                const double wgt = (double)k * alpha - (double)j * (alpha + 1.0);
                const double *x = a + (size_t)(k - j) * W, *y = c + (size_t)j * W;
#pragma omp simd
                FORL acc[l] = fma(wgt * x[l], y[l], acc[l]);
            }
            FORL c[(size_t)k * W + l] = (acc[l] * rk[k]) * inv[l];
        }
    } break;
    case HY_OP_EXP: {
        const double *a = ROW(o->a & 0x7fffffffu);
        double *c = ROW(o->dst & 0x7fffffffu);
        if (k == 0) {
            FORL c[l] = exp(a[l]);
        } else {
            vd acc;
            FORL acc[l] = 0;
            for (uint32_t j = 1; j <= k; ++j) FORL acc[l] = fma((double)j * a[(size_t)j * W + l], c[(size_t)(k - j) * W + l], acc[l]);
            FORL c[(size_t)k * W + l] = acc[l] * rk[k];
        }
    } break;
    case HY_OP_INTG: {
        const double *a = ROW(o->a & 0x7fffffffu), *b = ROW(o->b & 0x7fffffffu);
        double *c = ROW(o->dst & 0x7fffffffu);
        const int jet = (o->dst & 0x80000000u) != 0;
        if (k == 0) {
            const int code = (int)o->imm;
            FORL c[l] = code == 0 ? asin(a[l]) : (code == 1 ? acos(a[l]) : (code == 2 ? atan(a[l]) : erf(a[l])));
        } else {
            vd acc;
            FORL acc[l] = 0;
            for (uint32_t j = 1; j <= k; ++j) FORL acc[l] = fma((double)j * a[(size_t)j * W + l], b[(size_t)(k - j) * W + l], acc[l]);
            FORL c[(size_t)(jet ? k : 0) * W + l] = acc[l] * rk[k];
        }
    } break;
    case HY_OP_LOG: {
        const double *a = ROW(o->a & 0x7fffffffu);
        double *c = ROW(o->dst & 0x7fffffffu), *inv = ROW(o->dst2);
        if (k == 0) {
            FORL
            {
                inv[l] = 1.0 / a[l];
                c[l] = log(a[l]);
            }
        } else {
            vd acc;
            FORL acc[l] = 0;
            for (uint32_t j = 1; j < k; ++j) FORL acc[l] = fma((double)j * c[(size_t)j * W + l], a[(size_t)(k - j) * W + l], acc[l]);
            FORL c[(size_t)k * W + l] = (a[(size_t)k * W + l] - acc[l] * rk[k]) * inv[l];
        }
    } break;
    case HY_OP_SINCOS: {
        const double *a = ROW(o->a & 0x7fffffffu);
        double *s = ROW(o->dst & 0x7fffffffu), *c = ROW(o->dst2 & 0x7fffffffu);
        if (k == 0) {
            FORL
            {
                s[l] = sin(a[l]);
                c[l] = cos(a[l]);
            }
        } else {
            vd sa, ca;
            FORL sa[l] = ca[l] = 0;
            for (uint32_t j = 1; j <= k; ++j) FORL
                {
                    const double ja = (double)j * a[(size_t)j * W + l];
                    sa[l] = fma(ja, c[(size_t)(k - j) * W + l], sa[l]);
                    ca[l] = fma(ja, s[(size_t)(k - j) * W + l], ca[l]);
                }
            FORL
            {
                s[(size_t)k * W + l] = sa[l] * rk[k];
                c[(size_t)k * W + l] = -(ca[l] * rk[k]);
            }
        }
    } break;
    case HY_OP_TIME: {
        double *c = ROW(o->dst & 0x7fffffffu);
        FORL c[(size_t)k * W + l] = k == 0 ? tm[l] : (k == 1 ? 1.0 : 0.0);
    } break;
    case HY_OP_SVD: {
        const double *a = rowk(ws, o->a, k);
        double *x = ROW((o->dst & 0x7fffffffu) + k + 1);
        FORL x[l] = a[l] * rk[k + 1];
    } break;
    default: break;
    }
}

/* propagate_until for B lanes (FP64, no events); state [n, B] lane-fastest.
 * Returns the total number of accepted steps in *total_steps. */
int ora_simd_propagate_until(const hy_dims *d, const hy_op *ops, const hy_term *terms, uint32_t B, double *state,
                             const double *pars, const double *t0, const double *tf, uint64_t *n_steps, int nthreads)
{
    const uint32_t p = d->order, P1 = p + 1, n = d->n_state;
    if (p > MAXP || p < 2) return 1;
    double rk[MAXP + 2];
    rk[0] = 0;
    for (uint32_t k = 1; k <= p + 1; ++k) rk[k] = 1.0 / (double)k;
    const double rhofac = exp(-7.0 / (10.0 * (p - 1.0))) / (M_E * M_E), inv_p = 1.0 / p, inv_pm1 = 1.0 / (p - 1.0);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    const int64_t nb = ((int64_t)B + W - 1) / W;
#pragma omp parallel
    {
        double *ws = (double *)aligned_alloc(64, ((size_t)d->n_rows + 1) * W * sizeof(double));
        double *lp = (double *)aligned_alloc(64, ((size_t)d->n_par + 1) * W * sizeof(double));
        memset(ws, 0, ((size_t)d->n_rows + 1) * W * sizeof(double));
#pragma omp for schedule(dynamic, 1)
        for (int64_t b = 0; b < nb; ++b) {
            vd hi, lo, tfin, h;
            uint64_t ns[W];
            int done[W], lane[W];
            FORL
            {
                lane[l] = (int)(b * W + l < (int64_t)B ? b * W + l : B - 1); /* pad with a copy of the last lane */
                hi[l] = t0[lane[l]];
                lo[l] = 0;
                tfin[l] = tf[lane[l]];
                ns[l] = 0;
                done[l] = 0;
            }
            for (uint32_t i = 0; i < n; ++i) FORL ws[(size_t)(i * P1) * W + l] = state[(size_t)i * B + lane[l]];
            for (uint32_t i = 0; i < d->n_par; ++i) FORL lp[(size_t)i * W + l] = pars[(size_t)i * B + lane[l]];
            for (;;) {
                int all = 1;
                vd rem;
                FORL
                {
                    rem[l] = (tfin[l] - hi[l]) - lo[l];
                    if (rem[l] == 0) done[l] = 1;
                    all &= done[l];
                }
                if (all) break;
                for (uint32_t k = 0; k < p; ++k)
                    for (uint32_t i = 0; i < d->n_ops; ++i) exec_op(&ops[i], terms, ws, lp, rk, hi, k);
                vd n0, n1, n2;
                FORL n0[l] = n1[l] = n2[l] = 0;
                for (uint32_t i = 0; i < n; ++i) {
                    const double *x = ws + (size_t)(i * P1) * W;
                    FORL
                    {
                        n0[l] = fmax(n0[l], fabs(x[l]));
                        n1[l] = fmax(n1[l], fabs(x[(size_t)(p - 1) * W + l]));
                        n2[l] = fmax(n2[l], fabs(x[(size_t)p * W + l]));
                    }
                }
                FORL
                {
                    const double num = n0[l] < 1 ? 1 : n0[l];
                    const double r1 = pow(num / n2[l], inv_p), r2 = pow(num / n1[l], inv_pm1);
                    double hh = (r1 < r2 ? r1 : r2) * rhofac;
                    if (signbit(rem[l])) hh = -hh;
                    if (fabs(hh) > fabs(rem[l])) hh = rem[l];
                    h[l] = done[l] ? 0.0 : hh;
                }
                for (uint32_t i = 0; i < n; ++i) {
                    double *x = ws + (size_t)(i * P1) * W;
                    vd acc;
                    FORL acc[l] = x[(size_t)p * W + l];
                    for (uint32_t k = p; k-- > 0;) {
#pragma omp simd
                        FORL acc[l] = fma(acc[l], h[l], x[(size_t)k * W + l]);
                    }
                    FORL x[l] = acc[l];
                }
                FORL
                {
                    if (done[l]) continue;
                    ++ns[l];
                    if (h[l] == rem[l]) {
                        hi[l] = tfin[l];
                        lo[l] = 0;
                        done[l] = 1;
                    } else {
                        const double s = hi[l] + h[l], bb = s - hi[l];
                        double err = (hi[l] - (s - bb)) + (h[l] - bb);
                        err += lo[l];
                        const double nh = s + err;
                        lo[l] = err - (nh - s);
                        hi[l] = nh;
                    }
                }
            }
            FORL
            {
                if (b * W + l >= (int64_t)B) continue;
                for (uint32_t i = 0; i < n; ++i) state[(size_t)i * B + lane[l]] = ws[(size_t)(i * P1) * W + l];
                if (n_steps) n_steps[lane[l]] = ns[l];
            }
        }
        free(ws);
        free(lp);
    }
    return 0;
}
