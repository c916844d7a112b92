/*
 * hy_oracle_impl.h - body of the CPU oracle, included once per precision with
 * REAL / SUF / math-function macros defined by hy_oracle.c.
 *
 * TEST INFRASTRUCTURE ONLY (see hy_oracle.c).
 */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

typedef struct {
    const hy_dims *d;
    const hy_op *ops;
    const hy_term *terms;
    const uint32_t *ev_ref;
    REAL rk[HY_MAX_ORDER + 2]; /* rk[k] = 1/k */
    REAL rhofac;
    REAL inv_p, inv_pm1;
    int high_accuracy;
} FN(ora_sys);

static inline REAL FN(ld)(const REAL *ws, uint32_t ref, uint32_t k)
{
    return (ref & HY_REF_JET) ? ws[(ref & ~HY_REF_JET) + k] : ws[ref];
}
static inline REAL *FN(st)(REAL *ws, uint32_t ref, uint32_t k)
{
    return (ref & HY_REF_JET) ? &ws[(ref & ~HY_REF_JET) + k] : &ws[ref];
}
/* History access: only legal on jets. */
#define JET(ref) (ws + ((ref) & ~HY_REF_JET))

static REAL FN(pow0)(REAL x, double alpha)
{
    /* Order-0 value of x^alpha; half-integer exponents by sqrt/div
     * (nbody uses alpha = -3/2 on r^2). */
    if (alpha == -1.5) return (REAL)1 / (x * R_SQRT(x));
    if (alpha == -0.5) return (REAL)1 / R_SQRT(x);
    if (alpha == 1.5) return x * R_SQRT(x);
    if (alpha == -1.0) return (REAL)1 / x;
    if (alpha == -2.0) return (REAL)1 / (x * x);
    return R_POW(x, (REAL)alpha);
}

/* One op at order k.  Follows SURVEY.md A.3 (jet recurrences). */
static void FN(exec_op)(const FN(ora_sys) * S, const hy_op *o, REAL *ws, const REAL *pars, REAL tm,
                        uint32_t k)
{
    const hy_term *T = S->terms;
    switch (o->opcode) {
    case HY_OP_LINCOMB: {
        REAL acc = 0;
        for (uint32_t i = 0; i < o->n; ++i) {
            const hy_term *t = &T[o->b + i];
            REAL c = (REAL)t->coef;
            if (t->par >= 0) c = c * pars[t->par];
            REAL v = (t->src == HY_REF_ONE) ? (k == 0 ? (REAL)1 : (REAL)0) : FN(ld)(ws, t->src, k);
            acc = R_FMA(c, v, acc);
        }
        if (o->flags & HY_OPF_SVD)
            JET(o->dst)[k + 1] = acc * S->rk[k + 1]; /* fused x[k+1] = f[k]/(k+1) */
        else
            *FN(st)(ws, o->dst, k) = acc;
    } break;
    case HY_OP_ADDSUB: {
        REAL a = FN(ld)(ws, o->a, k), b = FN(ld)(ws, o->b, k);
        if (o->flags & HY_OPF_NEGA) a = -a;
        if (o->flags & HY_OPF_NEGB) b = -b;
        REAL acc = a + b;
        if (o->flags & HY_OPF_SVD)
            JET(o->dst)[k + 1] = acc * S->rk[k + 1];
        else
            *FN(st)(ws, o->dst, k) = acc;
    } break;
    case HY_OP_MUL: {
        const REAL *a = JET(o->a), *b = JET(o->b);
        REAL acc = 0;
        for (uint32_t j = 0; j <= k; ++j) acc = R_FMA(a[j], b[k - j], acc);
        *FN(st)(ws, o->dst, k) = acc;
    } break;
    case HY_OP_SQUARE: {
        const REAL *a = JET(o->a);
        REAL acc = 0;
        uint32_t half = (k + 1) / 2; /* j < half pairs with k-j > j */
        for (uint32_t j = 0; j < half; ++j) acc = R_FMA(a[j], a[k - j], acc);
        acc = acc + acc;
        if ((k & 1u) == 0) acc = R_FMA(a[k / 2], a[k / 2], acc);
        *FN(st)(ws, o->dst, k) = acc;
    } break;
    case HY_OP_SUMSQ: {
        REAL acc = 0;
        uint32_t half = (k + 1) / 2;
        for (uint32_t i = 0; i < o->n; ++i) {
            const REAL *a = JET(T[o->b + i].src);
            for (uint32_t j = 0; j < half; ++j) acc = R_FMA(a[j], a[k - j], acc);
        }
        acc = acc + acc;
        if ((k & 1u) == 0)
            for (uint32_t i = 0; i < o->n; ++i) {
                const REAL *a = JET(T[o->b + i].src);
                acc = R_FMA(a[k / 2], a[k / 2], acc);
            }
        *FN(st)(ws, o->dst, k) = acc;
    } break;
    case HY_OP_MULSH: {
        const REAL *b = JET(o->a);
        for (uint32_t i = 0; i < o->n; ++i) {
            const REAL *a = JET(T[o->b + i].src);
            REAL acc = 0;
            for (uint32_t j = 0; j <= k; ++j) acc = R_FMA(a[j], b[k - j], acc);
            *FN(st)(ws, T[o->b + i].dst, k) = acc;
        }
    } break;
    case HY_OP_DIV: {
        const REAL *b = JET(o->b);
        REAL *c = JET(o->dst);
        if (k == 0) ws[o->dst2] = (REAL)1 / b[0];
        REAL acc = FN(ld)(ws, o->a, k);
        for (uint32_t j = 1; j <= k; ++j) acc = R_FMA(-b[j], c[k - j], acc);
        c[k] = acc * ws[o->dst2];
    } break;
    case HY_OP_POW:
    case HY_OP_SQRT: {
        const REAL *a = JET(o->a);
        REAL *c = JET(o->dst);
        double alpha = o->opcode == HY_OP_SQRT ? 0.5 : o->imm;
        if (k == 0) {
            ws[o->dst2] = (REAL)1 / a[0];
            c[0] = o->opcode == HY_OP_SQRT ? R_SQRT(a[0]) : FN(pow0)(a[0], alpha);
        } else {
            REAL al = (REAL)alpha, al1 = (REAL)(alpha + 1.0), kal = (REAL)k * al;
            REAL acc = 0;
            for (uint32_t j = 0; j < k; ++j) {
                REAL w = R_FMA(-(REAL)j, al1, kal);
                acc = R_FMA(w * a[k - j], c[j], acc);
            }
            c[k] = (acc * S->rk[k]) * ws[o->dst2];
        }
    } break;
    case HY_OP_EXP: {
        const REAL *a = JET(o->a);
        REAL *c = JET(o->dst);
        if (k == 0) {
            c[0] = R_EXP(a[0]);
        } else {
            REAL acc = 0;
            for (uint32_t j = 1; j <= k; ++j) acc = R_FMA((REAL)j * a[j], c[k - j], acc);
            c[k] = acc * S->rk[k];
        }
    } break;
    case HY_OP_INTG: {
        /* dst = F(a), dF/da = b (include/hy_cuda.h): dst[k] = (1/k) sum_{j=1..k} j a[j] b[k-j] */
        const REAL *a = JET(o->a), *b = JET(o->b);
        if (k == 0) {
            const int code = (int)o->imm;
            *FN(st)(ws, o->dst, 0) = code == 0 ? R_ASIN(a[0]) : (code == 1 ? R_ACOS(a[0]) : (code == 2 ? R_ATAN(a[0]) : R_ERF(a[0])));
        } else {
            REAL acc = 0;
            for (uint32_t j = 1; j <= k; ++j) acc = R_FMA((REAL)j * a[j], b[k - j], acc);
            *FN(st)(ws, o->dst, k) = acc * S->rk[k];
        }
    } break;
    case HY_OP_LOG: {
        const REAL *a = JET(o->a);
        REAL *c = JET(o->dst);
        if (k == 0) {
            ws[o->dst2] = (REAL)1 / a[0];
            c[0] = R_LOG(a[0]);
        } else {
            REAL acc = 0;
            for (uint32_t j = 1; j < k; ++j) acc = R_FMA((REAL)j * c[j], a[k - j], acc);
            c[k] = R_FMA(-acc, S->rk[k], a[k]) * ws[o->dst2];
        }
    } break;
    case HY_OP_SINCOS: {
        const REAL *a = JET(o->a);
        REAL *s = JET(o->dst), *c = JET(o->dst2);
        if (k == 0) {
            s[0] = R_SIN(a[0]);
            c[0] = R_COS(a[0]);
        } else {
            REAL sa = 0, ca = 0;
            for (uint32_t j = 1; j <= k; ++j) {
                REAL ja = (REAL)j * a[j];
                sa = R_FMA(ja, c[k - j], sa);
                ca = R_FMA(ja, s[k - j], ca);
            }
            s[k] = sa * S->rk[k];
            c[k] = -(ca * S->rk[k]);
        }
    } break;
    case HY_OP_TIME: {
        REAL *c = JET(o->dst);
        c[k] = k == 0 ? tm : (k == 1 ? (REAL)1 : (REAL)0);
    } break;
    case HY_OP_SVD: {
        REAL *x = JET(o->dst);
        x[k + 1] = FN(ld)(ws, o->a, k) * S->rk[k + 1];
    } break;
    default: break;
    }
}

/* Build the jets of one lane: state in ws rows i*(p+1), orders 0..p. */
static void FN(build_jets)(const FN(ora_sys) * S, REAL *ws, const REAL *pars, REAL tm)
{
    const hy_dims *d = S->d;
    for (uint32_t k = 0; k < d->order; ++k)
        for (uint32_t i = 0; i < d->n_ops; ++i) FN(exec_op)(S, &S->ops[i], ws, pars, tm, k);
    if (d->n_events) {
        /* Order p of the event functions (and what they depend on). */
        for (uint32_t i = 0; i < d->n_ops; ++i)
            if ((S->ops[i].flags & HY_OPF_EVENT) && S->ops[i].opcode != HY_OP_SVD &&
                !(S->ops[i].flags & HY_OPF_SVD))
                FN(exec_op)(S, &S->ops[i], ws, pars, tm, d->order);
    }
}

/* Jorba-Zou step size; event rows take part in the norms. */
static REAL FN(step_size)(const FN(ora_sys) * S, const REAL *ws)
{
    const hy_dims *d = S->d;
    uint32_t p = d->order, P1 = p + 1;
    REAL n0 = 0, npm1 = 0, np_ = 0;
    int nan_seen = 0;
    for (uint32_t i = 0; i < d->n_state + d->n_events; ++i) {
        const REAL *x = i < d->n_state ? ws + i * P1 : JET(S->ev_ref[i - d->n_state]);
        REAL a0 = R_ABS(x[0]), a1 = R_ABS(x[p - 1]), a2 = R_ABS(x[p]);
        if (a0 != a0 || a1 != a1 || a2 != a2) nan_seen = 1;
        if (a0 > n0) n0 = a0;
        if (a1 > npm1) npm1 = a1;
        if (a2 > np_) np_ = a2;
    }
    if (nan_seen) return (REAL)NAN;
    REAL num = n0 < (REAL)1 ? (REAL)1 : n0;
    REAL rho_p = R_POW(num / np_, S->inv_p);
    REAL rho_pm1 = R_POW(num / npm1, S->inv_pm1);
    REAL rho = rho_p < rho_pm1 ? rho_p : rho_pm1;
    return rho * S->rhofac;
}

static void FN(time_add)(REAL *hi, REAL *lo, REAL h)
{
    /* Error-free addition + renormalisation. */
    volatile REAL s = *hi + h;
    volatile REAL bb = s - *hi;
    volatile REAL err = (*hi - (s - bb)) + (h - bb);
    err = err + *lo;
    volatile REAL nh = s + err;
    volatile REAL nl = err - (nh - s);
    *hi = nh;
    *lo = nl;
}

static REAL FN(time_sub)(REAL ahi, REAL alo, REAL bhi, REAL blo)
{
    volatile REAL s = ahi - bhi;
    volatile REAL bb = s - ahi;
    volatile REAL err = (ahi - (s - bb)) + (-bhi - bb);
    err = err + (alo - blo);
    return s + err;
}

/* One adaptive step of one lane.  `lim` is the signed step limit (its sign
 * selects the direction).  Returns the outcome; *h_out the step taken. */
static int64_t FN(lane_step)(const FN(ora_sys) * S, REAL *ws, const REAL *pars, REAL *t_hi,
                             REAL *t_lo, REAL lim, REAL *h_out, REAL *tc, size_t B, size_t l)
{
    const hy_dims *d = S->d;
    uint32_t p = d->order, P1 = p + 1;
    FN(build_jets)(S, ws, pars, *t_hi);
    REAL h = FN(step_size)(S, ws);
    if (signbit(lim)) h = -h;
    int64_t outcome = HY_OUTCOME_SUCCESS;
    if (R_ABS(h) > R_ABS(lim)) {
        h = lim;
        outcome = HY_OUTCOME_TIME_LIMIT;
    }
    int finite = 1;
    for (uint32_t i = 0; i < d->n_state; ++i) {
        REAL *x = ws + i * P1;
        REAL acc;
        if (tc)
            for (uint32_t k = 0; k <= p; ++k) tc[((size_t)i * P1 + k) * B + l] = x[k];
        if (!S->high_accuracy) {
            acc = x[p];
            for (uint32_t k = p; k-- > 0;) acc = R_FMA(acc, h, x[k]);
        } else {
            /* Compensated summation of the terms x[k] h^k. */
            REAL sum = x[0], comp = 0, hk = h;
            for (uint32_t k = 1; k <= p; ++k) {
                volatile REAL term = x[k] * hk;
                volatile REAL y = term - comp;
                volatile REAL tt = sum + y;
                comp = (tt - sum) - y;
                sum = tt;
                hk = hk * h;
            }
            acc = sum;
        }
        if (!isfinite(acc)) finite = 0;
        x[0] = acc; /* new state; x[1..p] keep the jet of the step (tc) */
    }
    FN(time_add)(t_hi, t_lo, h);
    *h_out = h;
    if (!finite) outcome = HY_OUTCOME_ERR_NF_STATE;
    return outcome;
}

/*
 * Batch propagate (reference call sites
 * expose_batch_integrators.cpp:243-314).  Lanes are independent; OpenMP over
 * lanes plays the role of the reference's thread-pool ensemble
 * (_ensemble_impl.py:23-68).  state is [n, B] lane-fastest.  When `single_step`
 * is non-zero exactly one step is taken per lane with `t` ignored and
 * max_delta_t as the (signed) limit: the reference's step().
 * tc (nullable) is [n, p+1, B]; h_log (nullable) is [B, h_log_cap].
 */
int FN(ora_propagate)(const hy_dims *d, const hy_op *ops, const hy_term *terms,
                      const uint32_t *ev_ref, double tol, int high_accuracy, uint32_t B, REAL *state,
                      const REAL *pars, REAL *t_hi, REAL *t_lo, const REAL *t, int is_delta,
                      uint64_t max_steps, const REAL *max_delta_t, int single_step, int backward,
                      int64_t *outcome, REAL *min_h, REAL *max_h, uint64_t *n_steps, REAL *last_h,
                      REAL *tc, REAL *h_log, uint64_t h_log_cap, int nthreads)
{
    if (d->order > HY_MAX_ORDER || d->order < 2) return 1;
    FN(ora_sys) S;
    S.d = d;
    S.ops = ops;
    S.terms = terms;
    S.ev_ref = ev_ref;
    S.high_accuracy = high_accuracy;
    uint32_t p = d->order, P1 = p + 1, n = d->n_state;
    S.rk[0] = 0;
    for (uint32_t k = 1; k <= p + 1; ++k) S.rk[k] = (REAL)(1.0 / (double)k);
    S.rhofac = (REAL)(exp(-7.0 / (10.0 * (p - 1.0))) / (M_E * M_E));
    S.inv_p = (REAL)(1.0 / p);
    S.inv_pm1 = (REAL)(1.0 / (p - 1.0));
    (void)tol;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    int err = 0;
#pragma omp parallel
    {
        REAL *ws = (REAL *)calloc(d->n_rows + 1, sizeof(REAL));
        REAL *lp = (REAL *)calloc(d->n_par + 1, sizeof(REAL));
        if (!ws || !lp) {
#pragma omp atomic write
            err = 2;
        } else {
#pragma omp for schedule(dynamic, 16)
            for (int64_t l = 0; l < (int64_t)B; ++l) {
                for (uint32_t i = 0; i < n; ++i) ws[i * P1] = state[(size_t)i * B + l];
                for (uint32_t i = 0; i < d->n_par; ++i) lp[i] = pars[(size_t)i * B + l];
                REAL hi = t_hi[l], lo = t_lo[l];
                REAL mdt = max_delta_t ? max_delta_t[l] : (REAL)INFINITY;
                int64_t oc = HY_OUTCOME_TIME_LIMIT;
                REAL mn = (REAL)INFINITY, mx = 0, h = 0;
                uint64_t ns = 0;
                if (single_step) {
                    REAL lim = max_delta_t ? mdt : (backward ? -(REAL)INFINITY : (REAL)INFINITY);
                    oc = FN(lane_step)(&S, ws, lp, &hi, &lo, lim, &h, tc, B, (size_t)l);
                    ns = 1;
                    if (h_log && h_log_cap) h_log[(size_t)l * h_log_cap] = h;
                } else {
                    REAL tf_hi, tf_lo = 0;
                    if (is_delta) {
                        tf_hi = hi;
                        tf_lo = lo;
                        FN(time_add)(&tf_hi, &tf_lo, t[l]);
                    } else {
                        tf_hi = t[l];
                    }
                    mdt = R_ABS(mdt);
                    for (;;) {
                        REAL rem = FN(time_sub)(tf_hi, tf_lo, hi, lo);
                        if (rem == 0) {
                            oc = HY_OUTCOME_TIME_LIMIT;
                            break;
                        }
                        REAL lim = R_ABS(rem) < mdt ? rem : R_COPYSIGN(mdt, rem);
                        int64_t so = FN(lane_step)(&S, ws, lp, &hi, &lo, lim, &h, tc, B, (size_t)l);
                        if (h_log && ns < h_log_cap) h_log[(size_t)l * h_log_cap + ns] = h;
                        ++ns;
                        if (so == HY_OUTCOME_ERR_NF_STATE) {
                            oc = so;
                            break;
                        }
                        if (so == HY_OUTCOME_SUCCESS) {
                            REAL ah = R_ABS(h);
                            if (ah < mn) mn = ah;
                            if (ah > mx) mx = ah;
                        }
                        if (so == HY_OUTCOME_TIME_LIMIT && h == rem) {
                            hi = tf_hi;
                            lo = tf_lo;
                            oc = HY_OUTCOME_TIME_LIMIT;
                            break;
                        }
                        if (max_steps && ns >= max_steps) {
                            oc = HY_OUTCOME_STEP_LIMIT;
                            break;
                        }
                    }
                }
                for (uint32_t i = 0; i < n; ++i) state[(size_t)i * B + l] = ws[i * P1];
                t_hi[l] = hi;
                t_lo[l] = lo;
                if (outcome) outcome[l] = oc;
                if (min_h) min_h[l] = mn;
                if (max_h) max_h[l] = mx;
                if (n_steps) n_steps[l] = ns;
                if (last_h) last_h[l] = h;
            }
        }
        free(ws);
        free(lp);
    }
    return err;
}
