"""numpy restatement of the reference's adaptive batch Taylor integrator.

TEST INFRASTRUCTURE ONLY.  Nothing under heyoka.py_b200/ may import this
module; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
use oracle/.  It is deliberately independent of the product's lowering
(heyoka.py_b200/hy_b200/decompose.py): jets are built by walking the
expression DAG node by node, so it also checks the tape lowering.

Where the algorithm comes from.  The arithmetic of the hot path lives in the
heyoka C++ library 7.11.x (CMakeLists.txt:127 `find_package(heyoka 7.11.0)`),
which is NOT present under /root/reference and cannot be built here.  This
file restates its published algorithm (README.md's
arXiv:2105.00800 / arXiv:2204.09948) and follows the reference's call sites:

 * constructor / order from tol ........ expose_batch_integrators.cpp:91-212
 * step(), step(max_delta_t) ........... expose_batch_integrators.cpp:233-241
 * propagate_for / propagate_until ..... expose_batch_integrators.cpp:243-314
 * propagate_grid ...................... expose_batch_integrators.cpp:315-392
 * update_d_output ..................... expose_batch_integrators.cpp:519-541
 * continuous output ................... taylor_expose_c_output.cpp:260-526
 * events .............................. taylor_expose_events.cpp:185-317

Parity status: PINNED against the absolute values the reference prints in its
notebooks by tests/test_oracle_golden.py; the
reference's unit tests hold no absolute values for this path.
"""

import math

import numpy as np

OUT_SUCCESS = -4294967297
OUT_STEP_LIMIT = -4294967298
OUT_TIME_LIMIT = -4294967299
OUT_ERR_NF_STATE = -4294967300
OUT_CB_STOP = -4294967301


def taylor_order(tol):
    # SURVEY A.2; e.g. eps64 -> 20, 1e-9 -> 12, eps32 -> 9, 1e-18 -> 22.
    return max(2, int(math.ceil(-math.log(tol) / 2.0 + 1.0)))


def _topo(roots):
    order, seen = [], set()
    stack = [(r, 0) for r in reversed(list(roots))]
    while stack:
        e, i = stack.pop()
        if i == 0:
            if id(e) in seen:
                continue
            seen.add(id(e))
        if i < len(e.args):
            stack.append((e, i + 1))
            if id(e.args[i]) not in seen:
                stack.append((e.args[i], 0))
        else:
            order.append(e)
    return order


def pow0(x, alpha, T):
    """Order-0 evaluation of x^alpha with the special cases the product uses
    (half-integer exponents by sqrt/div, the rest by pow)."""
    if alpha == -1.5:
        return T(1.0) / (x * np.sqrt(x))
    if alpha == -0.5:
        return T(1.0) / np.sqrt(x)
    if alpha == 1.5:
        return x * np.sqrt(x)
    if alpha == -1.0:
        return T(1.0) / x
    if alpha == -2.0:
        return T(1.0) / (x * x)
    return np.power(x, T(alpha))


class two_float:
    """Double-length time (hi, lo) with error-free addition (SURVEY A.6;
    reference views: expose_batch_integrators.cpp:407-427)."""

    @staticmethod
    def add(hi, lo, h):
        s = hi + h
        bb = s - hi
        err = (hi - (s - bb)) + (h - bb)
        err = err + lo
        nh = s + err
        nl = err - (nh - s)
        return nh, nl

    @staticmethod
    def sub_to_scalar(ahi, alo, bhi, blo):
        """(a - b) rounded to working precision."""
        s = ahi - bhi
        bb = s - ahi
        err = (ahi - (s - bb)) + (-bhi - bb)
        err = err + (alo - blo)
        return s + err


class NpTaylorBatch:
    """Lane-independent adaptive Taylor integrator on numpy arrays [B]."""

    def __init__(self, sys, state, time=None, pars=None, tol=0.0, fp_type=np.float64,
                 events=(), high_accuracy=False, ev_spec=None):
        # ev_spec: list (terminal events first) of dicts
        #   {"dir": -1|0|1, "terminal": bool, "cooldown": float (<0: auto)}
        # matching `events`; None = the event rows only take part in the norms.
        self.ev_spec = ev_spec
        self.ev_log = []   # (lane, ev_idx, t, d_sgn) in detection order
        self.cd = {}       # (lane, ev_idx) -> [elapsed, total]
        self.T = T = np.dtype(fp_type).type
        self.sys = list(sys)
        self.names = [lhs.name for lhs, _ in self.sys]
        self.n = len(self.sys)
        self.state = np.array(state, dtype=T).reshape(self.n, -1).copy()
        self.B = self.state.shape[1]
        eps = np.finfo(T).eps
        self.tol = float(eps) if tol == 0 else float(tol)
        self.order = taylor_order(self.tol)
        self.t_hi = np.zeros(self.B, dtype=T) if time is None else np.array(time, dtype=T)
        self.t_lo = np.zeros(self.B, dtype=T)
        self.events = list(events)  # expressions; rows take part in the norms
        self.roots = [r for _, r in self.sys] + self.events
        self.nodes = _topo(self.roots)
        npar = 0
        for nd in self.nodes:
            if nd.kind == "par":
                npar = max(npar, nd.value + 1)
        self.n_par = npar
        self.pars = (
            np.zeros((npar, self.B), dtype=T)
            if pars is None
            else np.array(pars, dtype=T).reshape(npar, self.B)
        )
        p = self.order
        # rhofac = exp(-7/(10(p-1))) / e^2   
        self.rhofac = T(math.exp(-7.0 / (10.0 * (p - 1))) / (math.e * math.e))
        self.tc = np.zeros((self.n, p + 1, self.B), dtype=T)
        self.ev_tc = np.zeros((len(self.events), p + 1, self.B), dtype=T)
        self.last_h = np.zeros(self.B, dtype=T)
        self.high_accuracy = high_accuracy

    # ---- jets ----
    def compute_jets(self, active=None):
        T, p, B, n = self.T, self.order, self.B, self.n
        sv = {nm: i for i, nm in enumerate(self.names)}
        jets = {}  # id(node) -> array [p+1, B]
        X = np.zeros((n, p + 1, B), dtype=T)
        X[:, 0, :] = self.state
        for nd in self.nodes:
            jets[id(nd)] = np.zeros((p + 1, B), dtype=T)
        cosj = {}  # id(sin/cos node arg) -> (s, c) jets
        with np.errstate(all="ignore"):
            for k in range(p + 1):
                for nd in self.nodes:
                    J = jets[id(nd)]
                    kd = nd.kind
                    if kd == "num":
                        J[k] = T(nd.value) if k == 0 else T(0)
                    elif kd == "par":
                        J[k] = self.pars[nd.value] if k == 0 else T(0)
                    elif kd == "time":
                        J[k] = self.t_hi if k == 0 else (T(1) if k == 1 else T(0))
                    elif kd == "var":
                        J[k] = X[sv[nd.name], k]
                    else:
                        a = [jets[id(c)] for c in nd.args]
                        nm = nd.name
                        if nm == "add":
                            J[k] = a[0][k] + a[1][k]
                        elif nm == "sub":
                            J[k] = a[0][k] - a[1][k]
                        elif nm == "neg":
                            J[k] = -a[0][k]
                        elif nm == "mul":
                            acc = np.zeros(B, dtype=T)
                            for j in range(k + 1):
                                acc = acc + a[0][j] * a[1][k - j]
                            J[k] = acc
                        elif nm == "div":
                            acc = a[0][k].copy()
                            for j in range(1, k + 1):
                                acc = acc - a[1][j] * J[k - j]
                            J[k] = acc / a[1][0]
                        elif nm in ("pow", "sqrt"):
                            if nm == "sqrt":
                                al = 0.5
                            else:
                                assert nd.args[1].kind == "num"
                                al = nd.args[1].value
                            b = a[0]
                            if al == 2.0:
                                acc = np.zeros(B, dtype=T)
                                for j in range(k + 1):
                                    acc = acc + b[j] * b[k - j]
                                J[k] = acc
                            elif k == 0:
                                J[0] = np.sqrt(b[0]) if al == 0.5 else pow0(b[0], al, T)
                            else:
                                acc = np.zeros(B, dtype=T)
                                for j in range(k):
                                    w = T(k * al - j * (al + 1.0))
                                    acc = acc + w * b[k - j] * J[j]
                                J[k] = acc / (T(k) * b[0])
                        elif nm == "exp":
                            if k == 0:
                                J[0] = np.exp(a[0][0])
                            else:
                                acc = np.zeros(B, dtype=T)
                                for j in range(1, k + 1):
                                    acc = acc + T(j) * a[0][j] * J[k - j]
                                J[k] = acc / T(k)
                        elif nm == "log":
                            if k == 0:
                                J[0] = np.log(a[0][0])
                            else:
                                acc = np.zeros(B, dtype=T)
                                for j in range(1, k):
                                    acc = acc + T(j) * J[j] * a[0][k - j]
                                J[k] = (a[0][k] - acc / T(k)) / a[0][0]
                        elif nm in ("sin", "cos"):
                            key = id(nd.args[0])
                            if key not in cosj:
                                cosj[key] = (
                                    np.zeros((p + 1, B), dtype=T),
                                    np.zeros((p + 1, B), dtype=T),
                                    [-1],
                                )
                            S, C, done = cosj[key]
                            if done[0] < k:
                                if k == 0:
                                    S[0] = np.sin(a[0][0])
                                    C[0] = np.cos(a[0][0])
                                else:
                                    sa = np.zeros(B, dtype=T)
                                    ca = np.zeros(B, dtype=T)
                                    for j in range(1, k + 1):
                                        ja = T(j) * a[0][j]
                                        sa = sa + ja * C[k - j]
                                        ca = ca + ja * S[k - j]
                                    S[k] = sa / T(k)
                                    C[k] = -ca / T(k)
                                done[0] = k
                            J[k] = S[k] if nm == "sin" else C[k]
                        elif nm in ("asin", "acos", "atan", "erf"):
                            # F(a) with dF/da = g(a): F[k] = (1/k) sum_{j=1..k} j a[j] g[k-j]; the jet
                            # of g is carried along with its own recurrences (square, pow / exp)
                            key = id(nd)
                            if key not in cosj:
                                cosj[key] = [np.zeros((p + 1, B), dtype=T) for _ in range(3)]
                            SQ, U, Gj = cosj[key]  # a^2, 1 -+ a^2 (or -a^2), g
                            A = a[0]
                            acc = np.zeros(B, dtype=T)
                            for j in range(k + 1):
                                acc = acc + A[j] * A[k - j]
                            SQ[k] = acc
                            if nm in ("asin", "acos"):
                                U[k] = (T(1) if k == 0 else T(0)) - SQ[k]
                                al = -0.5
                            elif nm == "atan":
                                U[k] = (T(1) if k == 0 else T(0)) + SQ[k]
                                al = -1.0
                            else:
                                U[k] = -SQ[k]
                            if nm == "erf":
                                if k == 0:
                                    Gj[0] = np.exp(U[0])
                                else:
                                    acc = np.zeros(B, dtype=T)
                                    for j in range(1, k + 1):
                                        acc = acc + T(j) * U[j] * Gj[k - j]
                                    Gj[k] = acc / T(k)
                                scale = T(2.0 / math.sqrt(math.pi))
                            else:
                                if k == 0:
                                    Gj[0] = pow0(U[0], al, T)
                                else:
                                    acc = np.zeros(B, dtype=T)
                                    for j in range(k):
                                        acc = acc + T(k * al - j * (al + 1.0)) * U[k - j] * Gj[j]
                                    Gj[k] = acc / (T(k) * U[0])
                                scale = T(-1) if nm == "acos" else T(1)
                            if k == 0:
                                f0 = {"asin": np.arcsin, "acos": np.arccos, "atan": np.arctan,
                                      "erf": np.vectorize(math.erf, otypes=[T])}[nm]
                                J[0] = f0(A[0]).astype(T)
                            else:
                                acc = np.zeros(B, dtype=T)
                                for j in range(1, k + 1):
                                    acc = acc + T(j) * A[j] * (scale * Gj[k - j])
                                J[k] = acc / T(k)
                        else:
                            raise NotImplementedError(nm)
                # State recurrence x_i[k+1] = f_i[k] / (k+1).
                if k < p:
                    for i, (_, r) in enumerate(self.sys):
                        X[i, k + 1] = jets[id(r)][k] / T(k + 1)
        EV = np.zeros((len(self.events), p + 1, B), dtype=T)
        for e, ex in enumerate(self.events):
            EV[e] = jets[id(ex)]
        return X, EV

    # ---- step size ----
    def step_size(self, X, EV):
        T, p = self.T, self.order
        rows = np.concatenate([X, EV], axis=0) if EV.shape[0] else X
        with np.errstate(all="ignore"):
            n0 = np.max(np.abs(rows[:, 0, :]), axis=0)
            npm1 = np.max(np.abs(rows[:, p - 1, :]), axis=0)
            np_ = np.max(np.abs(rows[:, p, :]), axis=0)
            num = np.where(n0 < 1, T(1), n0).astype(T)
            rho_p = np.power(num / np_, T(1.0 / p)).astype(T)
            rho_pm1 = np.power(num / npm1, T(1.0 / (p - 1))).astype(T)
            rho = np.minimum(rho_p, rho_pm1)
            h = rho * self.rhofac
        return h.astype(T)

    @staticmethod
    def horner(X, h):
        p = X.shape[1] - 1
        acc = X[:, p, :].copy()
        for k in range(p - 1, -1, -1):
            acc = acc * h + X[:, k, :]
        return acc

    step_hook = None  # optional callable(mask, X, t0_hi, t0_lo, h) run after every step

    def step(self, max_delta_t=None, backward=False, mask=None):
        """One adaptive step on the lanes selected by ``mask``.
        Returns (outcome[B], h[B])."""
        T, B = self.T, self.B
        if mask is None:
            mask = np.ones(B, dtype=bool)
        X, EV = self.compute_jets()
        h = self.step_size(X, EV)
        if max_delta_t is None:
            lim = np.full(B, -np.inf if backward else np.inf, dtype=T)
        else:
            lim = np.array(max_delta_t, dtype=T)
        neg = np.signbit(lim) | (backward & (lim == 0))
        h = np.where(neg, -h, h)
        with np.errstate(all="ignore"):
            clamp = np.abs(h) > np.abs(lim)
            # NaN step sizes must not pass silently as "success".
            h = np.where(clamp, lim, h).astype(T)
        outcome = np.where(clamp, OUT_TIME_LIMIT, OUT_SUCCESS).astype(np.int64)
        if self.ev_spec:
            # Rigorous pre-filter (not the product's interval Horner): with q(s) = g(h s),
            # |q(s) - q_0| <= sum_{k>=1} |q_k| on [0, 1], so |q_0| > sum_{k>=1} |q_k| excludes
            # a root in the step.  Lanes with a running cooldown clock are always visited.
            with np.errstate(all="ignore"):
                hp = np.abs(h.astype(np.float64))[None, :] ** np.arange(self.order + 1)[:, None]
                q = np.abs(EV.astype(np.float64)) * hp[None, :, :]
                maybe = ~(q[:, 0, :] > 1.0000001 * q[:, 1:, :].sum(axis=1))
            need = mask & np.any(maybe, axis=0)
            for (l_, _e) in self.cd:
                need[l_] = need[l_] or mask[l_]
            for l in np.nonzero(need)[0]:
                h[l], te = self._detect_lane(int(l), EV[:, :, l], h[l])
                if te >= 0:
                    outcome[l] = -te - 1
        with np.errstate(all="ignore"):
            new_state = self.horner(X, h)
        bad = ~np.all(np.isfinite(new_state), axis=0)
        outcome = np.where(bad, OUT_ERR_NF_STATE, outcome)
        upd = mask
        self.state[:, upd] = new_state[:, upd]
        self.tc[:, :, upd] = X[:, :, upd]
        if EV.shape[0]:
            self.ev_tc[:, :, upd] = EV[:, :, upd]
        if self.step_hook is not None:
            self.step_hook(upd.copy(), X, self.t_hi.copy(), self.t_lo.copy(), h.copy())
        nh, nl = two_float.add(self.t_hi, self.t_lo, h)
        self.t_hi = np.where(upd, nh, self.t_hi).astype(T)
        self.t_lo = np.where(upd, nl, self.t_lo).astype(T)
        self.last_h = np.where(upd, h, self.last_h).astype(T)
        return outcome, h

    def propagate_until(self, t, max_steps=0, max_delta_t=None, record=None):
        """SURVEY A.7.  Each lane steps until it reaches its own final time;
        finished lanes are frozen (the reference takes zero-length steps with
        them, which leaves state/time untouched).  ``record(step_idx, mask,
        h, outcome)`` is an optional per-step hook for tests."""
        T, B = self.T, self.B
        tf = np.broadcast_to(np.array(t, dtype=T), (B,)).copy()
        if max_delta_t is None:
            mdt = np.full(B, np.inf, dtype=T)
        else:
            mdt = np.abs(np.broadcast_to(np.array(max_delta_t, dtype=T), (B,))).copy()
        active = np.ones(B, dtype=bool)
        out = np.full(B, OUT_TIME_LIMIT, dtype=np.int64)
        min_h = np.full(B, np.inf, dtype=T)
        max_h = np.zeros(B, dtype=T)
        nst = np.zeros(B, dtype=np.uint64)
        zero = np.zeros(B, dtype=T)
        rem0 = two_float.sub_to_scalar(tf, zero, self.t_hi, self.t_lo)
        if not np.all(np.isfinite(tf)):
            raise ValueError("A non-finite time was passed to propagate_until()")
        active &= rem0 != 0
        it = 0
        while np.any(active):
            rem = two_float.sub_to_scalar(tf, zero, self.t_hi, self.t_lo).astype(T)
            lim = np.where(np.abs(rem) < mdt, rem, np.copysign(mdt, rem)).astype(T)
            # direction from the sign of the remaining time
            oc, h = self.step(max_delta_t=lim, backward=False, mask=active)
            nst[active] += 1
            succ = active & (oc == OUT_SUCCESS)
            ah = np.abs(h)
            min_h = np.where(succ & (ah < min_h), ah, min_h).astype(T)
            max_h = np.where(succ & (ah > max_h), ah, max_h).astype(T)
            if record is not None:
                record(it, active.copy(), h.copy(), oc.copy())
            it += 1
            bad = active & (oc == OUT_ERR_NF_STATE)
            out[bad] = OUT_ERR_NF_STATE
            active &= ~bad
            tev = active & (oc > OUT_SUCCESS)  # terminal event codes are small integers
            out[tev] = oc[tev]
            active &= ~tev
            fin = active & (oc == OUT_TIME_LIMIT) & (h == rem)
            self.t_hi = np.where(fin, tf, self.t_hi).astype(T)
            self.t_lo = np.where(fin, T(0), self.t_lo).astype(T)
            out[fin] = OUT_TIME_LIMIT
            active &= ~fin
            if max_steps:
                lim_hit = active & (nst >= max_steps)
                out[lim_hit] = OUT_STEP_LIMIT
                active &= ~lim_hit
        return out, min_h, max_h, nst

    def propagate_for(self, dt, **kw):
        T = self.T
        dt = np.broadcast_to(np.array(dt, dtype=T), (self.B,))
        hi, lo = two_float.add(self.t_hi, self.t_lo, dt)
        # final time as a working-precision number (hi part)
        return self.propagate_until(hi, **kw)

    def dense(self, t, rel_time=False):
        """update_d_output."""
        T = self.T
        t = np.broadcast_to(np.array(t, dtype=T), (self.B,))
        if rel_time:
            tau = self.last_h + t
        else:
            tau = t - (self.t_hi - self.last_h)
        return self.horner(self.tc, tau.astype(T))


    # ---- continuous output / grid built on the step hook ----
    def propagate_until_recorded(self, t, **kw):
        """propagate_until while recording, per lane, the start time and the
        Taylor coefficients of every step.  Returns (result, rec) with
        rec[lane] = list of (t0_hi, t0_lo, h, tc[n, p+1])."""
        rec = [[] for _ in range(self.B)]

        def hook(mask, X, t0h, t0l, h):
            for l in np.nonzero(mask)[0]:
                rec[l].append((t0h[l], t0l[l], h[l], X[:, :, l].copy()))

        self.step_hook = hook
        try:
            res = self.propagate_until(t, **kw)
        finally:
            self.step_hook = None
        return res, rec

    @staticmethod
    def eval_record(rec_lane, tq):
        """Continuous-output evaluation for one lane at time tq: the step whose
        [t_start, t_end) contains tq, clamped to the first/last step."""
        fwd = rec_lane[-1][2] >= 0
        s = len(rec_lane) - 1
        for i, (t0h, t0l, h, X) in enumerate(rec_lane):
            te = t0h + h
            if (tq < te) if fwd else (tq > te):
                s = i
                break
        t0h, t0l, h, X = rec_lane[s]
        tau = (tq - t0h) - t0l
        p = X.shape[1] - 1
        acc = X[:, p].copy()
        for k in range(p - 1, -1, -1):
            acc = acc * tau + X[:, k]
        return acc

    def propagate_grid(self, grid, **kw):
        """grid is [K, B] (monotonic per lane).  Steps are clamped only by the
        last grid time; interior points come from dense output."""
        T = self.T
        grid = np.array(grid, dtype=T).reshape(-1, self.B)
        K = grid.shape[0]
        out = np.full((K, self.n, self.B), np.nan, dtype=T)
        start_state = self.state.copy()
        start_t = self.t_hi.copy()
        res, rec = self.propagate_until_recorded(grid[-1], **kw)
        for l in range(self.B):
            fwd = grid[-1, l] >= start_t[l]
            for q in range(K):
                g = grid[q, l]
                if (g <= start_t[l]) if fwd else (g >= start_t[l]):
                    out[q, :, l] = start_state[:, l]
                    continue
                # the step during which g is reached (end-inclusive)
                for (t0h, t0l, h, X) in rec[l]:
                    tau = (g - t0h) - t0l
                    if abs(tau) <= abs(h):
                        p = X.shape[1] - 1
                        acc = X[:, p].copy()
                        for k in range(p - 1, -1, -1):
                            acc = acc * tau + X[:, k]
                        out[q, :, l] = acc
                        break
        return res, out


    # ---- events, independent method: the critical points of the
    # step polynomial split [0, h) into monotonic pieces; a sign change on a
    # piece brackets exactly one root, refined by bisection in extended precision.
    @staticmethod
    def poly_roots_01(q):
        """Real roots of sum q[k] s^k in [0, 1), ascending."""
        L = np.longdouble
        q = np.array(q, dtype=L)
        p = len(q) - 1
        roots = []
        if np.all(q == 0):
            return roots

        def ev(c, s):
            acc = L(0)
            for a in c[::-1]:
                acc = acc * s + a
            return acc

        if q[0] == 0:
            roots.append(0.0)
        # critical points: real roots of q' in (0, 1), found recursively
        def real_roots_in_01(c):
            c = np.array(c, dtype=L)
            while len(c) > 1 and c[-1] == 0:
                c = c[:-1]
            if len(c) <= 1:
                return []
            if len(c) == 2:
                r = -c[0] / c[1]
                return [r] if 0 < r < 1 else []
            dc = np.array([k * c[k] for k in range(1, len(c))], dtype=L)
            brk = [L(0)] + sorted(real_roots_in_01(dc)) + [L(1)]
            out = []
            for a, b in zip(brk[:-1], brk[1:]):
                fa, fb = ev(c, a), ev(c, b)
                if fa == 0 and a > 0:
                    out.append(a)
                if fa * fb < 0:
                    lo, hi = a, b
                    for _ in range(200):
                        m = (lo + hi) / 2
                        fm = ev(c, m)
                        if fm == 0 or m == lo or m == hi:
                            lo = hi = m
                            break
                        if (fm > 0) == (fa > 0):
                            lo = m
                        else:
                            hi = m
                    out.append((lo + hi) / 2)
            return out

        for r in real_roots_in_01(q):
            if 0 < r < 1:
                roots.append(float(r))
        return sorted(roots)

    def _detect_lane(self, l, ev_tc, h):
        """Events of lane l in the step [0, h).  Returns (h_eff, terminal idx or -1)."""
        T = self.T
        p = self.order
        if h == 0 or not np.isfinite(h):
            return h, -1
        cands = []
        for e, spec in enumerate(self.ev_spec):
            g = ev_tc[e].astype(np.longdouble)
            q = np.array([g[k] * np.longdouble(h) ** k for k in range(p + 1)])
            for s_ in self.poly_roots_01(q):
                dq = sum(k * q[k] * np.longdouble(s_) ** (k - 1) for k in range(1, p + 1))
                sg = int(np.sign(dq)) * (1 if h > 0 else -1)
                if spec["dir"] != 0 and spec["dir"] != sg:
                    continue
                tau = T(s_ * h)
                if spec["terminal"]:
                    cd = self.cd.get((l, e))
                    if cd is not None and abs(tau) < cd[1] - cd[0]:
                        continue
                cands.append((abs(tau), e, tau, sg, float(dq / h) if h != 0 else 0.0))
        cands.sort(key=lambda c: (c[0], c[1]))
        term = next((c for c in cands if self.ev_spec[c[1]]["terminal"]), None)
        h_eff, te = h, -1
        for c in cands:
            if term is not None and c is not term and not (c[0] < term[0]):
                continue
            if self.ev_spec[c[1]]["terminal"] and c is not term:
                continue
            self.ev_log.append((l, c[1], float((self.t_hi[l] + c[2]) + self.t_lo[l]), c[3]))
        if term is not None:
            h_eff, te = T(term[2]), term[1]
            spec = self.ev_spec[te]
            if spec["cooldown"] >= 0:
                cdv = spec["cooldown"]
            else:
                gm = max(1.0, max(abs(float(ev_tc[te][k]) * float(h) ** k) for k in range(p + 1)))
                cdv = 10 * self.tol * gm / abs(term[4]) if term[4] != 0 else 0.0
            self.cd[(l, te)] = [-abs(float(h_eff)), cdv]
        # advance the cooldown clocks
        for key in [k for k in self.cd if k[0] == l]:
            el, tot = self.cd[key]
            el += abs(float(h_eff))
            if el >= tot:
                del self.cd[key]
            else:
                self.cd[key] = [el, tot]
        return h_eff, te
