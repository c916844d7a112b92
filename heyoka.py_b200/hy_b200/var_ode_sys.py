"""Variational equations of arbitrary order (``var_ode_sys``).

Mirror of /root/reference/heyoka/expose_var_ode_sys.cpp:29-58: the state is
augmented with the sensitivities d x_i / d a_j of the solution w.r.t. the
selected arguments (initial conditions and/or parameters).  Ordering of the
augmented state follows the reference (_test_var_integrator.py:190-201): by
total order, then by component, then reverse-lexicographic multi-index - for
order 1 that is, for each component i, d x_i/d a_0, d x_i/d a_1, ...

Any order is built by repeated total differentiation of the equations one order below.
"""

import enum

import numpy as np

from . import _expression as E


class var_args(enum.IntFlag):
    vars = 1
    params = 2
    time = 4
    all = 7


def _mindices(na, m):
    """All multi-indices of total order m over na arguments, in the reference's order
    (descending lexicographic: (2,0), (1,1), (0,2))."""
    out = []

    def rec(pos, left, cur):
        if pos == na - 1:
            out.append(tuple(cur + [left]))
            return
        for v in range(left, -1, -1):
            rec(pos + 1, left - v, cur + [v])

    rec(0, m, [])
    return out


def _vname(base, alpha):
    """Reference naming: the sparse multi-index, e.g. '∂[(0, 1), (1, 1)]x' (var_ode_sys.ipynb:229-262)."""
    sp = ", ".join("({}, {})".format(j, a) for j, a in enumerate(alpha) if a)
    return "∂[{}]{}".format(sp, base)


class var_ode_sys:
    """Variational equations of arbitrary order (reference: expose_var_ode_sys.cpp:29-58).

    The augmented state is ordered by total differentiation order, then by component, then by
    descending lexicographic multi-index (var_ode_sys.ipynb:361-529, _test_var_integrator.py:190-201).
    The equation of d^alpha x_i is the total derivative, w.r.t. one of the arguments, of the
    equation one order below: d/da_j g = sum_s (dg/ds) d_j s (+ dg/da_j for a parameter), where s
    runs over the (variational) state symbols of g and d_j of the symbol d^beta x_k is the symbol
    d^(beta + e_j) x_k."""

    def __init__(self, sys, args, order=1):
        sys = [(l, E._wrap(r)) for l, r in sys]
        if order < 1:
            raise ValueError("The 'order' argument to the var_ode_sys constructor must be nonzero")
        n = len(sys)
        names = [l.name for l, _ in sys]
        rhs = [r for _, r in sys]
        # Parameters appearing in the system.
        npar = 0
        for nd in E.topo_order(rhs):
            if nd.kind == "par":
                npar = max(npar, nd.value + 1)
        self._vargs_in = args
        if isinstance(args, var_args):
            if int(args) == 0 or int(args) > 7:
                raise ValueError("Invalid var_args enumerator detected")
            al = []
            if args & var_args.vars:
                al += [E.expression(nm) for nm in names]
            if args & var_args.params:
                al += [E.par[i] for i in range(npar)]
            if args & var_args.time:
                al += [E.time]
            if not al:
                raise ValueError("Cannot formulate the variational equations with an empty list of arguments")
        else:
            al = list(args)
            if not al:
                raise ValueError("Cannot formulate the variational equations with an empty list of arguments")
            for a in al:
                if not isinstance(a, E.expression) or a.kind not in ("var", "par", "time"):
                    raise ValueError(
                        "var_ode_sys: the arguments must be state variables, parameters or the time"
                    )
                if a.kind == "var" and a.name not in names:
                    raise ValueError("var_ode_sys: '{}' is not a state variable".format(a.name))
            if len(set(("time",) if a.kind == "time" else (a.kind, a.name if a.kind == "var" else a.value) for a in al)) != len(al):
                raise ValueError("Duplicate entries detected in the list of variational arguments")
        self.vargs_list = al
        self.n_orig_sv = n
        self.order = order
        self._names = names
        na = len(al)
        self._na = na
        # multi-indices per total order; position of every (order, component, alpha) in the state
        self._mi = [[tuple([0] * na)]] + [_mindices(na, m) for m in range(1, order + 1)]
        self._off = [0]
        for m in range(order + 1):
            self._off.append(self._off[-1] + n * len(self._mi[m]))
        zero = tuple([0] * na)
        sym = {(i, zero): E.expression(names[i]) for i in range(n)}
        for m in range(1, order + 1):
            for i in range(n):
                for al_ in self._mi[m]:
                    sym[(i, al_)] = E.expression(_vname(names[i], al_))
        owner = {sym[k].name: k for k in sym}  # symbol name -> (component, alpha)

        def total_diff(g, j):
            a = al[j]
            terms = []
            for nm in E.get_variables([g]):
                k, beta = owner[nm]
                b2 = list(beta)
                b2[j] += 1
                dg = E.diff(g, sym[(k, beta)])
                if dg.kind == "num" and dg.value == 0.0:
                    continue
                terms.append(dg * sym[(k, tuple(b2))])
            if a.kind == "par":
                dg = E.diff(g, a)
                if not (dg.kind == "num" and dg.value == 0.0):
                    terms.append(dg)
            return E.sum(terms) if terms else E.expression(0.0)

        eq = {(i, zero): rhs[i] for i in range(n)}
        eqs = list(sys)
        for m in range(1, order + 1):
            for i in range(n):
                for al_ in self._mi[m]:
                    # differentiate the equation one order below w.r.t. the LAST argument of alpha
                    j = max(q for q in range(na) if al_[q] > 0)
                    lower = list(al_)
                    lower[j] -= 1
                    eq[(i, al_)] = total_diff(eq[(i, tuple(lower))], j)
                    eqs.append((sym[(i, al_)], eq[(i, al_)]))
        self.sys = eqs
        # ---- the initial time as an argument (var_args.time): d^beta x / ... d t0^m of the flow x(t; x0, alpha, t0).
        # The equations need nothing new - the right-hand side does not depend on t0, so only the chain terms above
        # appear - but the initial conditions do: differentiating the identity x(t0; x0, alpha, t0) = x0 gives, for a
        # multi-index gamma and beta = gamma + e_t0,
        #     IC_beta = d/dt0 [IC_gamma](x0, alpha, t0)  -  (equation of gamma, evaluated on the initial conditions)
        # e.g. dx/dt0 = -f(x0, alpha, t0), d2x/dt0 dx0_j = -df/dx_j, d2x/dt0^2 = -f_t + f_x f.  They are built here as
        # expressions of the original state symbols, the parameters and the time, and evaluated by the integrator's
        # constructor on its initial state (`_initial_var_state_at`).  (Derived from the definition of the flow in
        # var_ode_sys.ipynb; the reference holds no numerical value for a time argument to pin this against - the
        # test checks it against finite differences of the flow.)
        self._jt = [j for j, a in enumerate(al) if a.kind == "time"]
        self._ic_sym = None
        if self._jt:
            jt = self._jt[0]
            ic = {}
            for m in range(1, order + 1):
                for i in range(n):
                    for al_ in self._mi[m]:
                        if al_[jt] == 0:
                            # no time component: identity for d x_i / d x_i(0), zero otherwise
                            one = m == 1 and al[al_.index(1)].kind == "var" and al[al_.index(1)].name == names[i]
                            ic[(i, al_)] = E.expression(1.0 if one else 0.0)
            smap_names = {}
            for m in range(1, order + 1):
                for i in range(n):
                    for al_ in self._mi[m]:
                        if al_[jt] == 0:
                            continue
                        g = list(al_)
                        g[jt] -= 1
                        g = tuple(g)
                        lower = E.expression(names[i]) if g == zero else ic[(i, g)]
                        rhs_g = eq[(i, g)]
                        sub = {sym[k].name: v for k, v in ic.items()}
                        val = E.diff(lower, E.time) - (E.subs(rhs_g, sub) if sub else rhs_g)
                        ic[(i, al_)] = val
            self._ic_sym = ic

    @property
    def vargs(self):
        return self.vargs_list

    def _initial_var_state(self, fp):
        n, na = self.n_orig_sv, self._na
        ic = np.zeros(self._off[-1] - n, dtype=fp)
        for i in range(n):
            for j, a in enumerate(self.vargs_list):
                if a.kind == "var" and a.name == self._names[i]:
                    ic[i * na + j] = 1  # order 1: d x_i / d x_i(0) = 1; every higher order starts at 0
        return ic

    def _initial_var_state_at(self, x0, pars, t0, fp):
        """Initial conditions of the variational variables for the given initial state [n, B], parameters
        [m, B] and times [B] (they depend on them only when the initial time is an argument)."""
        B = x0.shape[1]
        out = np.repeat(self._initial_var_state(fp)[:, None], B, axis=1)
        if self._ic_sym is None:
            return out
        n = self.n_orig_sv
        vv = {nm: np.asarray(x0[i], dtype=np.float64) for i, nm in enumerate(self._names)}
        pp = [np.asarray(p, dtype=np.float64) for p in pars]
        tt = np.asarray(t0, dtype=np.float64)
        for m in range(1, self.order + 1):
            nm_ = len(self._mi[m])
            for i in range(n):
                for q, al_ in enumerate(self._mi[m]):
                    if al_[self._jt[0]] == 0:
                        continue
                    v = E.eval_numpy(self._ic_sym[(i, al_)], vv, pars=pp, tm=tt)
                    out[self._off[m] - n + i * nm_ + q] = np.broadcast_to(np.asarray(v, dtype=fp), (B,))
        return out

    def get_vslice(self, order, component=None):
        n = self.n_orig_sv
        if order > self.order:
            raise ValueError(
                "The derivative order {} is larger than the maximum order {}".format(order, self.order)
            )
        if component is not None and not (0 <= component < n):
            raise ValueError("Invalid component {}".format(component))
        nm = len(self._mi[order])
        if component is None:
            return slice(self._off[order], self._off[order + 1])
        return slice(self._off[order] + component * nm, self._off[order] + (component + 1) * nm)

    def get_mindex(self, i):
        if not (0 <= i < self._off[-1]):
            raise IndexError("Invalid index {} passed to get_mindex()".format(i))
        m = max(q for q in range(self.order + 1) if self._off[q] <= i)
        comp, j = divmod(i - self._off[m], len(self._mi[m]))
        return [comp] + list(self._mi[m][j])

    def eval_taylor_map(self, state, inputs):
        """x_i + sum_{|alpha| >= 1} (d^alpha x_i / alpha!) delta^alpha
        (reference: taylor map evaluation, var_ode_sys.ipynb:620-707)."""
        import math

        n = self.n_orig_sv
        out = np.array(state[:n], dtype=state.dtype, copy=True)
        for m in range(1, self.order + 1):
            mis = self._mi[m]
            for i in range(n):
                base = self._off[m] + i * len(mis)
                for j, al_ in enumerate(mis):
                    mono = 1.0
                    fact = 1.0
                    for q, e in enumerate(al_):
                        if e:
                            mono = mono * inputs[q] ** e
                            fact *= math.factorial(e)
                    out[i] = out[i] + state[base + j] * mono / fact
        return out

    def __repr__(self):
        return "var_ode_sys(order={}, n_orig_sv={}, vargs={})".format(
            self.order, self.n_orig_sv, self.vargs_list
        )


def _aname(a):
    return a.name if a.kind == "var" else "p{}".format(a.value)
