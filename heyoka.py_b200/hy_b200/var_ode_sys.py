"""First-order variational equations (``var_ode_sys``).

Mirror of /root/reference/heyoka/expose_var_ode_sys.cpp:29-58: the state is
augmented with the sensitivities d x_i / d a_j of the solution w.r.t. the
selected arguments (initial conditions and/or parameters).  Ordering of the
augmented state follows the reference (_test_var_integrator.py:190-201): by
total order, then by component, then reverse-lexicographic multi-index - for
order 1 that is, for each component i, d x_i/d a_0, d x_i/d a_1, ...

Only order 1 is built; higher orders raise.
"""

import enum

import numpy as np

from . import _expression as E


class var_args(enum.IntFlag):
    vars = 1
    params = 2
    time = 4
    all = 7


class var_ode_sys:
    def __init__(self, sys, args, order=1):
        sys = [(l, E._wrap(r)) for l, r in sys]
        if order < 1:
            raise ValueError("The 'order' argument to the var_ode_sys constructor must be nonzero")
        if order != 1:
            raise NotImplementedError(
                "var_ode_sys: only first-order variational equations are available in this build"
            )
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
            if args & var_args.time:
                raise NotImplementedError("var_ode_sys: derivatives w.r.t. the initial time")
            al = []
            if args & var_args.vars:
                al += [E.expression(nm) for nm in names]
            if args & var_args.params:
                al += [E.par[i] for i in range(npar)]
        else:
            al = list(args)
            if not al:
                raise ValueError("Cannot formulate the variational equations with an empty list of arguments")
            for a in al:
                if not isinstance(a, E.expression) or a.kind not in ("var", "par"):
                    raise ValueError(
                        "var_ode_sys: the arguments must be state variables or parameters"
                    )
                if a.kind == "var" and a.name not in names:
                    raise ValueError("var_ode_sys: '{}' is not a state variable".format(a.name))
            if len(set(id(a) for a in al)) != len(al):
                raise ValueError("Duplicate entries detected in the list of variational arguments")
        self.vargs_list = al
        self.n_orig_sv = n
        self.order = order
        self._names = names
        na = len(al)
        # Sensitivity variables, component-major.
        svar = [[E.expression("d{}_d{}".format(names[i], _aname(a))) for a in al] for i in range(n)]
        jac = [[E.diff(rhs[i], E.expression(names[k])) for k in range(n)] for i in range(n)]
        eqs = list(sys)
        for i in range(n):
            for j, a in enumerate(al):
                terms = [jac[i][k] * svar[k][j] for k in range(n)]
                if a.kind == "par":
                    terms.append(E.diff(rhs[i], a))
                eqs.append((svar[i][j], E.sum(terms)))
        self.sys = eqs
        self._na = na

    @property
    def vargs(self):
        return self.vargs_list

    def _initial_var_state(self, fp):
        n, na = self.n_orig_sv, self._na
        ic = np.zeros(n * na, dtype=fp)
        for i in range(n):
            for j, a in enumerate(self.vargs_list):
                if a.kind == "var" and a.name == self._names[i]:
                    ic[i * na + j] = 1
        return ic

    def get_vslice(self, order, component=None):
        n, na = self.n_orig_sv, self._na
        if order > self.order:
            raise ValueError(
                "The derivative order {} is larger than the maximum order {}".format(order, self.order)
            )
        if component is not None and not (0 <= component < n):
            raise ValueError("Invalid component {}".format(component))
        if order == 0:
            return slice(0, n) if component is None else slice(component, component + 1)
        if component is None:
            return slice(n, n + n * na)
        return slice(n + component * na, n + (component + 1) * na)

    def get_mindex(self, i):
        n, na = self.n_orig_sv, self._na
        if not (0 <= i < n + n * na):
            raise IndexError("Invalid index {} passed to get_mindex()".format(i))
        if i < n:
            return [i] + [0] * na
        i -= n
        comp, j = divmod(i, na)
        mi = [0] * na
        mi[j] = 1
        return [comp] + mi

    def eval_taylor_map(self, state, inputs):
        """x_i + sum_j (d x_i / d a_j) * delta_j   (first order)."""
        n, na = self.n_orig_sv, self._na
        sens = state[n:].reshape(n, na, -1)
        return state[:n] + np.einsum("ijb,jb->ib", sens, inputs)

    def __repr__(self):
        return "var_ode_sys(order={}, n_orig_sv={}, vargs={})".format(
            self.order, self.n_orig_sv, self.vargs_list
        )


def _aname(a):
    return a.name if a.kind == "var" else "p{}".format(a.value)
