"""The initial time as a variational argument (var_args.time / `time` in the argument list).

CPU: the reference's own var_ode_sys scenarios (/root/reference/heyoka/_test_var_ode_sys.py:14-78), and the
formulation checked against finite differences of the flow x(t; x0, t0) on the numpy oracle.
GPU: the integrator fills the initial conditions of the time derivatives from its initial state / pars / time."""

from copy import copy, deepcopy
from pickle import dumps, loads

import numpy as np
import pytest

import hy_b200 as hy
from hy_b200 import make_vars, par, sin, time, var_args, var_ode_sys
from oracle.np_oracle import NpTaylorBatch


def test_reference_var_ode_sys_scenarios():
    assert var_args.vars | var_args.time | var_args.params == var_args.all
    assert (var_args.vars | var_args.time) & var_args.time
    x, v = make_vars("x", "v")
    orig_sys = [(x, v), (v, -par[0] * sin(x) + time)]
    vsys = var_ode_sys(orig_sys, var_args.vars)
    assert orig_sys == vsys.sys[:2] and vsys.vargs == [x, v] and vsys.n_orig_sv == 2 and vsys.order == 1
    vsys = var_ode_sys(sys=orig_sys, args=[v, time, x], order=2)
    assert orig_sys == vsys.sys[:2] and vsys.vargs == [v, time, x] and vsys.n_orig_sv == 2 and vsys.order == 2
    vsys = var_ode_sys(orig_sys, var_args.vars | var_args.params)
    assert vsys.vargs == [x, v, par[0]] and vsys.order == 1
    for f in (copy, deepcopy, lambda o: loads(dumps(o))):
        v2 = f(vsys)
        assert orig_sys == v2.sys[:2] and v2.vargs == [x, v, par[0]] and v2.n_orig_sv == 2 and v2.order == 1
    vall = var_ode_sys(orig_sys, var_args.all)
    assert vall.vargs == [x, v, par[0], time]


def _setup():
    x, v = make_vars("x", "v")
    orig = [(x, v), (v, hy.cos(time) - par[0] * v - sin(x))]
    return x, v, orig, var_ode_sys(orig, [x, time], order=2)


def test_time_argument_against_finite_differences_of_the_flow():
    x, v, orig, vs = _setup()
    x0, p, t0, T = np.array([[0.2], [0.3]]), np.array([[0.4]]), np.array([0.5]), 3.0
    ic = vs._initial_var_state_at(x0, p, t0, np.float64)
    # dx/dt0 (t0) = -f(x0, t0)
    f0 = np.array([0.3, np.cos(0.5) - 0.4 * 0.3 - np.sin(0.2)])
    s1 = ic[vs.get_vslice(1).start - 2: vs.get_vslice(1).stop - 2, 0]
    assert np.allclose(s1, [1.0, -f0[0], 0.0, -f0[1]], rtol=0, atol=1e-15)
    o = NpTaylorBatch(vs.sys, np.vstack([x0, ic]), time=t0, pars=p)
    o.propagate_until(T)
    st = o.state[:, 0]

    def flow(a, b, t):
        q = NpTaylorBatch(orig, np.array([[a], [b]]), time=np.array([t]), pars=p)
        q.propagate_until(T)
        return q.state[:, 0].copy()

    d = 1e-4
    c = flow(0.2, 0.3, 0.5)
    assert np.allclose(st[:2], c, rtol=0, atol=1e-13)
    dt = (flow(0.2, 0.3, 0.5 + d) - flow(0.2, 0.3, 0.5 - d)) / (2 * d)
    dx = (flow(0.2 + d, 0.3, 0.5) - flow(0.2 - d, 0.3, 0.5)) / (2 * d)
    assert np.allclose(st[vs.get_vslice(1)], [dx[0], dt[0], dx[1], dt[1]], rtol=0, atol=1e-7)
    dtt = (flow(0.2, 0.3, 0.5 + d) - 2 * c + flow(0.2, 0.3, 0.5 - d)) / d**2
    dxx = (flow(0.2 + d, 0.3, 0.5) - 2 * c + flow(0.2 - d, 0.3, 0.5)) / d**2
    dxt = (flow(0.2 + d, 0.3, 0.5 + d) - flow(0.2 + d, 0.3, 0.5 - d) - flow(0.2 - d, 0.3, 0.5 + d)
           + flow(0.2 - d, 0.3, 0.5 - d)) / (4 * d * d)
    # order 2, per component: (2,0), (1,1), (0,2)
    assert np.allclose(st[vs.get_vslice(2)], [dxx[0], dxt[0], dtt[0], dxx[1], dxt[1], dtt[1]], rtol=0, atol=5e-7)


@pytest.mark.gpu
def test_time_argument_on_the_gpu_integrator():
    x, v, orig, vs = _setup()
    ic0 = np.array([[0.2, 0.21], [0.3, 0.31]])
    pars, t0 = np.array([[0.4, 0.41]]), np.array([0.5, 0.51])
    ta = hy.taylor_adaptive_batch(vs, ic0, pars=pars, time=t0)
    want = vs._initial_var_state_at(ic0, pars, t0, np.float64)
    assert np.array_equal(ta.state[2:], want) and np.all(ta.state[3] == -ic0[1])      # dx/dt0 = -v0
    ta.propagate_until(3.0)
    o = NpTaylorBatch(vs.sys, np.vstack([ic0, want]), time=t0, pars=pars)
    o.propagate_until(3.0)
    assert np.max(np.abs(ta.state - o.state) / np.maximum(1.0, np.abs(o.state))) < 1e-12
    # Taylor map in (dx0, dt0): a run displaced in x0 and started later
    dx0, dt0 = 1e-4, 2e-4
    tm = ta.eval_taylor_map(np.array([[dx0, dx0], [dt0, dt0]]))
    ref = hy.taylor_adaptive_batch(orig, ic0 + np.array([[dx0], [0.0]]), pars=pars, time=t0 + dt0)
    ref.propagate_until(3.0)
    assert np.max(np.abs(tm - ref.state)) < 1e-10
