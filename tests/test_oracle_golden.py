"""Pin the oracles (oracle/np_oracle.py: DAG walker, oracle/hy_oracle.c: tape
interpreter) to the absolute values printed in the reference's notebooks
(tests/golden/notebook_golden.json, produced by make_notebook_golden.py).
CPU only."""

import json
import os

import numpy as np
import pytest

import hy_b200 as hy
from hy_b200 import decompose as D
from oracle.np_oracle import NpTaylorBatch
from oracle.c_oracle import COracle

import common

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "notebook_golden.json")))


def ulps(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b) / np.spacing(np.abs(b)))


def make(kind, sys_, ic, order=None, **kw):
    ic = np.asarray(ic, dtype=float)
    if ic.ndim == 1:
        ic = ic[:, None]
    if kind == "np":
        return NpTaylorBatch(sys_, ic, **kw)
    tol = kw.get("tol", 0.0)
    p = D.taylor_order(tol if tol else np.finfo(float).eps)
    return COracle(D.decompose(sys_, p), ic, **kw)


@pytest.mark.parametrize("kind", ["np", "c"])
def test_scalar_pendulum_A1_A3(kind):
    g = G["pendulum_scalar"]
    sys_ = common.pendulum_sys()
    ta = make(kind, sys_, g["ic"])
    oc, h = ta.step()
    assert oc[0] == hy.taylor_outcome.success
    assert ulps(h[0], g["first_step_h"]) <= 2
    assert ulps(ta.state[:, 0], g["first_step_state"]) <= 4
    # step backward, then two clamped steps (A2)
    oc, h = ta.step(backward=True)
    assert ulps(h[0], g["step_backward_h"]) <= 2
    oc, h = ta.step(max_delta_t=[0.01])
    assert oc[0] == hy.taylor_outcome.time_limit and h[0] == 0.01
    oc, h = ta.step(max_delta_t=[-0.02])
    assert oc[0] == hy.taylor_outcome.time_limit and h[0] == -0.02
    assert abs(ta.t_hi[0] - g["time_after_clamped_steps"]) < 1e-16
    # A3
    ta = make(kind, sys_, g["ic"])
    for call, arg, key in (("propagate_for", 5.0, "propagate_for_5"),
                           ("propagate_until", 20.0, "propagate_until_20"),
                           ("propagate_until", 0.0, "propagate_until_0")):
        r = getattr(ta, call)(arg)
        oc, mn, mx, ns = r[0], r[1], r[2], r[3]
        assert oc[0] == hy.taylor_outcome.time_limit
        assert int(ns[0]) == g[key][0]
        assert ulps(mn[0], g[key][1]) <= 8 and ulps(mx[0], g[key][2]) <= 8
    assert np.max(np.abs(ta.state[:, 0] - g["final_state"])) < 5e-16
    assert ta.t_hi[0] == 0.0


@pytest.mark.parametrize("kind", ["np", "c"])
def test_tol_1e9_A5(kind):
    g = G["pendulum_tol1e-9"]
    assert D.taylor_order(1e-9) == g["order"]
    ta = make(kind, common.pendulum_sys(), [0.05, 0.025], tol=1e-9)
    ta.propagate_until(10.0)
    ta.propagate_until(0.0)
    assert np.max(np.abs(ta.state[:, 0] - g["state_after_10_and_back"])) < 2e-14


def test_orders():
    # SURVEY.md A.2
    assert D.taylor_order(np.finfo(np.float64).eps) == 20
    assert D.taylor_order(float(np.finfo(np.float32).eps)) == 9
    assert D.taylor_order(1e-18) == 22
    assert D.taylor_order(1e-9) == 12


@pytest.mark.parametrize("kind", ["np", "c"])
def test_harmonic_first_step_A6(kind):
    g = G["harmonic"]
    x, v = hy.make_vars("x", "v")
    ta = make(kind, [(x, v), (v, -x)], [0.0, 1.0])
    oc, h = ta.step()
    assert ulps(h[0], g["first_step_h"]) <= 2
    # tc = (+-1/k!) pattern
    import math

    tc = ta.tc[:, :, 0]
    for k in range(1, 21, 2):
        assert abs(tc[0, k] - (-1) ** ((k - 1) // 2) / math.factorial(k)) < 1e-16


@pytest.mark.parametrize("kind", ["np", "c"])
def test_batch_forced_pendulum_A7_A8(kind):
    g = G["batch_forced_pendulum"]
    sys_ = common.forced_pendulum_sys()
    ta = make(kind, sys_, g["ic"], pars=g["pars"])
    oc, h = ta.step()
    assert ulps(h, g["first_step_h"]) <= 2
    assert np.max(np.abs(ta.state - np.array(g["first_step_state"]))) < 1e-8
    oc, h = ta.step(max_delta_t=g["clamped_step"])
    assert np.all(oc == hy.taylor_outcome.time_limit)
    assert np.max(np.abs(ta.state - np.array(g["clamped_state"]))) < 1e-8
    r = ta.propagate_for(g["propagate_for"]["delta"])
    res = np.array(g["propagate_for"]["res"])
    assert list(r[3]) == list(res[:, 2].astype(int))
    assert np.max(np.abs(r[1] - res[:, 0]) / res[:, 0]) < 1e-13
    assert np.max(np.abs(r[2] - res[:, 1]) / res[:, 1]) < 1e-13
    assert np.max(np.abs(ta.state - np.array(g["propagate_for"]["state"]))) < 1e-8
    r = ta.propagate_until(g["propagate_until"]["t"])
    res = np.array(g["propagate_until"]["res"])
    assert list(r[3]) == list(res[:, 2].astype(int))
    assert np.max(np.abs(r[1] - res[:, 0]) / res[:, 0]) < 1e-13
    assert np.max(np.abs(r[2] - res[:, 1]) / res[:, 1]) < 1e-13
    assert np.max(np.abs(ta.state - np.array(g["propagate_until"]["state"]))) < 1e-8
    assert np.all(ta.t_hi == np.array(g["propagate_until"]["t"]))


@pytest.mark.parametrize("kind", ["np", "c"])
def test_batch_tc_A9(kind):
    g = G["batch_forced_pendulum"]
    ta = make(kind, common.forced_pendulum_sys(), g["ic"], pars=g["pars"])
    ta.step()
    assert np.max(np.abs(ta.tc[0, 2] - np.array(g["tc_x_order2"]))) < 1e-9
    assert np.max(np.abs(ta.tc[0, 3] - np.array(g["tc_x_order3"]))) < 1e-9


@pytest.mark.parametrize("kind", ["np", "c"])
def test_cr3bp_A13(kind):
    # tol=1e-18 -> order 22; propagate_grid does not clamp at interior grid
    # points, so the step sequence equals propagate_until(200).
    g = G["cr3bp"]
    sys_ = common.cr3bp_sys(g["mu"])
    ta = make(kind, sys_, g["ic"], tol=g["tol"])
    r = ta.propagate_until(g["t_end"])
    assert int(r[3][0]) == g["steps"]
    assert abs(r[1][0] - g["min_h"]) / g["min_h"] < 1e-12
    assert abs(r[2][0] - g["max_h"]) / g["max_h"] < 1e-12
    assert np.max(np.abs(ta.state[:, 0] - np.array(g["final_state"]))) < 1e-8


def test_c_oracle_vs_np_oracle_nbody():
    # The tape lowering (fused SUMSQ/MULSH ops) against the DAG walker.
    sys_ = common.oss_sys()
    ic = common.oss_ensemble(3, amp=1e-3)
    a = COracle(D.decompose(sys_, 20), ic)
    b = NpTaylorBatch(sys_, ic)
    ra = a.propagate_until(20.0)
    rb = b.propagate_until(20.0)
    assert list(ra[3]) == list(rb[3])
    assert np.max(np.abs(a.state - b.state) / np.maximum(1, np.abs(b.state))) < 1e-13
    c = COracle(D.decompose(sys_, 20, fuse=False), ic)
    rc = c.propagate_until(20.0)
    assert list(rc[3]) == list(ra[3])
    assert np.max(np.abs(a.state - c.state) / np.maximum(1, np.abs(c.state))) < 1e-13


def test_simd_baseline_matches_scalar_oracle():
    # oracle/hy_baseline_simd.c (bench.py's cpu_baseline) against the pinned scalar oracle
    from oracle.c_oracle import simd_propagate_until

    sys_ = common.oss_sys()
    ic = common.oss_ensemble(19, amp=1e-3)   # not a multiple of the SIMD width
    dc = D.decompose(sys_, 20)
    o = COracle(dc, ic)
    r = o.propagate_until(30.0)
    st, ns = simd_propagate_until(dc, ic, 30.0, nthreads=2)
    assert np.array_equal(ns, r[3])
    assert np.max(np.abs(st - o.state) / np.maximum(1, np.abs(o.state))) < 1e-12
    # a system with sin/cos, parameters and time
    g = G["batch_forced_pendulum"]
    dc = D.decompose(common.forced_pendulum_sys(), 20)
    st, ns = simd_propagate_until(dc, g["ic"], [10.0, 11.0, 12.0, 13.0], pars=g["pars"])
    o = COracle(dc, g["ic"], pars=g["pars"])
    r = o.propagate_until([10.0, 11.0, 12.0, 13.0])
    assert np.array_equal(ns, r[3]) and np.max(np.abs(st - o.state)) < 1e-12


@pytest.mark.parametrize("kind", ["np", "c"])
def test_variational_order2_A12(kind):
    # var_ode_sys.ipynb:361,529,707: second-order variational equations of the forced damped
    # pendulum - the 12-vector at t = 3, the order-2 slice / multi-indices, the Taylor map.
    g = G["var_pendulum_order2"]
    x, v = hy.make_vars("x", "v")
    sys_ = [(x, v), (v, hy.cos(hy.time) - hy.par[0] * v - hy.sin(x))]
    vs = hy.var_ode_sys(sys_, hy.var_args.vars, order=g["order"])
    assert vs.order == 2 and vs.n_orig_sv == 2 and [a.name for a in vs.vargs] == ["x", "v"]
    assert len(vs.sys) == 12
    ic = np.zeros(12)
    ic[:2] = g["ic"]
    ic[2:] = vs._initial_var_state(float)
    assert list(ic) == g["initial_state"]
    sl = vs.get_vslice(2)
    assert [sl.start, sl.stop] == g["order2_slice"]
    assert [vs.get_mindex(i) for i in range(sl.start, sl.stop)] == g["order2_mindex"]
    ta = make(kind, vs.sys, ic, pars=np.array([[g["par"]]]))
    ta.propagate_until(g["t_end"])
    st = ta.state[:, 0]
    assert np.max(np.abs(st - np.array(g["final_state_8digits"]))) < 6e-9
    assert np.max(np.abs(st[sl] - np.array(g["order2_values"])) / np.abs(g["order2_values"])) < 1e-13
    tm = vs.eval_taylor_map(ta.state, np.array(g["taylor_map_inputs"])[:, None])[:, 0]
    assert np.max(np.abs(tm - np.array(g["taylor_map_state_8digits"]))) < 6e-9
    # the Taylor map against a run from the displaced initial conditions (the notebook prints 6.6e-13)
    tb = make(kind, sys_, np.array(g["ic"]) + np.array(g["taylor_map_inputs"]), pars=np.array([[g["par"]]]))
    tb.propagate_until(g["t_end"])
    err = tm - tb.state[:, 0]
    assert np.max(np.abs(err - np.array(g["taylor_map_error_vs_displaced_run"]))) < 5e-15
