"""GPU: variational equations of order 2 and the Taylor map, pinned on the reference notebook
(var_ode_sys.ipynb:361,529,707 - golden A12) and on the reference's own batch test
(/root/reference/heyoka/_test_var_integrator.py:144-258: layout, tstate, eval_taylor_map errors)."""

import json
import os

import numpy as np
import pytest

import hy_b200 as hy
from hy_b200 import decompose as D
from oracle.c_oracle import COracle

pytestmark = pytest.mark.gpu

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "notebook_golden.json")))["var_pendulum_order2"]


def _sys():
    x, v = hy.make_vars("x", "v")
    return [(x, v), (v, hy.cos(hy.time) - hy.par[0] * v - hy.sin(x))], x, v


def test_order2_notebook_golden_A12():
    sys_, x, v = _sys()
    vs = hy.var_ode_sys(sys_, hy.var_args.vars, order=2)
    B = 5
    ic = np.array(G["ic"])[:, None] * np.ones((1, B))
    ta = hy.taylor_adaptive_batch(vs, ic, pars=np.full((1, B), G["par"]), compact_mode=True)
    assert ta.dim == 12 and ta.n_orig_sv == 2 and ta.is_variational and ta.vorder == 2
    assert ta.vargs == [x, v]
    assert np.array_equal(ta.state[:, 0], G["initial_state"])
    ta.propagate_until(G["t_end"])
    sl = ta.get_vslice(order=2)
    assert [sl.start, sl.stop] == G["order2_slice"]
    assert [ta.get_mindex(i) for i in range(sl.start, sl.stop)] == G["order2_mindex"]
    for b in range(B):
        assert np.max(np.abs(ta.state[:, b] - np.array(G["final_state_8digits"]))) < 6e-9
        assert np.max(np.abs(ta.state[sl, b] - np.array(G["order2_values"])) / np.abs(G["order2_values"])) < 1e-12
    # Taylor map: the notebook's displaced run
    dx = np.array(G["taylor_map_inputs"])[:, None] * np.ones((1, B))
    ts = ta.eval_taylor_map(dx)
    assert np.max(np.abs(ts[:, 0] - np.array(G["taylor_map_state_8digits"]))) < 6e-9
    tb = hy.taylor_adaptive_batch(sys_, ic + dx, pars=np.full((1, B), G["par"]))
    tb.propagate_until(G["t_end"])
    err = ts - tb.state
    assert np.max(np.abs(err[:, 0] - np.array(G["taylor_map_error_vs_displaced_run"]))) < 2e-14
    # against the C oracle on the same 12-variable tape
    orc = COracle(D.decompose(vs.sys, ta.order), np.array(G["initial_state"])[:, None] * np.ones((1, B)),
                  pars=np.full((1, B), G["par"]))
    oc, mn, mx, ns, _ = orc.propagate_until(G["t_end"])
    assert list(ta.propagate_res_arrays[3]) == list(ns)
    assert np.max(np.abs(ta.state - orc.state)) < 1e-12


def test_reference_batch_test_layout_and_errors():
    # _test_var_integrator.py:144-258
    sys_, x, v = _sys()
    vsys = hy.var_ode_sys(sys_, hy.var_args.vars, order=2)
    ta = hy.taylor_adaptive_batch(vsys, [[0.2, 0.21], [0.3, 0.31]], pars=[[0.4, 0.41]], time=[0.5, 0.51],
                                  compact_mode=True)
    assert ta.dim > 2 and ta.n_orig_sv == 2 and ta.is_variational and ta.vorder == 2 and ta.vargs == [x, v]
    ts = ta.tstate
    assert ts.shape == (2, 2) and np.all(ts == 0.0)
    with pytest.raises(ValueError):
        ta.tstate[0] = 0.5
    assert ta.get_vslice(order=0) == slice(0, 2, None)
    assert ta.get_vslice(order=0, component=1) == slice(1, 2, None)
    assert ta.get_vslice(order=1) == slice(2, 6, None)
    assert ta.get_vslice(order=1, component=1) == slice(4, 6, None)
    assert [ta.get_mindex(i) for i in range(6)] == [[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [1, 0, 1]]
    ta.propagate_until(3.0)
    ts2 = ta.eval_taylor_map([[0.0, 0.0], [0.0, 0.0]])
    assert np.shares_memory(ts, ts2) and np.all(ts2 == ta.state[:2])
    ts2 = ta.eval_taylor_map(np.array([[0.0, 0.0], [0.0, 0.0]]))
    assert np.shares_memory(ts, ts2) and np.all(ts2 == ta.state[:2])
    with pytest.raises(TypeError, match="Invalid dtype detected for the inputs of a Taylor map evaluation:"):
        ta.eval_taylor_map(np.array([0.0, 0.0], dtype=np.int32))
    with pytest.raises(ValueError, match="the array is not C-style contiguous, please "):
        ta.eval_taylor_map(np.array([0.0, 0.0, 0.0, 0.0])[::2])
    with pytest.raises(ValueError, match=r"has 1 dimension\(s\), "):
        ta.eval_taylor_map(np.array([0.0, 0.0]))
    with pytest.raises(ValueError, match=r"has 1 row\(s\), but it must have 2 row\(s\) instead"):
        ta.eval_taylor_map(np.array([[0.0, 0.0]]))
    with pytest.raises(ValueError, match=r"has 1 column\(s\), but it must have 2 column\(s\) instead"):
        ta.eval_taylor_map(np.array([[0.0], [0.0]]))
    with pytest.raises(ValueError, match="may overlap"):
        ta.eval_taylor_map(ta.state[:2])
    with pytest.raises(ValueError, match="may overlap"):
        ta.eval_taylor_map(ta.tstate)


def test_order2_wrt_parameter_matches_finite_differences():
    # d x / d alpha and d2 x / d alpha2 of the damped pendulum against central differences of runs
    sys_, x, v = _sys()
    vs = hy.var_ode_sys(sys_, [hy.par[0]], order=2)
    ic = np.array([[0.2], [0.3]])
    ta = hy.taylor_adaptive_batch(vs, ic, pars=[[0.4]])
    ta.propagate_until(2.0)
    e = 1e-4
    runs = []
    for a in (0.4 - e, 0.4, 0.4 + e):
        t = hy.taylor_adaptive_batch(sys_, ic, pars=[[a]])
        t.propagate_until(2.0)
        runs.append(t.state[:, 0].copy())
    d1 = (runs[2] - runs[0]) / (2 * e)
    d2 = (runs[2] - 2 * runs[1] + runs[0]) / (e * e)
    assert np.max(np.abs(ta.state[ta.get_vslice(1), 0] - d1)) < 1e-7
    assert np.max(np.abs(ta.state[ta.get_vslice(2), 0] - d2)) < 1e-5
