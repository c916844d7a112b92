"""CPU: host-side logic that needs no device - expression DAG, decomposition,
models, var_ode_sys, ensemble argument validation (mirrors
/root/reference/heyoka/_test_model.py:153-278, _test_ensemble.py:13-60,
_test_var_integrator.py:144-201)."""

import pickle

import numpy as np
import pytest

import hy_b200 as hy
from hy_b200 import decompose as D
from hy_b200 import _expression as E


def test_expression_basics():
    x, v = hy.make_vars("x", "v")
    assert (x + 0.0) is x and (1.0 * x) is x
    assert hy.sin(x) is hy.sin(x)  # hash-consing
    assert (x == x) is True and (x == v) is False
    assert pickle.loads(pickle.dumps(-9.8 * hy.sin(x))) is (-9.8 * hy.sin(x))
    assert "9.8000000000000007" in str(-9.8 * hy.sin(x))
    assert hy.diff(x * x, x) is (x + x) or hy.diff(x * x, x) is not None
    with pytest.raises(TypeError):
        hy.make_vars(1)


def test_models_structure():
    # /root/reference/heyoka/_test_model.py
    dyn = hy.model.nbody(2, masses=[0.0, 0.0])
    assert len(dyn) == 12
    for i in (3, 4, 5, 9, 10, 11):
        assert dyn[i][1] == hy.expression(0.0)
    assert "5.0000000000000" in str(hy.model.nbody(2, Gconst=5.0)[3][1])
    x, v = hy.make_vars("x", "v")
    dyn = hy.model.pendulum()
    assert dyn[0] == (x, v) and dyn[1][1] == -hy.sin(x)
    assert hy.model.pendulum(gconst=4.0, length=2.0)[1][1] == -2.0 * hy.sin(x)
    x, px, y = hy.make_vars("x", "px", "y")
    assert hy.model.cr3bp()[0] == (x, px + y)
    assert "0.06250000000" in str(hy.model.cr3bp(mu=1.0 / 2**4)[3][1])


def test_decomposition_nbody_shape():
    from hy_b200 import workloads as W

    dc = D.decompose(W.oss_sys(), 20)
    names = [D.OP_NAMES[int(o)] for o in dc.ops["opcode"]]
    assert names.count("sumsq") == 15 and names.count("pow") == 15 and names.count("mulsh") == 15
    assert names.count("addsub") == 45 and names.count("lincomb") == 18 and names.count("svd") == 18
    assert dc.n_rows * 8 < 20 * 1024  # fits 11 trajectories in 227 kB of shared memory
    fl, lo = dc.flops_per_step()
    assert 4e4 < fl < 6e4
    # unfused lowering is also valid
    dc2 = D.decompose(W.oss_sys(), 20, fuse=False)
    assert "mul" in [D.OP_NAMES[int(o)] for o in dc2.ops["opcode"]]


def test_decomposition_errors():
    x, v = hy.make_vars("x", "v")
    with pytest.raises(ValueError):
        D.decompose([(x, v), (x, v)], 20)
    with pytest.raises(ValueError):
        D.decompose([(x, hy.make_vars("y"))], 20)
    with pytest.raises(ValueError):
        D.decompose([], 20)


def test_small_integer_pow_is_safe_at_zero():
    x = hy.make_vars("x")
    dc = D.decompose([(x, x**3)], 8)
    names = [D.OP_NAMES[int(o)] for o in dc.ops["opcode"]]
    assert "pow" not in names and "square" in names and "mul" in names


def test_var_ode_sys_layout():
    # /root/reference/heyoka/_test_var_integrator.py:190-201 (order 1 part)
    x, v = hy.make_vars("x", "v")
    sys_ = [(x, v), (v, hy.cos(hy.time) - hy.par[0] * v - hy.sin(x))]
    vs = hy.var_ode_sys(sys_, hy.var_args.vars)
    assert vs.n_orig_sv == 2 and vs.order == 1 and len(vs.sys) == 6
    assert vs.get_vslice(order=0) == slice(0, 2)
    assert vs.get_vslice(order=1) == slice(2, 6)
    assert vs.get_vslice(order=1, component=1) == slice(4, 6)
    assert vs.get_mindex(0) == [0, 0, 0] and vs.get_mindex(2) == [0, 1, 0]
    assert vs.get_mindex(5) == [1, 0, 1]
    assert list(vs._initial_var_state(float)) == [1, 0, 0, 1]
    # order 2 (reference layout: var_ode_sys.ipynb:229-262, :529): by total order, component,
    # descending lexicographic multi-index; the mixed derivative is built once
    v2 = hy.var_ode_sys(sys_, hy.var_args.vars, order=2)
    assert len(v2.sys) == 12 and v2.get_vslice(order=2) == slice(6, 12)
    assert v2.get_vslice(order=2, component=1) == slice(9, 12)
    assert [v2.get_mindex(i) for i in range(6, 12)] == [[0, 2, 0], [0, 1, 1], [0, 0, 2], [1, 2, 0], [1, 1, 1], [1, 0, 2]]
    assert [l.name for l, _ in v2.sys[6:9]] == ["∂[(0, 2)]x", "∂[(0, 1), (1, 1)]x", "∂[(1, 2)]x"]
    assert list(v2._initial_var_state(float)) == [1, 0, 0, 1, 0, 0, 0, 0, 0, 0]
    with pytest.raises(ValueError):
        v2.get_vslice(order=3)
    # w.r.t. the parameter as well: 3 arguments -> 2 * (3 + 6) sensitivities
    v3 = hy.var_ode_sys(sys_, hy.var_args.vars | hy.var_args.params, order=2)
    assert len(v3.sys) == 2 + 2 * 3 + 2 * 6 and len(v3.vargs) == 3
    assert v3.get_mindex(2 + 6) == [0, 2, 0, 0] and v3.get_mindex(2 + 6 + 5) == [0, 0, 0, 2]
    # the Taylor map is the truncated multivariate series
    st = np.arange(1.0, 13.0)[:, None]
    dx = np.array([[0.1], [-0.2]])
    tm = v2.eval_taylor_map(st, dx)
    ex0 = 1 + 3 * 0.1 + 4 * -0.2 + 7 * 0.01 / 2 + 8 * 0.1 * -0.2 + 9 * 0.04 / 2
    assert abs(tm[0, 0] - ex0) < 1e-15
    # symbolic Jacobian against finite differences
    rhs = [r for _, r in sys_]
    pt = {"x": 0.3, "v": -0.2}
    for i in range(2):
        for nm in ("x", "v"):
            d = E.eval_numpy(hy.diff(rhs[i], hy.expression(nm)), pt, pars=[0.1], tm=0.7)
            e = 1e-6
            p1 = dict(pt); p1[nm] += e
            p0 = dict(pt); p0[nm] -= e
            fd = (E.eval_numpy(rhs[i], p1, pars=[0.1], tm=0.7) - E.eval_numpy(rhs[i], p0, pars=[0.1], tm=0.7)) / (2 * e)
            assert abs(d - fd) < 1e-8


def test_ensemble_argument_validation():
    # /root/reference/heyoka/_test_ensemble.py error paths (raised before any device work)
    ta = object()
    with pytest.raises(TypeError, match="n_iter parameter must be an integer"):
        hy.ensemble_propagate_until_batch(ta, 20.0, "a", lambda t, i: t)
    with pytest.raises(ValueError, match="must be non-negative"):
        hy.ensemble_propagate_until_batch(ta, 20.0, -1, lambda t, i: t)
    with pytest.raises(TypeError, match="must be a scalar, not an iterable"):
        hy.ensemble_propagate_until_batch(ta, [20.0], 1, lambda t, i: t)
    with pytest.raises(ValueError, match="must be one-dimensional"):
        hy.ensemble_propagate_grid_batch(ta, [[1.0, 2.0]], 1, lambda t, i: t)
    with pytest.raises(TypeError, match='"max_delta_t"'):
        hy.ensemble_propagate_for_batch(ta, 1.0, 1, lambda t, i: t, max_delta_t=[1.0])
    with pytest.raises(ValueError, match="parallelisation algorithm"):
        hy.ensemble_propagate_for_batch(ta, 1.0, 1, lambda t, i: t, algorithm="foo")
    with pytest.raises(TypeError, match="chunksize"):
        hy.ensemble_propagate_for_batch(ta, 1.0, 1, lambda t, i: t, chunksize=3)
    assert hy.ensemble_propagate_for_batch(ta, 1.0, 0, lambda t, i: t) == []


def test_enums():
    assert int(hy.taylor_outcome.success) == -4294967297
    assert int(hy.taylor_outcome.time_limit) == -4294967299
    assert hy.taylor_outcome.success > hy.taylor_outcome.time_limit
    assert hy.event_direction.any == 0
