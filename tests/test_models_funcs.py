"""CPU: the model builders and elementary functions added for SURVEY.md section 8 (f3 / f4):
np1body, fixed_centres (+ energies / potentials) with the reference's structure checks
(/root/reference/heyoka/_test_model.py:101-208), hyperbolic functions / inverses / sigmoid as
compositions (expose_expression.cpp:288-306), integrated by the numpy oracle."""

import numpy as np
import pytest

import hy_b200 as hy
from hy_b200 import _expression as E
from hy_b200 import model
from oracle.np_oracle import NpTaylorBatch


def _vals(sys_, st):
    return {l.name: st[i] for i, (l, _) in enumerate(sys_)}


def test_np1body_structure_and_equivalence_with_nbody():
    dyn = model.np1body(2, masses=[0.0, 0.0])
    assert len(dyn) == 6
    assert dyn[3][1] == hy.expression(0.0) and dyn[5][1] == hy.expression(0.0)
    dyn = model.np1body(2, Gconst=5.0)
    assert all("10.0000000000000" in str(dyn[i][1]) for i in (3, 4, 5))
    assert model.np1body_energy(2, masses=[]) == hy.expression(0.0)
    assert "5.0000000000000" in str(model.np1body_energy(2, Gconst=5.0))
    assert [l.name for l, _ in dyn] == ["x_1", "y_1", "z_1", "vx_1", "vy_1", "vz_1"]
    # the relative dynamics of a 3-body system = the 3-body system itself, body 0 subtracted
    m = [1.0, 1e-3, 3e-4]
    ic = np.array([0, 0, 0, 0, 0, 0, 1.0, 0, 0, 0, 1.0, 0.1, 0, 2.0, 0.1, -0.7, 0, 0], dtype=float)[:, None]
    a = NpTaylorBatch(model.nbody(3, masses=m), ic)
    a.propagate_until(3.0)
    rel = ic[6:] - np.tile(ic[:6], (2, 1))
    sys_r = model.np1body(3, masses=m)
    b = NpTaylorBatch(sys_r, rel)
    b.propagate_until(3.0)
    assert np.max(np.abs((a.state[6:] - np.tile(a.state[:6], (2, 1))) - b.state)) < 1e-13
    en = model.np1body_energy(3, masses=m)
    e0 = E.eval_numpy(en, _vals(sys_r, rel[:, 0]))
    e1 = E.eval_numpy(en, _vals(sys_r, b.state[:, 0]))
    assert abs(e1 - e0) < 1e-15
    pot = E.eval_numpy(model.np1body_potential(3, masses=m), _vals(sys_r, rel[:, 0]))
    assert pot < 0 and pot < e0


def test_fixed_centres_structure_errors_and_energy():
    x = hy.make_vars("x")
    dyn = model.fixed_centres(Gconst=1.5, masses=[1.1], positions=[[1.0, 2.0, 3.0]])
    assert dyn[0][0] == x and dyn[0][1] == hy.expression("vx") and len(dyn) == 6
    model.fixed_centres_energy(Gconst=1.5, masses=[1.1], positions=[[1.0, 2.0, 3.0]])
    model.fixed_centres_potential(Gconst=1.5, masses=[1.1], positions=[[1.0, 2.0, 3.0]])
    with pytest.raises(ValueError, match="the number of dimensions must be 2, but it is 1 instead"):
        model.fixed_centres(Gconst=1.5, masses=[1.1], positions=[1.0, 2.0, 3.0])
    with pytest.raises(ValueError, match="the number of columns must be 3, but it is 4 instead"):
        model.fixed_centres(Gconst=1.5, masses=[1.1], positions=[[1.0, 2.0, 3.0, 4.0]])
    with pytest.raises(TypeError, match="could not be converted into an array of expressions"):
        model.fixed_centres(Gconst=1.5, masses=[1.1], positions=[[{}, {}, {}]])
    model.fixed_centres(Gconst=np.single(1.5), masses=[np.single(1.1)], positions=[[1.0, 2.0, 3.0]])
    # two centres: the energy is conserved along the trajectory
    kw = dict(Gconst=1.0, masses=[1.0, 0.5], positions=[[-1.0, 0.0, 0.0], [1.0, 0.0, 0.2]])
    sys_ = model.fixed_centres(**kw)
    ic = np.array([0.1, 1.3, 0.2, 0.6, 0.0, 0.1])[:, None]
    ta = NpTaylorBatch(sys_, ic)
    ta.propagate_until(5.0)
    en = model.fixed_centres_energy(**kw)
    assert abs(E.eval_numpy(en, _vals(sys_, ta.state[:, 0])) - E.eval_numpy(en, _vals(sys_, ic[:, 0]))) < 1e-13


def test_hyperbolic_functions_and_sigmoid():
    x = hy.make_vars("x")
    for f, g, pt in ((hy.sinh, np.sinh, 0.3), (hy.cosh, np.cosh, 0.3), (hy.tanh, np.tanh, 0.3),
                     (hy.asinh, np.arcsinh, 0.3), (hy.acosh, np.arccosh, 1.7), (hy.atanh, np.arctanh, 0.3),
                     (hy.sigmoid, lambda t: 1.0 / (1.0 + np.exp(-t)), 0.3)):
        assert abs(E.eval_numpy(f(x), {"x": pt}) - g(pt)) < 1e-15
        # derivative against central differences
        d = E.eval_numpy(hy.diff(f(x), x), {"x": pt})
        fd = (g(pt + 1e-6) - g(pt - 1e-6)) / 2e-6
        assert abs(d - fd) < 1e-8
        assert f(pt).kind == "num"  # numbers fold
    # x' = tanh(t) - x sigmoid(x): integrate and compare with a fine RK4
    sys_ = [(x, hy.tanh(hy.time) - x * hy.sigmoid(x) + 0.1 * hy.asinh(x))]
    ta = NpTaylorBatch(sys_, np.array([[0.4]]))
    ta.propagate_until(2.0)

    def rhs(t, y):
        return np.tanh(t) - y / (1.0 + np.exp(-y)) + 0.1 * np.arcsinh(y)

    y, t, h = 0.4, 0.0, 1e-4
    for _ in range(20000):
        k1 = rhs(t, y); k2 = rhs(t + h / 2, y + h / 2 * k1); k3 = rhs(t + h / 2, y + h / 2 * k2); k4 = rhs(t + h, y + h * k3)
        y += h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
        t += h
    assert abs(ta.state[0, 0] - y) < 1e-12


def _intg_sys():
    x, y = hy.make_vars("x", "y")
    return [(x, hy.asin(0.5 * hy.sin(hy.time)) + hy.atan(y)),
            (y, hy.erf(x) - 0.1 * hy.acos(0.3 * hy.cos(y)))]


def test_asin_acos_atan_erf_tape_and_oracles():
    # One new recurrence (HY_OP_INTG, include/hy_cuda.h) serves the four functions: F(a) with
    # dF/da = g(a), g built from existing ops.  Reference: expose_expression.cpp:288-306.
    import math
    from hy_b200 import decompose as D
    from oracle.c_oracle import COracle

    x = hy.make_vars("x")
    for f, g, pt in ((hy.asin, np.arcsin, 0.3), (hy.acos, np.arccos, 0.3), (hy.atan, np.arctan, 1.3),
                     (hy.erf, math.erf, 0.4)):
        assert abs(E.eval_numpy(f(x), {"x": pt}) - g(pt)) < 1e-15
        d = E.eval_numpy(hy.diff(f(x), x), {"x": pt})
        assert abs(d - (g(pt + 1e-6) - g(pt - 1e-6)) / 2e-6) < 1e-8
        assert f(pt).kind == "num"
    sys_ = _intg_sys()
    dc = D.decompose(sys_, 20)
    kinds = [D.OP_NAMES[int(o["opcode"])] for o in dc.ops]
    assert kinds.count("intg") == 4
    ic = np.array([[0.1, -0.2], [0.3, 0.5]])
    a = NpTaylorBatch(sys_, ic)
    a.propagate_until(3.0)
    b = COracle(dc, ic)
    b.propagate_until(3.0)
    assert np.max(np.abs(a.state - b.state)) < 1e-13      # DAG walker vs tape interpreter

    def rhs(t, s):
        return np.array([math.asin(0.5 * math.sin(t)) + math.atan(s[1]),
                         math.erf(s[0]) - 0.1 * math.acos(0.3 * math.cos(s[1]))])

    s, t, h = ic[:, 0].copy(), 0.0, 2e-4
    for _ in range(15000):
        k1 = rhs(t, s); k2 = rhs(t + h / 2, s + h / 2 * k1); k3 = rhs(t + h / 2, s + h / 2 * k2); k4 = rhs(t + h, s + h * k3)
        s = s + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
        t += h
    assert np.max(np.abs(s - a.state[:, 0])) < 1e-11
