"""GPU: the elementary functions and model builders of SURVEY.md section 8 (f3 / f4) through the
C ABI - tape interpreter and run-time compiled kernel against the C oracle (same tape) and the
numpy oracle (expression DAG)."""

import os

import numpy as np
import pytest

import hy_b200 as hy
from hy_b200 import decompose as D
from hy_b200 import model
from oracle.c_oracle import COracle
from oracle.np_oracle import NpTaylorBatch

pytestmark = pytest.mark.gpu


def _jit(make):
    os.environ["HY_CUDA_JIT"] = "1"
    try:
        ta = make()
        ta._ctx
    finally:
        del os.environ["HY_CUDA_JIT"]
    return ta


def _check(sys_, ic, t_end, pars=None, tol=1e-12, fp=np.float64, bitwise=True):
    kw = {} if pars is None else {"pars": pars}
    a = hy.taylor_adaptive_batch(sys_, ic.astype(fp), fp_type=fp, compact_mode=True, **kw)
    b = _jit(lambda: hy.taylor_adaptive_batch(sys_, ic.astype(fp), fp_type=fp, **kw))
    assert a._ctx.launch_info()["kernel_variant"] == 0 and b._ctx.launch_info()["kernel_variant"] == 1000
    for ta in (a, b):
        ta.propagate_until(fp(t_end))
    orc = COracle(D.decompose(sys_, a.order), ic.astype(fp), fp_type=fp, **kw)
    oc, mn, mx, ns, _ = orc.propagate_until(fp(t_end))
    for ta in (a, b):
        assert list(ta.propagate_res_arrays[3]) == list(ns)
        err = np.max(np.abs(ta.state - orc.state) / np.maximum(1.0, np.abs(orc.state)))
        assert err < tol, err
    if bitwise:
        assert np.array_equal(a.state, b.state)   # interpreter and compiled kernel: bit for bit
    else:
        # (central-force pairs: the interpreter's fused pair op sums in another order)
        assert np.max(np.abs(a.state - b.state) / np.maximum(1.0, np.abs(b.state))) < tol
    return a


@pytest.mark.parametrize("fp,tol", [(np.float64, 1e-12), (np.float32, 1e-5)])
def test_asin_acos_atan_erf(fp, tol):
    x, y = hy.make_vars("x", "y")
    sys_ = [(x, hy.asin(0.5 * hy.sin(hy.time)) + hy.atan(y)),
            (y, hy.erf(x) - 0.1 * hy.acos(0.3 * hy.cos(y)))]
    B = 48
    ic = np.stack([np.linspace(-0.3, 0.3, B), np.linspace(0.0, 0.8, B)])
    ta = _check(sys_, ic, 3.0, tol=tol, fp=fp)
    if fp == np.float64:
        ref = NpTaylorBatch(sys_, ic)
        ref.propagate_until(3.0)
        assert np.max(np.abs(ta.state - ref.state)) < 1e-12


def test_hyperbolic_functions_and_sigmoid():
    x, y = hy.make_vars("x", "y")
    sys_ = [(x, hy.tanh(hy.time) - x * hy.sigmoid(y) + 0.1 * hy.asinh(x)),
            (y, hy.sinh(0.3 * x) - 0.2 * hy.cosh(0.5 * y) + 0.05 * hy.atanh(0.5 * hy.sin(x)))]
    B = 32
    ic = np.stack([np.linspace(-0.4, 0.4, B), np.linspace(0.1, 0.5, B)])
    _check(sys_, ic, 2.0)


def test_np1body_and_fixed_centres():
    m = [1.0, 1e-3, 3e-4, 5e-5]
    sys_ = model.np1body(4, masses=m, Gconst=1.0)
    B = 24
    base = np.array([1.0, 0, 0, 0, 1.0, 0.05, 0, 1.8, 0.1, -0.75, 0, 0, -2.6, 0.1, 0, 0, -0.62, 0.02])
    ic = base[:, None] * (1.0 + 1e-3 * np.linspace(-1, 1, B))[None, :]
    ta = _check(sys_, ic, 6.0, bitwise=False)
    en = model.np1body_energy(4, masses=m)
    from hy_b200 import _expression as E

    names = [l.name for l, _ in sys_]
    e0 = E.eval_numpy(en, {n: ic[i] for i, n in enumerate(names)})
    e1 = E.eval_numpy(en, {n: ta.state[i] for i, n in enumerate(names)})
    assert np.max(np.abs(e1 - e0) / np.abs(e0)) < 1e-13
    kw = dict(Gconst=1.0, masses=[1.0, 0.5], positions=[[-1.0, 0.0, 0.0], [1.0, 0.0, 0.2]])
    sys_ = model.fixed_centres(**kw)
    ic = np.array([0.1, 1.3, 0.2, 0.6, 0.0, 0.1])[:, None] * np.ones((1, B)) + 1e-3 * np.arange(B)[None, :]
    ta = _check(sys_, ic, 5.0, bitwise=False)
    en = model.fixed_centres_energy(**kw)
    names = [l.name for l, _ in sys_]
    e0 = E.eval_numpy(en, {n: ic[i] for i, n in enumerate(names)})
    e1 = E.eval_numpy(en, {n: ta.state[i] for i, n in enumerate(names)})
    assert np.max(np.abs(e1 - e0)) < 1e-13
