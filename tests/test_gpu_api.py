"""GPU: behaviour of the drop-in class, mirroring the reference's own tests:
/root/reference/heyoka/_test_batch_integrator.py (ctor, copy, propagate_for/until,
time/dtime, pickling, callbacks), _test_ensemble.py (ensemble == serial,
bit-exact) and _test_var_integrator.py::test_batch."""

import pickle
from copy import copy, deepcopy

import numpy as np
import pytest

import hy_b200 as hy

import common

pytestmark = pytest.mark.gpu


def _ta(fp=np.float64, **kw):
    return hy.taylor_adaptive_batch(common.pendulum_sys(), common.PEND_IC.astype(fp), fp_type=fp, **kw)


def test_ctor_and_properties():
    # _test_batch_integrator.py:13-88
    ta = _ta()
    assert ta.batch_size == 4 and ta.dim == 2 and ta.order == 20
    assert ta.tol == np.finfo(float).eps and not ta.high_accuracy and not ta.compact_mode
    assert not ta.with_events and not ta.is_variational
    assert np.all(ta.state == common.PEND_IC) and np.all(ta.time == 0)
    assert ta.sys[0][1] == hy.make_vars("v")
    assert len(ta.decomposition) > 0 and "libhy_cuda" in repr(ta)
    assert ta.llvm_state.opt_level == 3
    ta2 = hy.taylor_adaptive_batch(common.pendulum_sys(), common.PEND_IC, tol=1e-9,
                                   high_accuracy=True, compact_mode=True, time=[1.0, 2.0, 3.0, 4.0],
                                   opt_level=2, fast_math=False)
    assert ta2.order == 12 and ta2.high_accuracy and ta2.compact_mode
    assert np.all(ta2.time == [1, 2, 3, 4])
    with pytest.raises(ValueError, match="number of dimensions is 2"):
        hy.taylor_adaptive_batch(common.pendulum_sys(), [0.0, 1.0])
    with pytest.raises(ValueError, match="Invalid parameter vector"):
        hy.taylor_adaptive_batch(common.forced_pendulum_sys(), common.PEND_IC, pars=[[1.0, 2.0]])
    with pytest.raises(ValueError, match="Invalid time vector"):
        hy.taylor_adaptive_batch(common.pendulum_sys(), common.PEND_IC, time=[0.0])
    with pytest.raises(TypeError):
        hy.taylor_adaptive_batch(common.pendulum_sys(), common.PEND_IC.astype(np.longdouble))
    with pytest.raises(TypeError):
        hy.taylor_adaptive_batch(common.pendulum_sys(), common.PEND_IC, fp_type=np.float32, tol=1e-3)
    with pytest.raises(TypeError):
        hy.taylor_adaptive_batch(common.pendulum_sys(), common.PEND_IC, fp_type=np.longdouble)


@pytest.mark.parametrize("fp", [np.float64, np.float32])
def test_propagate_scalar_vs_vector_identical(fp):
    # _test_batch_integrator.py:128-170: bit-identical state and propagate_res
    ic = common.PEND_IC.astype(fp)
    ta = _ta(fp)
    ta.propagate_for([fp(10.0)] * 4)
    st, res = deepcopy(ta.state), deepcopy(ta.propagate_res)
    ta.set_time(fp(0.0))
    ta.state[:] = ic
    ta.propagate_for(fp(10.0))
    assert np.all(ta.state == st) and res == ta.propagate_res
    ta.set_time(fp(0.0))
    ta.state[:] = ic
    ta.propagate_for([fp(10.0)] * 4, max_delta_t=[fp(1e-2)] * 4)
    st, res = deepcopy(ta.state), deepcopy(ta.propagate_res)
    ta.set_time(fp(0.0))
    ta.state[:] = ic
    ta.propagate_for(fp(10.0), max_delta_t=fp(1e-2))
    assert np.all(ta.state == st) and res == ta.propagate_res
    assert all(r[3] >= 1000 for r in res)
    ta.set_time(fp(0.0))
    ta.state[:] = ic
    ta.propagate_until(fp(10.0), max_steps=5)
    assert all(r[0] == hy.taylor_outcome.step_limit and r[3] == 5 for r in ta.propagate_res)
    if fp == np.float32:
        with pytest.raises(TypeError):
            ta.propagate_for(10.0)


def test_step_callbacks():
    # _test_batch_integrator.py:170-260, The adaptive integrator.ipynb cells 19-21
    ta = _ta()

    def cb(ta):
        ta.counter = getattr(ta, "counter", 0) + 1
        return True

    ta.propagate_for(10.0, callback=cb)
    ref = _ta()
    ref.propagate_for(10.0)
    assert ta.counter == max(r[3] for r in ref.propagate_res)
    assert [r[3] for r in ta.propagate_res] == [r[3] for r in ref.propagate_res]
    assert np.max(np.abs(ta.state - ref.state)) < 1e-14

    class CB:
        def __init__(self):
            self.pre = 0

        def __call__(_, ta):
            assert id(_) == _.orig_id
            return True

        def pre_hook(self, ta):
            self.pre += 1

    c = CB()
    c.orig_id = id(c)
    ret = ta.propagate_for(1.0, callback=c)
    assert ret[1] is c and c.pre == 1 and ret[0] is None
    ret = ta.propagate_for(1.0, callback=[c, cb])
    assert ret[1][0] is c and ret[1][1] is cb
    with pytest.raises(TypeError, match="not callable"):
        ta.propagate_for(10.0, callback="hello world")
    with pytest.raises(TypeError, match="expected to return a boolean"):
        ta.propagate_for(10.0, callback=lambda ta: "hello")
    # stopping callback -> cb_stop
    ta = _ta()
    ta.propagate_until(10.0, callback=lambda ta: False)
    assert all(r[0] == hy.taylor_outcome.cb_stop and r[3] == 1 for r in ta.propagate_res)
    # max_delta_t with callback: The adaptive integrator.ipynb cell 19
    ta = hy.taylor_adaptive_batch(common.pendulum_sys(), np.array([[0.05] * 2, [0.025] * 2]))
    seen = []
    ta.propagate_until(0.5, max_delta_t=0.1, callback=lambda t: seen.append(float(t.time[0])) or True)
    assert np.allclose(seen, [0.1, 0.2, 0.30000000000000004, 0.4, 0.5], rtol=0, atol=1e-15)
    # angle reducer
    x, v = hy.make_vars("x", "v")
    ta = hy.taylor_adaptive_batch(common.pendulum_sys(), np.array([[0.05] * 2, [5.0] * 2]))
    ta.propagate_until(10.0, callback=hy.callback.angle_reducer([x]))
    assert np.all((ta.state[0] >= 0) & (ta.state[0] < 2 * np.pi))


def test_time_and_dtime():
    # _test_batch_integrator.py:415-488
    ta = _ta()
    assert not ta.time.flags.writeable
    ta.set_time(1.5)
    assert np.all(ta.time == 1.5)
    ta.set_time([1.0, 2.0, 3.0, 4.0])
    assert np.all(ta.time == [1, 2, 3, 4])
    with pytest.raises(ValueError):
        ta.set_time([1.0, 2.0])
    ta.set_time(0.0)
    ta.propagate_until(1000.1)
    hi, lo = ta.dtime
    assert np.all(hi == 1000.1) and not hi.flags.writeable
    ta.propagate_for(0.1)
    assert np.any(ta.dtime[1] != 0)
    ta.set_dtime(1.0, 0.5)
    assert np.all(ta.dtime[0] == 1.5) and np.all(ta.dtime[1] == 0.0)
    ta.set_dtime([1.0] * 4, [0.5] * 4)
    assert np.all(ta.dtime[0] == 1.5)
    with pytest.raises(TypeError):
        ta.set_dtime(1.0, [0.5] * 4)


def test_copy_deepcopy_pickle():
    # _test_batch_integrator.py:89-126, :602-690
    ta = _ta()
    ta.foo = [1, 2, 3]
    ta.propagate_until(1.0)
    c1, c2 = copy(ta), deepcopy(ta)
    assert c1.foo is ta.foo and c2.foo == ta.foo and c2.foo is not ta.foo
    assert np.all(c2.state == ta.state) and np.all(c2.time == ta.time)
    c2.state[:] = 0
    assert not np.all(ta.state == 0)
    p = pickle.loads(pickle.dumps(ta))
    assert p.foo == ta.foo and p.propagate_res == ta.propagate_res
    ta.step()
    p.step()
    assert np.all(p.state == ta.state) and np.all(p.time == ta.time)
    assert p.step_res == ta.step_res


def test_ensemble_equals_serial_bit_exact():
    # _test_ensemble.py:13-111
    ta = hy.taylor_adaptive_batch(common.pendulum_sys(), np.zeros((2, 4)))
    rng = np.random.default_rng(3)
    ics = rng.uniform(-0.3, 0.3, (10, 2, 4))

    def gen(t, i):
        t.state[:] = ics[i]
        return t

    for algo in ("thread", "process"):
        kw = {"algorithm": algo}
        if algo == "thread":
            kw["max_workers"] = 4
        ret = hy.ensemble_propagate_until_batch(ta, 20.0, 10, gen, **kw)
        assert len(ret) == 10
        for i in range(10):
            ser = gen(deepcopy(ta), i)
            ser.propagate_until(20.0)
            assert np.all(ret[i][0].state == ser.state)
            assert np.all(ret[i][0].time == ser.time)
            assert ret[i][0].propagate_res == ser.propagate_res
            assert ret[i][1] is None and ret[i][2] is None
    ret = hy.ensemble_propagate_for_batch(ta, 5.0, 3, gen, c_output=True)
    for i in range(3):
        ser = gen(deepcopy(ta), i)
        co, _ = ser.propagate_for(5.0, c_output=True)
        assert np.all(ret[i][1](2.5) == co(2.5))
    grid = np.linspace(0.0, 3.0, 7)
    ret = hy.ensemble_propagate_grid_batch(ta, grid, 3, gen)
    for i in range(3):
        ser = gen(deepcopy(ta), i)
        _, out = ser.propagate_grid(np.repeat(grid, 4).reshape(-1, 4))
        assert np.all(ret[i][2] == out)
    # callbacks are deep-copied per iteration
    class CB:
        def __call__(self, ta):
            return True

    cb = CB()
    ret = hy.ensemble_propagate_until_batch(ta, 1.0, 3, gen, callback=cb)
    assert all(r[2] is not cb and isinstance(r[2], CB) for r in ret)


def test_lane_results_independent_of_batch_composition():
    # determinism: a trajectory gives bit-identical results alone or inside a big batch
    sys_ = common.oss_sys()
    ic = common.oss_ensemble(64)
    big = hy.taylor_adaptive_batch(sys_, ic)
    big.propagate_until(20.0)
    small = hy.taylor_adaptive_batch(sys_, ic[:, 5:9].copy())
    small.propagate_until(20.0)
    assert np.all(big.state[:, 5:9] == small.state)
    assert big.propagate_res[5:9] == small.propagate_res


def test_variational_batch():
    # _test_var_integrator.py:144-258 (order 1)
    x, v = hy.make_vars("x", "v")
    sys_ = [(x, v), (v, hy.cos(hy.time) - hy.par[0] * v - hy.sin(x))]
    vs = hy.var_ode_sys(sys_, hy.var_args.vars)
    B = 4
    ic = np.array([[0.2 + 0.01 * i for i in range(B)], [0.3] * B])
    ta = hy.taylor_adaptive_batch(vs, ic, pars=np.full((1, B), 0.4))
    assert ta.is_variational and ta.dim == 6 and ta.n_orig_sv == 2 and ta.vorder == 1
    assert np.all(ta.state[2:, 0] == [1, 0, 0, 1])  # identity ICs auto-filled
    assert ta.get_vslice(order=1) == slice(2, 6) and ta.get_mindex(3) == [0, 0, 1]
    ta.propagate_until(3.0)
    # sensitivities against finite differences of the plain system
    e = 1e-6
    for j in range(2):
        d = np.zeros((2, B))
        d[j] = e
        tp = hy.taylor_adaptive_batch(sys_, ic + d, pars=np.full((1, B), 0.4))
        tm = hy.taylor_adaptive_batch(sys_, ic - d, pars=np.full((1, B), 0.4))
        tp.propagate_until(3.0)
        tm.propagate_until(3.0)
        fd = (tp.state - tm.state) / (2 * e)
        sens = ta.state[2:].reshape(2, 2, B)[:, j, :]
        assert np.max(np.abs(fd - sens)) < 1e-8
    ts = ta.eval_taylor_map(np.zeros((2, B)))
    assert np.all(ts == ta.state[:2]) and not ts.flags.writeable
    with pytest.raises(ValueError):
        ta.eval_taylor_map(np.zeros((3, B)))
    with pytest.raises(ValueError):
        _ta().vorder


def test_high_accuracy_and_backward():
    ta = hy.taylor_adaptive_batch(common.pendulum_sys(), common.PEND_IC, high_accuracy=True, tol=1e-18)
    assert ta.order == 22
    ta.propagate_until(50.0)
    ta.propagate_until(0.0)
    assert np.max(np.abs(ta.state - common.PEND_IC)) < 1e-13
    ta = _ta()
    ta.step_backward()
    assert all(r[1] < 0 for r in ta.step_res)
    ta.step([1e-3] * 4)
    assert all(r[0] == hy.taylor_outcome.time_limit and r[1] == 1e-3 for r in ta.step_res)
    # non-finite state -> err_nf_state, not an exception
    x = hy.make_vars("x")
    ta = hy.taylor_adaptive_batch([(x, x * x)], np.array([[1.0, 1.0]]))
    ta.propagate_until(2.0)
    assert all(r[0] in (hy.taylor_outcome.err_nf_state, hy.taylor_outcome.time_limit)
               for r in ta.propagate_res)


@pytest.mark.parametrize("fp", [np.float32, np.float64])
def test_propagate_grid_reference_scenarios(fp):
    # /root/reference/heyoka/_test_batch_integrator.py:692-857, scenario by scenario (the scalar integrator of the
    # comparison there is replaced by one-lane batch integrators)
    from copy import deepcopy

    x, v = hy.make_vars("x", "v")
    eqns = [(x, v), (v, -9.8 * hy.sin(x))]
    x_ic = np.array([0.06, 0.07, 0.08, 0.09], dtype=fp)
    v_ic = np.array([0.025, 0.026, 0.027, 0.028], dtype=fp)
    ta = hy.taylor_adaptive_batch(eqns, [x_ic, v_ic], fp_type=fp)

    with pytest.raises(ValueError) as cm:
        ta.propagate_grid(np.array([], dtype=fp))
    assert ("Invalid grid passed to the propagate_grid() method of a batch integrator: the expected number of "
            "dimensions is 2, but the input array has a dimension of 1") in str(cm.value)
    with pytest.raises(ValueError) as cm:
        ta.propagate_grid(np.array([[1, 2], [3, 4]], dtype=fp))
    assert ("Invalid grid passed to the propagate_grid() method of a batch integrator: the shape must be (n, 4) "
            "but the number of columns is 2 instead") in str(cm.value)

    grid = np.array([[-0.1, -0.2, -0.3, -0.4], [0.01, 0.02, 0.03, 0.9], [1.0, 1.1, 1.2, 1.3],
                     [11.0, 11.1, 11.2, 11.3]], dtype=fp)
    ta.propagate_until(grid[0, :])
    bres = ta.propagate_grid(grid)
    assert bres[0] is None and bres[1].shape == (4, 2, 4)
    for idx in range(4):
        one = hy.taylor_adaptive_batch(eqns, [x_ic[idx:idx + 1], v_ic[idx:idx + 1]], fp_type=fp)
        one.propagate_until(grid[0, idx:idx + 1])
        sres = one.propagate_grid(grid[:, idx:idx + 1].copy())
        assert np.max(np.abs(sres[1][:, :, 0] - bres[1][:, :, idx])) < np.finfo(fp).eps * 100

    # vector / scalar max_delta_t
    def restart():
        ta.set_time(fp(0.0))
        ta.state[:] = [x_ic, v_ic]
        ta.propagate_until(grid[0, :])

    restart()
    b1 = ta.propagate_grid(grid, max_delta_t=[fp(1e-3)] * 4)
    res = deepcopy(ta.propagate_res)
    restart()
    b2 = ta.propagate_grid(grid, max_delta_t=fp(1e-3))
    assert np.all(b1[1] == b2[1]) and ta.propagate_res == res

    # dynamic attributes through the callback; no copies of the callback
    def cb(t):
        t.counter = t.counter + 1 if hasattr(t, "counter") else 0
        return True

    restart()
    ta.propagate_grid(grid, callback=cb)
    assert ta.counter > 0

    class cb_id:
        def __call__(self_, t):
            assert id(self_) == self_.orig_id
            return True

    inst = cb_id()
    inst.orig_id = id(inst)
    restart()
    ta.propagate_grid(grid, callback=inst)

    with pytest.raises(TypeError) as cm:
        restart()
        ta.propagate_grid(grid, callback="hello world")
    assert "cannot be used as a step callback because it is not callable" in str(cm.value)

    class broken_cb:
        def __call__(self_, t):
            return []

    with pytest.raises(TypeError) as cm:
        restart()
        ta.propagate_grid(grid, callback=broken_cb())
    assert "The call operator of a step callback is expected to return a boolean, but a value of type" in str(cm.value)

    class cb_hook:
        def __call__(self_, t):
            return True

        def pre_hook(self_, t):
            t.foo = True

    restart()
    ta.propagate_grid(grid, callback=cb_hook())
    assert ta.foo


@pytest.mark.parametrize("fp", [np.float32, np.float64])
def test_reference_scenarios_d_output_time_dtime_basic(fp):
    # /root/reference/heyoka/_test_batch_integrator.py:340-553 (test_update_d_output, test_set_time, test_dtime,
    # test_basic), scenario by scenario
    from copy import deepcopy
    from sys import getrefcount

    x, v = hy.make_vars("x", "v")
    sys_ = [(x, v), (v, -9.8 * hy.sin(x))]

    # ---- update_d_output (:340-413)
    ta = hy.taylor_adaptive_batch(sys=sys_, state=np.array([[0.0, 0.1, 0.2, 0.3], [0.25, 0.26, 0.27, 0.28]], dtype=fp),
                                  fp_type=fp)
    ta.step(write_tc=True)
    with pytest.raises(ValueError):
        ta.update_d_output(fp(0.3))[0] = fp(0.5)
    d_out = ta.update_d_output(fp(0.3))
    assert d_out.shape == (2, 4)
    rc = getrefcount(ta)
    tmp_out = ta.update_d_output(fp(0.2))
    assert getrefcount(ta) == rc + 1
    with pytest.raises(ValueError):
        ta.update_d_output(np.array([0.3, 0.4, 0.45, 0.46], dtype=fp))[0] = fp(0.5)
    d_out2 = ta.update_d_output(np.array([0.3, 0.4, 0.45, 0.46], dtype=fp))
    assert d_out2.shape == (2, 4)
    rc = getrefcount(ta)
    tmp_out2 = ta.update_d_output(np.array([0.31, 0.41, 0.66, 0.67], dtype=fp))
    assert getrefcount(ta) == rc + 1
    cp = deepcopy(ta.update_d_output(fp(0.3)))
    assert np.all(cp == ta.update_d_output([fp(0.3)] * 4))
    ta.set_time(fp(0.0))
    ta.state[:] = [[0.0, 0.01, 0.02, 0.03], [0.205, 0.206, 0.207, 0.208]]
    ta.step(write_tc=True)
    ta.update_d_output(ta.time)
    eps10 = np.finfo(fp).eps * 10
    assert np.allclose(ta.d_output, ta.state, rtol=eps10, atol=eps10)
    ta.update_d_output(fp(0.0), rel_time=True)
    assert np.allclose(ta.d_output, ta.state, rtol=eps10, atol=eps10)
    del tmp_out, tmp_out2

    # ---- set_time (:415-438), dtime (:440-488)
    ta = hy.taylor_adaptive_batch(sys=sys_, state=np.array([[0.0, 0.1], [0.25, 0.26]], dtype=fp), fp_type=fp)
    assert np.all(ta.time == [0, 0])
    ta.set_time([fp(-1.0), fp(1.0)])
    assert np.all(ta.time == [-1, 1])
    ta.set_time(fp(5.0))
    assert np.all(ta.time == [5, 5])
    ta = hy.taylor_adaptive_batch(sys=sys_, state=np.array([[0.0, 0.1], [0.25, 0.26]], dtype=fp), fp_type=fp)
    assert np.all(ta.dtime[0] == [0, 0]) and np.all(ta.dtime[1] == [0, 0])
    with pytest.raises(ValueError):
        ta.dtime[0][0] = 0.5
    with pytest.raises(ValueError):
        ta.dtime[1][0] = 0.5
    ta.step()
    ta.propagate_for(fp(1000.1))
    assert not np.all(ta.dtime[1] == [0, 0])
    ta.set_dtime(fp(1.0), fp(0.5))
    assert np.all(ta.dtime[0] == [1.5, 1.5]) and np.all(ta.dtime[1] == [0, 0])
    ta.set_dtime([fp(1.0), fp(2.0)], [fp(0.5), fp(0.25)])
    assert np.all(ta.dtime[0] == [1.5, 2.25]) and np.all(ta.dtime[1] == [0, 0])
    with pytest.raises(TypeError) as cm:
        ta.set_dtime([fp(1.0), fp(2.0)], fp(0.5))
    assert "The two arguments to the set_dtime() method must be of the same type" in str(cm.value)

    # ---- basic (:490-553)
    ta = hy.taylor_adaptive_batch(sys=sys_, state=np.array([[0.0, 0.1], [0.25, 0.26]], dtype=fp),
                                  t_events=[hy.t_event_batch(v, fp_type=fp)], fp_type=fp)
    assert ta.with_events and not ta.compact_mode and not ta.high_accuracy and ta.sys == sys_
    ta = hy.taylor_adaptive_batch(sys=sys_, state=np.array([[0.0, 0.1], [0.25, 0.26]], dtype=fp), compact_mode=True,
                                  high_accuracy=True, fp_type=fp)
    assert not ta.with_events and ta.compact_mode and ta.high_accuracy
    assert not ta.llvm_state.fast_math and not ta.llvm_state.force_avx512 and ta.llvm_state.opt_level == 3
    ta = hy.taylor_adaptive_batch(sys=sys_, state=np.empty((0, 2), dtype=fp), compact_mode=True, high_accuracy=True,
                                  fp_type=fp)
    assert np.all(ta.state == np.zeros((2, 2), dtype=fp))
    ta = hy.taylor_adaptive_batch(sys=sys_, state=np.array([[0.0, 0.1], [0.25, 0.26]], dtype=fp), compact_mode=True,
                                  high_accuracy=True, force_avx512=True, fast_math=True, opt_level=0, fp_type=fp)
    assert ta.llvm_state.fast_math and ta.llvm_state.force_avx512 and ta.llvm_state.opt_level == 0


def test_views_keep_the_integrator_alive():
    # _test_batch_integrator.py:13-50 (llvm_state reference counting) and the view semantics of
    # expose_batch_integrators.cpp:394-518: every live view holds the integrator, whose page-locked buffers it aliases
    import gc
    from sys import getrefcount

    x, v = hy.make_vars("x", "v")
    sys_ = [(x, v), (v, -9.8 * hy.sin(x))]
    ta = hy.taylor_adaptive_batch(sys_, [[0.0, 0.0], [0.0, 0.0]])
    rc = getrefcount(ta)
    tmp = ta.llvm_state
    assert getrefcount(ta) == rc + 1
    assert not ta.llvm_state.force_avx512 and not ta.llvm_state.slp_vectorize
    tb = hy.taylor_adaptive_batch(sys_, [[0.0, 0.0], [0.0, 0.0]], force_avx512=True, slp_vectorize=True, parjit=True,
                                  compact_mode=True, code_model=hy.code_model.large)
    rc = getrefcount(tb)
    tmp = tb.llvm_state
    assert getrefcount(tb) == rc + 1
    assert tb.llvm_state.force_avx512 and tb.llvm_state.slp_vectorize and tb.llvm_state.code_model == hy.code_model.large
    # views outlive `del ta`
    tc_ = hy.taylor_adaptive_batch(sys_, [[0.1, 0.2], [0.3, 0.4]])
    tc_.step(write_tc=True)
    st, tm, lh, tcs = tc_.state, tc_.time, tc_.last_h, tc_.tc
    keep = [np.array(a) for a in (st, tm, lh, tcs)]
    st[0, 0] = 5.0                                   # state is writable and aliases the integrator's buffer
    assert tc_.state[0, 0] == 5.0
    keep[0][0, 0] = 5.0
    with pytest.raises(ValueError):
        tm[0] = 1.0
    del tc_
    gc.collect()
    junk = [hy.taylor_adaptive_batch(sys_, [[1.0, 2.0], [3.0, 4.0]]) for _ in range(4)]   # would reuse freed buffers
    for a, k in zip((st, tm, lh, tcs), keep):
        assert np.array_equal(a, k)
    del junk
