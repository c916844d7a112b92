"""GPU: the scenarios of the reference's own batch-integrator tests, one by one
(/root/reference/heyoka/_test_batch_integrator.py; the line ranges are cited per test), run against
`hy_b200` - same calls, same expectations, same error messages."""

import pickle
from copy import copy, deepcopy

import numpy as np
import pytest

import hy_b200 as hy

pytestmark = pytest.mark.gpu

FP = [np.float32, np.float64]


def _pend():
    x, v = hy.make_vars("x", "v")
    return x, v, [(x, v), (v, -9.8 * hy.sin(x))]


def test_type_conversions():
    # :51-87
    x, v, sys_ = _pend()
    ta = hy.taylor_adaptive_batch(sys=sys_, state=((0.0, 0.1), (0.25, 0.26)), tol=1e-4)
    assert np.all(ta.state == ((0.0, 0.1), (0.25, 0.26)))
    ta = hy.taylor_adaptive_batch(sys=sys_, state=np.array([[0.0, 0.1], [0.25, 0.26]]), tol=1e-4)
    assert np.all(ta.state == ((0.0, 0.1), (0.25, 0.26)))
    if np.finfo(np.double).nmant == np.finfo(np.longdouble).nmant:
        return
    ld = np.longdouble
    with pytest.raises(TypeError):
        hy.taylor_adaptive_batch(sys=sys_, state=((ld(0.0), ld(0.1)), (ld(0.25), ld(0.26))), tol=1e-4)
    with pytest.raises(TypeError):
        hy.taylor_adaptive_batch(sys=sys_, state=np.array([[0.0, 0.1], [0.25, 0.26]], dtype=ld), tol=1e-4)


def test_copy():
    # :89-126
    x, v, sys_ = _pend()

    def cb0(ta, t, d_sgn, bidx):
        pass

    ta = hy.taylor_adaptive_batch(sys=sys_, state=[[0, 0.01], [0.25, 0.26]],
                                  nt_events=[hy.nt_event_batch(v * v - 1e-10, cb0)])
    for _ in range(4):
        ta.step()

    class foo:
        pass

    ta.bar = foo()
    assert id(ta.bar) == id(copy(ta).bar)
    assert id(ta.bar) != id(deepcopy(ta).bar)
    assert np.all(ta.state == copy(ta).state)
    assert np.all(ta.state == deepcopy(ta).state)
    ta_dc = deepcopy(ta)
    assert ta_dc.state[0, 0] == ta.state[0, 0]
    ta.state[0, 0] += 1
    assert ta_dc.state[0, 0] != ta.state[0, 0]


@pytest.mark.parametrize("fp", FP)
@pytest.mark.parametrize("which", ["for", "until"])
def test_propagate_for_until(fp, which):
    # :128-338 (propagate_for and propagate_until run the same scenarios)
    ic = np.array([[0.0, 0.1, 0.2, 0.3], [0.25, 0.26, 0.27, 0.28]], dtype=fp)
    x, v, sys_ = _pend()
    ta = hy.taylor_adaptive_batch(sys=sys_, state=ic, fp_type=fp)
    prop = lambda *a, **k: getattr(ta, "propagate_" + which)(*a, **k)

    def restart():
        ta.set_time(fp(0.0))
        ta.state[:] = ic

    prop([fp(10.0)] * 4)
    st, res = deepcopy(ta.state), deepcopy(ta.propagate_res)
    restart()
    prop(fp(10.0))
    assert np.all(ta.state == st) and res == ta.propagate_res
    restart()
    prop([fp(10.0)] * 4, max_delta_t=[fp(1e-4)] * 4)
    st, res = deepcopy(ta.state), deepcopy(ta.propagate_res)
    restart()
    prop(fp(10.0), max_delta_t=fp(1e-4))
    assert np.all(ta.state == st) and res == ta.propagate_res

    def cb(t):
        t.counter = t.counter + 1 if hasattr(t, "counter") else 0
        return True

    tgt = fp(10.0) if which == "for" else fp(20.0)
    prop(tgt, callback=cb)
    assert ta.counter > 0

    class cb_id:
        def __call__(self_, t):
            assert id(self_) == self_.orig_id
            return True

    inst = cb_id()
    inst.orig_id = id(inst)
    prop(fp(10.0) if which == "for" else fp(30.0), callback=inst)
    with pytest.raises(TypeError) as cm:
        restart()
        prop(fp(10.0), callback="hello world")
    assert "cannot be used as a step callback because it is not callable" in str(cm.value)

    class broken_cb:
        def __call__(self_, t):
            return []

    with pytest.raises(TypeError) as cm:
        restart()
        prop(fp(10.0), callback=broken_cb())
    assert "The call operator of a step callback is expected to return a boolean, but a value of type" in str(cm.value)

    class cb_hook:
        def __call__(self_, t):
            return True

        def pre_hook(self_, t):
            t.foo = True

    restart()
    prop(fp(10.0), callback=cb_hook())
    assert ta.foo


@pytest.mark.parametrize("fp", FP)
def test_events(fp):
    # :555-600
    x, v, sys_ = _pend()

    def cb0(ta, t, d_sgn, bidx):
        pass

    ta = hy.taylor_adaptive_batch(sys=sys_, state=np.array([[0.0, 0.001], [0.25, 0.2501]], dtype=fp),
                                  nt_events=[hy.nt_event_batch(v * v - 1e-6, cb0, fp_type=fp)],
                                  t_events=[hy.t_event_batch(v, fp_type=fp)], fp_type=fp)
    assert ta.with_events and len(ta.t_events) == 1 and len(ta.nt_events) == 1
    ta.propagate_until([fp(1e9), fp(1e9)])
    assert all(int(_[0]) == -1 for _ in ta.propagate_res)
    assert ta.te_cooldowns[0][0] is not None and ta.te_cooldowns[1][0] is not None
    ta.reset_cooldowns(0)
    assert ta.te_cooldowns[0][0] is None and ta.te_cooldowns[1][0] is not None
    ta.reset_cooldowns()
    assert ta.te_cooldowns[0][0] is None and ta.te_cooldowns[1][0] is None


def _s11n_cb0(ta, t, d_sgn, bidx):
    pass


class _s11n_cb1:
    def __init__(self):
        self.n = 0

    def __call__(self, ta, d_sgn, bidx):
        self.n = self.n + 1
        return True


@pytest.mark.parametrize("fp", FP)
def test_s11n(fp):
    # :602-690
    x, v, sys_ = _pend()
    ta = hy.taylor_adaptive_batch(sys=sys_, state=np.array([[0, 0.01], [0.25, 0.26]], dtype=fp),
                                  nt_events=[hy.nt_event_batch(v * v - 1e-6, _s11n_cb0, fp_type=fp)], fp_type=fp)
    for _ in range(4):
        ta.step()
    ta2 = pickle.loads(pickle.dumps(ta))
    assert np.all(ta.state == ta2.state) and np.all(ta.time == ta2.time)
    assert len(ta.t_events) == len(ta2.t_events) and len(ta.nt_events) == len(ta2.nt_events)
    ta.step()
    ta2.step()
    assert np.all(ta.state == ta2.state) and np.all(ta.time == ta2.time)
    ta = hy.taylor_adaptive_batch(sys=sys_, state=np.array([[0, 0.01], [0.25, 0.26]], dtype=fp), tol=fp(1e-6), fp_type=fp)
    assert ta.tol == fp(1e-6)
    ta.foo = "hello world"
    ta = pickle.loads(pickle.dumps(ta))
    assert ta.foo == "hello world"
    clb = _s11n_cb1()
    ta = hy.taylor_adaptive_batch(sys=sys_, state=np.array([[0, 0.01], [0.25, 0.26]], dtype=fp),
                                  t_events=[hy.t_event_batch(v, callback=clb, fp_type=fp)], fp_type=fp)
    assert id(clb) != id(ta.t_events[0].callback)
    assert ta.t_events[0].callback.n == 0
    ta.propagate_until([fp(100.0), fp(100.0)])
    ta2 = pickle.loads(pickle.dumps(ta))
    assert ta.t_events[0].callback.n == ta2.t_events[0].callback.n


@pytest.mark.parametrize("fp", FP)
def test_step_callback(fp):
    # :859-921
    from hy_b200.callback import angle_reducer

    x, v, sys_ = _pend()

    class cb_hook:
        def __call__(self_, ta):
            return True

        def pre_hook(self_, ta):
            ta.foo = True

    ta = hy.taylor_adaptive_batch(sys=sys_, state=[[fp(0.0), fp(0.01)], [fp(10.0), fp(10.01)]], fp_type=fp)
    cb1, cb2 = cb_hook(), cb_hook()
    res = ta.propagate_for(fp(10.0), callback=[cb1, cb2])
    assert isinstance(res[1], list) and isinstance(res[1][0], cb_hook) and isinstance(res[1][1], cb_hook)
    assert id(res[1][0]) == id(cb1) and id(res[1][1]) == id(cb2)
    assert hasattr(ta, "foo")
    res = ta.propagate_until(fp(20.0), callback=[cb1, angle_reducer([x]), cb2])
    assert isinstance(res[1], list) and isinstance(res[1][0], cb_hook)
    assert isinstance(res[1][1], angle_reducer) and isinstance(res[1][2], cb_hook)
    assert id(res[1][0]) == id(cb1) and id(res[1][2]) == id(cb2)
    assert 0 <= ta.state[0, 0] < fp(6.29) and 0 <= ta.state[0, 1] < fp(6.29)
    res = ta.propagate_grid([[fp(20.0), fp(20.0)], [fp(30.0), fp(30.1)]], callback=cb1)
    assert isinstance(res[0], cb_hook) and id(res[0]) == id(cb1)
    res = ta.propagate_for(fp(10.0), callback=angle_reducer([x]))
    assert isinstance(res[1], angle_reducer)
    assert 0 <= ta.state[0, 0] < fp(6.29) and 0 <= ta.state[0, 1] < fp(6.29)


def test_ensemble_batch():
    # /root/reference/heyoka/_test_ensemble.py:13-236 (thread algorithm throughout; the process algorithm for the
    # propagate_until scenarios, with and without c_output - every worker process creates its own CUDA context and
    # the continuous outputs travel back pickled)
    from hy_b200.callback import angle_reducer

    x, v, sys_ = _pend()
    ta = hy.taylor_adaptive_batch(sys=sys_, state=[[0.0] * 4] * 2)
    ics = np.zeros((10, 2, 4))
    for i in range(10):
        ics[i, 0] = [0.05 + i / 100, 0.051 + i / 100, 0.052 + i / 100, 0.053 + i / 100.0]
        ics[i, 0] = [0.025 + i / 100, 0.026 + i / 100, 0.027 + i / 100, 0.028 + i / 100.0]

    def gen(t, idx):
        t.set_time(0.0)
        t.state[:] = ics[idx]
        return t

    for algo in ("thread", "process"):
        kw = dict(algorithm=algo, max_workers=8) if algo == "thread" else dict(algorithm=algo, max_workers=2, chunksize=3)
        ret = hy.ensemble_propagate_until_batch(ta, 20.0, 10, gen, **kw)
        assert len(ret) == 10
        for i in range(10):
            ta.set_time(0.0)
            ta.state[:] = ics[i]
            ta.propagate_until(20.0)
            assert all(abs(ret[i][0].time[j] - 20.0) < 1e-7 for j in range(4))
            assert np.all(ta.state == ret[i][0].state) and ret[i][1] is None
            assert np.all(ta.time == ret[i][0].time) and ta.propagate_res == ret[i][0].propagate_res
        ret = hy.ensemble_propagate_until_batch(ta, 20.0, 10, gen, c_output=True, **kw)
        for i in range(10):
            ta.set_time(0.0)
            ta.state[:] = ics[i]
            loc = ta.propagate_until(20.0, c_output=True)
            assert np.all(ta.state == ret[i][0].state) and ret[i][1] is not None
            assert ta.propagate_res == ret[i][0].propagate_res
            assert np.all(loc[0](5.0) == ret[i][1](5.0))

    def gen10(t, idx):
        t.set_time(10.0)
        t.state[:] = ics[idx]
        return t

    ret = hy.ensemble_propagate_for_batch(ta, 20.0, 10, gen10, algorithm="thread", max_workers=8)
    for i in range(10):
        ta.set_time(10.0)
        ta.state[:] = ics[i]
        ta.propagate_for(20.0)
        assert all(abs(ret[i][0].time[j] - 30.0) < 1e-7 for j in range(4))
        assert np.all(ta.state == ret[i][0].state) and ret[i][1] is None
        assert ta.propagate_res == ret[i][0].propagate_res

    grid = np.linspace(0.0, 20.0, 80)
    splat = np.repeat(grid, 4).reshape(-1, 4)
    ret = hy.ensemble_propagate_grid_batch(ta, grid, 10, gen, algorithm="thread", max_workers=8)
    for i in range(10):
        ta.set_time(0.0)
        ta.state[:] = ics[i]
        loc = ta.propagate_grid(splat)
        assert np.all(loc[1] == ret[i][2])
        assert np.all(ta.state == ret[i][0].state) and ta.propagate_res == ret[i][0].propagate_res

    class step_cb:
        def __call__(self_, t):
            assert id(self_) != self_.orig_id      # callbacks are deep-copied per iteration
            return True

    cb = step_cb()
    cb.orig_id = id(cb)
    ret = hy.ensemble_propagate_for_batch(ta, 20.0, 10, gen, algorithm="thread", max_workers=8, callback=cb)
    assert all(isinstance(r[2], step_cb) for r in ret)
    ret = hy.ensemble_propagate_for_batch(ta, 20.0, 10, gen, algorithm="thread", max_workers=8,
                                          callback=[cb, angle_reducer([x])])
    for r in ret:
        assert isinstance(r[2], list) and len(r[2]) == 2
        assert isinstance(r[2][0], step_cb) and isinstance(r[2][1], angle_reducer)


@pytest.mark.parametrize("fp,c_out_t", [(np.float32, "continuous_output_batch_flt"), (np.float64, "continuous_output_batch_dbl")])
def test_c_output_batch(fp, c_out_t):
    # /root/reference/heyoka/test.py:1414-1771 (the scalar integrators of the comparison are one-lane batch integrators)
    from pickle import dumps, loads
    from sys import getrefcount

    x, v, sys_ = _pend()
    c_out = getattr(hy, c_out_t)()
    msg = "Cannot use a default-constructed continuous_output_batch object"

    def check_default(c):
        with pytest.raises(ValueError) as cm:
            c.n_steps
        assert msg in str(cm.value)
        assert c.batch_size == 0 and "forward" not in repr(c) and c.llvm_state.ir != ""

    for arg in (np.zeros((0,), dtype=fp), fp(1)):
        with pytest.raises(ValueError) as cm:
            c_out(arg)
        assert msg in str(cm.value)
    with pytest.raises(ValueError) as cm:
        c_out(time=[fp(0), fp(0)])
    assert msg in str(cm.value)
    assert c_out.output is None and c_out.times is None and c_out.tcs is None
    with pytest.raises(ValueError) as cm:
        c_out.bounds
    assert msg in str(cm.value)
    check_default(c_out)
    check_default(copy(c_out))
    check_default(deepcopy(c_out))
    check_default(loads(dumps(c_out)))

    ic = [[fp(0), fp(0.01), fp(0.02), fp(0.03)], [fp(0.25), fp(0.26), fp(0.27), fp(0.28)]]
    arr_ic = np.array(ic)
    ta = hy.taylor_adaptive_batch(sys=sys_, state=ic, fp_type=fp)

    def reset():
        ta.state[:] = ic
        ta.set_time([fp(0)] * 4)

    final_tm = [fp(10), fp(10.4), fp(10.5), fp(11.0)]
    check_tm = [fp(0.1), fp(1.3), fp(5.6), fp(9.1)]
    c_out, cb = ta.propagate_until(final_tm)
    assert cb is None and c_out is None
    reset()
    c_out, cb = ta.propagate_until(final_tm, c_output=False)
    assert cb is None and c_out is None
    reset()
    c_out, cb = ta.propagate_until(final_tm, c_output=True)
    assert cb is None and c_out is not None
    assert c_out(check_tm).shape == (2, 4)
    with pytest.raises(ValueError):
        c_out(check_tm)[0] = 0.5
    rc = getrefcount(c_out)
    tmp_out = c_out(check_tm)
    assert getrefcount(c_out) == rc + 1
    with pytest.raises(ValueError) as cm:
        c_out(np.zeros((1, 1, 1), dtype=fp))
    assert ("Invalid time array passed to a continuous_output_batch object: the number of dimensions must be 1 or 2, "
            "but it is 3 instead") in str(cm.value)
    for n_ in (1, 0, 5):
        with pytest.raises(ValueError) as cm:
            c_out(np.zeros((n_,), dtype=fp))
        assert ("Invalid time array passed to a continuous_output_batch object: the length must be 4 but it is {} "
                "instead".format(n_)) in str(cm.value)

    # one-lane integrators for comparison
    eps10 = np.finfo(fp).eps * 10
    scal = []
    for idx in range(4):
        one = hy.taylor_adaptive_batch(sys=sys_, state=arr_ic[:, idx:idx + 1], fp_type=fp)
        scal.append(one.propagate_until(final_tm[idx:idx + 1], c_output=True)[0])
    c_out(check_tm)
    for idx in range(4):
        scal[idx](check_tm[idx:idx + 1])
        assert np.allclose(scal[idx].output[:, 0], c_out.output[:, idx], rtol=eps10, atol=eps10)
    rc = getrefcount(c_out)
    tmp_out2 = c_out(check_tm)
    assert getrefcount(c_out) == rc + 1
    scal_res = deepcopy(c_out(fp(0.42)))
    assert np.all(scal_res == c_out([fp(0.42)] * 4))
    nc_check_tm = np.vstack([check_tm, np.zeros((4,), dtype=fp)]).T.flatten()[::2]
    c_out(nc_check_tm)
    for idx in range(4):
        assert np.allclose(scal[idx].output[:, 0], c_out.output[:, idx], rtol=eps10, atol=eps10)
    with pytest.raises(ValueError) as cm:
        c_out(np.zeros((5, 3), dtype=fp))
    assert ("Invalid time array passed to a continuous_output_batch object: the number of columns must be 4 but it is 3 "
            "instead") in str(cm.value)
    b_check_tm = np.repeat(check_tm, 5, axis=0).reshape((4, 5)).T
    out_b = c_out(b_check_tm)
    assert out_b.shape == (5, 2, 4)
    for idx in range(4):
        scal[idx](check_tm[idx:idx + 1])
        for j in range(5):
            assert np.allclose(scal[idx].output[:, 0], out_b[j, :, idx], rtol=eps10, atol=eps10)
    assert c_out(np.zeros((0, 4), dtype=fp)).shape == (0, 2, 4)

    assert c_out.times.shape == (c_out.n_steps + 1, 4) and np.all(np.isfinite(c_out.times))
    with pytest.raises(ValueError):
        c_out.times[0] = 0.5
    rc = getrefcount(c_out)
    tmp_out3 = c_out.times
    assert getrefcount(c_out) == rc + 1
    assert c_out.tcs.shape == (c_out.n_steps, 2, ta.order + 1, 4)
    with pytest.raises(ValueError):
        c_out.tcs[0] = 0.5
    rc = getrefcount(c_out)
    tmp_out4 = c_out.tcs
    assert getrefcount(c_out) == rc + 1
    assert np.all(c_out.bounds[0] == [0.0] * 4)
    assert np.allclose(c_out.bounds[1], final_tm, rtol=eps10, atol=eps10)
    assert c_out.batch_size == 4 and "forward" in repr(c_out) and c_out.llvm_state.ir != ""
    c_out = copy(c_out)
    assert c_out.tcs.shape == (c_out.n_steps, 2, ta.order + 1, 4) and c_out.llvm_state.ir != ""
    c_out = deepcopy(c_out)
    assert c_out.tcs.shape == (c_out.n_steps, 2, ta.order + 1, 4)
    c_out = loads(dumps(c_out))
    assert c_out.llvm_state.ir != "" and c_out.tcs.shape == (c_out.n_steps, 2, ta.order + 1, 4)

    class foo:
        pass

    c_out_copy = deepcopy(c_out)
    c_out_copy.bar = foo()
    assert id(c_out_copy.bar) == id(copy(c_out_copy).bar)
    assert id(c_out_copy.bar) != id(deepcopy(c_out_copy).bar)
    assert np.all(c_out_copy(fp(0.1)) == copy(c_out_copy)(fp(0.1)))
    assert np.all(c_out_copy(fp(0.1)) == deepcopy(c_out_copy)(fp(0.1)))
    c_out.foo = []
    c_out = loads(dumps(c_out))
    assert c_out.tcs.shape == (c_out.n_steps, 2, ta.order + 1, 4) and c_out.foo == []
    del tmp_out, tmp_out2, tmp_out3, tmp_out4


def test_event_detection_batch():
    # /root/reference/heyoka/test.py:694-999

    x, v, sys_ = _pend()
    counter, cur_time = [0] * 2, [0.0] * 2
    ta_id = [None]

    def cb0(ta, t, d_sgn, bidx):
        assert t > cur_time[bidx]
        assert counter[bidx] % 3 == 0 or counter[bidx] % 3 == 2
        assert ta_id[0] == id(ta)
        counter[bidx] += 1
        cur_time[bidx] = t

    def cb1(ta, t, d_sgn, bidx):
        assert t > cur_time[bidx]
        assert counter[bidx] % 3 == 1
        assert ta_id[0] == id(ta)
        counter[bidx] += 1
        cur_time[bidx] = t

    ta = hy.taylor_adaptive_batch(sys=sys_, state=[[0.0, 0.001], [0.25, 0.2501]],
                                  nt_events=[hy.nt_event_batch(v * v - 1e-10, cb0), hy.nt_event_batch(v, cb1)])
    ta_id[0] = id(ta)
    ta.propagate_until([4.0, 4.0])
    assert all(_[0] == hy.taylor_outcome.time_limit for _ in ta.propagate_res)
    assert counter == [12, 12]

    class ccb0:
        def __init__(self):
            self.lst = []

        def __call__(self, ta, t, d_sgn, bidx):
            pass

    class ccb1(ccb0):
        pass

    ta = hy.taylor_adaptive_batch(sys=sys_, state=[[0.0, 0.001], [0.25, 0.2501]],
                                  nt_events=[hy.nt_event_batch(v * v - 1e-10, ccb0()), hy.nt_event_batch(v, ccb1()),
                                             hy.nt_event_batch(v, ccb1())])
    # (the reference's event objects alias storage inside the C++ integrator and therefore hold it -
    #  getrefcount(ta) grows by 3 there; here they are plain Python objects owned by the integrator)
    nt_list = ta.nt_events
    assert len(nt_list) == 3
    for i in range(3):
        assert id(ta.nt_events[i].callback) == id(ta.nt_events[i].callback)
        assert id(ta.nt_events[i].callback.lst) == id(ta.nt_events[i].callback.lst)
    ta_copy = deepcopy(ta)
    for i in range(3):
        assert id(ta_copy.nt_events[i].callback) != id(ta.nt_events[i].callback)
        assert id(ta_copy.nt_events[i].callback.lst) != id(ta.nt_events[i].callback.lst)
    del nt_list

    def cb2(ta, t):
        pass

    ta = hy.taylor_adaptive_batch(sys=sys_, state=[[0.0, 0.001], [0.25, 0.2501]],
                                  nt_events=[hy.nt_event_batch(v * v - 1e-10, cb2)])
    with pytest.raises(RuntimeError):
        ta.propagate_until([4.0, 4.0])

    # terminal events
    counter_t, counter_nt, cur_time = [0] * 2, [0] * 2, [0.0] * 2

    def ncb(ta, t, d_sgn, bidx):
        assert t > cur_time[bidx] and ta_id[0] == id(ta)
        counter_nt[bidx] += 1
        cur_time[bidx] = t

    def tcb(ta, d_sgn, bidx):
        assert ta.time[bidx] > cur_time[bidx] and ta_id[0] == id(ta)
        counter_t[bidx] += 1
        cur_time[bidx] = ta.time[bidx]
        return True

    ta = hy.taylor_adaptive_batch(sys=sys_, state=[[0.0, 0.001], [0.25, 0.2501]],
                                  nt_events=[hy.nt_event_batch(v * v - 1e-10, ncb)],
                                  t_events=[hy.t_event_batch(v, callback=tcb)])
    ta_id[0] = id(ta)
    while True:
        ta.step()
        if all(_[0] > hy.taylor_outcome.success for _ in ta.step_res):
            break
    assert all(int(_[0]) == 0 for _ in ta.step_res) and all(_ < 1 for _ in ta.time)
    assert all(_ == 1 for _ in counter_nt) and all(_ == 1 for _ in counter_t)
    while True:
        ta.step()
        if all(_[0] > hy.taylor_outcome.success for _ in ta.step_res):
            break
    assert all(int(_[0]) == 0 for _ in ta.step_res) and all(_ > 1 for _ in ta.time)
    assert all(_ == 3 for _ in counter_nt) and all(_ == 2 for _ in counter_t)

    class tcb0:
        def __init__(self):
            self.lst = []

        def __call__(self, ta, d_sgn, bidx):
            pass

    class tcb1(tcb0):
        pass

    ta = hy.taylor_adaptive_batch(sys=sys_, state=[[0.0, 0.001], [0.25, 0.2501]],
                                  t_events=[hy.t_event_batch(v * v - 1e-10, callback=tcb0()),
                                            hy.t_event_batch(v, callback=tcb1()), hy.t_event_batch(v, callback=tcb1())])
    t_list = ta.t_events
    assert len(t_list) == 3
    for i in range(3):
        assert id(ta.t_events[i].callback) == id(ta.t_events[i].callback)
        assert id(ta.t_events[i].callback.lst) == id(ta.t_events[i].callback.lst)
    ta_copy = deepcopy(ta)
    for i in range(3):
        assert id(ta_copy.t_events[i].callback) != id(ta.t_events[i].callback)
        assert id(ta_copy.t_events[i].callback.lst) != id(ta.t_events[i].callback.lst)
    del t_list

    ta = hy.taylor_adaptive_batch(sys=sys_, state=[[0.0, 0.001], [0.25, 0.2501]],
                                  t_events=[hy.t_event_batch(v * v - 1e-10, callback=cb2)])
    with pytest.raises(RuntimeError):
        ta.propagate_until([4.0, 4.0])

    def cb3(ta, d_sgn, bidx):
        return "hello"

    ta = hy.taylor_adaptive_batch(sys=sys_, state=[[0.0, 0.001], [0.25, 0.2501]],
                                  t_events=[hy.t_event_batch(v * v - 1e-10, callback=cb3)])
    with pytest.raises(RuntimeError) as cm:
        ta.propagate_until([4.0, 4.0])
    assert "in the construction of the return value of an event callback" in str(cm.value)


def test_var_integrator_batch():
    # /root/reference/heyoka/_test_var_integrator.py:143-258
    from sys import getrefcount

    x, v = hy.make_vars("x", "v")
    orig_sys = [(x, v), (v, hy.cos(hy.time) - hy.par[0] * v - hy.sin(x))]
    vsys = hy.var_ode_sys(orig_sys, hy.var_args.vars, order=2)
    ta = hy.taylor_adaptive_batch(vsys, [[0.2, 0.21], [0.3, 0.31]], pars=[[0.4, 0.41]], time=[0.5, 0.51],
                                  compact_mode=True)
    assert ta.dim > 2 and ta.n_orig_sv == 2 and ta.is_variational and ta.vorder == 2 and ta.vargs == [x, v]
    rc = getrefcount(ta)
    ts = ta.tstate
    assert getrefcount(ta) == rc + 1
    assert ts.shape == (2, 2) and np.all(ts == [[0.0, 0.0], [0.0, 0.0]])
    with pytest.raises(ValueError):
        ta.tstate[0] = 0.5
    assert ta.get_vslice(order=0) == slice(0, 2, None) and ta.get_vslice(order=0, component=1) == slice(1, 2, None)
    assert ta.get_vslice(order=1) == slice(2, 6, None) and ta.get_vslice(order=1, component=1) == slice(4, 6, None)
    assert ta.get_mindex(0) == [0, 0, 0] and ta.get_mindex(1) == [1, 0, 0] and ta.get_mindex(2) == [0, 1, 0]
    assert ta.get_mindex(3) == [0, 0, 1] and ta.get_mindex(4) == [1, 1, 0] and ta.get_mindex(i=5) == [1, 0, 1]
    ta.propagate_until(3.0)
    ts2 = ta.eval_taylor_map([[0.0, 0.0], [0.0, 0.0]])
    assert getrefcount(ta) == rc + 2
    assert np.shares_memory(ts, ts2) and np.all(ts2 == ta.state[:2])
    ts2 = ta.eval_taylor_map(np.array([[0.0, 0.0], [0.0, 0.0]]))
    assert np.shares_memory(ts, ts2) and np.all(ts2 == ta.state[:2])
    with pytest.raises(TypeError) as cm:
        ta.eval_taylor_map(np.array([0.0, 0.0], dtype=np.int32))
    assert "Invalid dtype detected for the inputs of a Taylor map evaluation:" in str(cm.value)
    with pytest.raises(ValueError) as cm:
        ta.eval_taylor_map(np.array([0.0, 0.0, 0.0, 0.0])[::2])
    assert "Invalid inputs array detected in a Taylor map evaluation: the array is not C-style contiguous, please " in str(cm.value)
    with pytest.raises(ValueError) as cm:
        ta.eval_taylor_map(np.array([0.0, 0.0]))
    assert "The array of inputs provided for the evaluation of a Taylor map has 1 dimension(s), " in str(cm.value)
    with pytest.raises(ValueError) as cm:
        ta.eval_taylor_map(np.array([[0.0, 0.0]]))
    assert ("The array of inputs provided for the evaluation of a Taylor map has 1 row(s), but it must have 2 row(s) "
            "instead") in str(cm.value)
    with pytest.raises(ValueError) as cm:
        ta.eval_taylor_map(np.array([[0.0], [0.0]]))
    assert ("The array of inputs provided for the evaluation of a Taylor map has 1 column(s), but it must have 2 "
            "column(s) instead") in str(cm.value)
    with pytest.raises(ValueError) as cm:
        ta.eval_taylor_map(ta.state[:2])
    assert "may overlap" in str(cm.value)
    with pytest.raises(ValueError) as cm:
        ta.eval_taylor_map(ta.tstate)
    assert "may overlap" in str(cm.value)


def test_ensemble_argument_errors():
    # /root/reference/heyoka/_test_ensemble.py:394-438 (the checks are shared by the scalar and the batch functions)
    x, v, sys_ = _pend()
    ta = hy.taylor_adaptive_batch(sys=sys_, state=[[0.0] * 4] * 2)

    def gen(t, idx):
        return t

    with pytest.raises(TypeError) as cm:
        hy.ensemble_propagate_until_batch(ta, 20.0, "a", gen)
    assert "The n_iter parameter must be an integer, but an object of type" in str(cm.value)
    with pytest.raises(ValueError) as cm:
        hy.ensemble_propagate_until_batch(ta, 20.0, -1, gen)
    assert "The n_iter parameter must be non-negative" in str(cm.value)
    for fn in (hy.ensemble_propagate_until_batch, hy.ensemble_propagate_for_batch):
        with pytest.raises(TypeError) as cm:
            fn(ta, [20.0], 10, gen)
        assert ("Cannot perform an ensemble propagate_until/for(): the final epoch/time interval must be a scalar, "
                "not an iterable object") in str(cm.value)
    with pytest.raises(ValueError) as cm:
        hy.ensemble_propagate_grid_batch(ta, [[20.0, 20.0]], 10, gen)
    assert ("Cannot perform an ensemble propagate_grid(): the input time grid must be one-dimensional, but instead it "
            "has 2 dimensions") in str(cm.value)
    with pytest.raises(TypeError) as cm:
        hy.ensemble_propagate_until_batch(ta, 20.0, 10, gen, max_delta_t=[10])
    assert ('Cannot perform an ensemble propagate_until/for/grid(): the "max_delta_t" argument must be a scalar, '
            "not an iterable object") in str(cm.value)
    with pytest.raises(TypeError):
        hy.ensemble_propagate_until_batch(ta, 20.0, 10, gen, chunksize=1)   # not recognised in threaded mode
    assert hy.ensemble_propagate_until_batch(ta, 20.0, 0, gen) == []
