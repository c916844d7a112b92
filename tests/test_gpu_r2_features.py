"""GPU: round-2 features of the C ABI and the front end - resumable launches with an active-lane
mask, single-pass appendable continuous output, hy_clone / context pool, lane-sharded multi-device
integrators, the device-side angle reducer, propagate_grid with callbacks, process ensembles."""

import copy
import pickle

import numpy as np
import pytest

import hy_b200 as hy
from hy_b200 import _cabi, _devctx
from hy_b200 import decompose as D
from hy_b200 import workloads as W
from oracle.c_oracle import COracle
from oracle.np_oracle import NpTaylorBatch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))))


class _CountCb:
    def __init__(self):
        self.n = 0

    def __call__(self, ta):
        self.n += 1
        return True


# ---------------------------------------------------------------- resumable launches
@pytest.mark.parametrize("which", ["pendulum", "oss", "cr3bp"])
def test_step_callback_run_is_bit_identical_to_plain_run(which):
    # A do-nothing step callback forces one launch per batch step (resume + active mask): state,
    # times, step counts, min/max h must equal the single-launch run bit for bit
    # (reference: the callback does not change the numerics, step_cb_utils.cpp:70-98).
    if which == "pendulum":
        sys_, ic, T = W.pendulum_sys(), W.PEND_IC, [10.0, 11.0, 12.0, 13.0]
    elif which == "oss":
        sys_, ic, T = W.oss_sys(), W.oss_ensemble(6), 30.0
    else:
        sys_, ic, T = W.cr3bp_sys(0.01), W.cr3bp_ensemble(7), 5.0
    a = hy.taylor_adaptive_batch(sys_, ic)
    b = hy.taylor_adaptive_batch(sys_, ic)
    a.propagate_until(T)
    cb = _CountCb()
    _, ret = b.propagate_until(T, callback=cb)
    assert ret is cb
    assert np.array_equal(a.state, b.state)
    assert np.array_equal(a.time, b.time) and np.array_equal(a.dtime[1], b.dtime[1])
    ra, rb = a.propagate_res_arrays, b.propagate_res_arrays
    for x, y in zip(ra, rb):
        assert np.array_equal(x, y)
    assert cb.n == int(ra[3].max())            # one call per batch step
    # propagate_for with max_steps and max_delta_t through the same machinery
    a.propagate_for(3.0, max_steps=7, max_delta_t=0.05)
    b.propagate_for(3.0, max_steps=7, max_delta_t=0.05, callback=_CountCb())
    assert np.array_equal(a.state, b.state)
    for x, y in zip(a.propagate_res_arrays, b.propagate_res_arrays):
        assert np.array_equal(x, y)
    assert all(r[0] == hy.taylor_outcome.step_limit for r in b.propagate_res)


def test_cb_stop_and_grid_with_callback():
    sys_ = W.pendulum_sys()
    ta = hy.taylor_adaptive_batch(sys_, W.PEND_IC)

    class Stop:
        n = 0

        def __call__(self, ta):
            self.n += 1
            return self.n < 5

    ta.propagate_until(100.0, callback=Stop())
    assert all(r[0] == hy.taylor_outcome.cb_stop for r in ta.propagate_res)
    assert all(r[3] == 5 for r in ta.propagate_res)
    # propagate_grid(..., callback=) (expose_batch_integrators.cpp:315-392)
    grid = np.repeat(np.linspace(0.0, 10.0, 41), 4).reshape(-1, 4)
    a = hy.taylor_adaptive_batch(sys_, W.PEND_IC)
    b = hy.taylor_adaptive_batch(sys_, W.PEND_IC)
    _, oa = a.propagate_grid(grid)
    cb = _CountCb()
    ret, ob = b.propagate_grid(grid, callback=cb)
    assert ret is cb and cb.n > 10
    assert np.array_equal(oa, ob)
    for x, y in zip(a.propagate_res_arrays, b.propagate_res_arrays):
        assert np.array_equal(x, y)
    # early stop: NaN past the exit
    c = hy.taylor_adaptive_batch(sys_, W.PEND_IC)
    _, oc = c.propagate_grid(grid, callback=Stop())
    assert np.all(np.isnan(oc[-1])) and np.all(np.isfinite(oc[0]))


def test_nt_event_callbacks_with_c_output_fire_once_and_cover_the_interval():
    # c_output + events/callbacks returned a truncated record and fired twice.
    x, v = hy.make_vars("x", "v")
    log = []

    def cb(ta, t, d_sgn, bidx):
        log.append((bidx, float(t), d_sgn))

    ic = np.array([[-0.05, -0.06], [0.0, 0.0]])
    ta = hy.taylor_adaptive_batch(W.pendulum_sys(), ic, nt_events=[hy.nt_event_batch(v, cb)])
    c_out, _ = ta.propagate_until(5.0, c_output=True)
    # golden A10 (Event detection.ipynb:289): v = 0 at 0, 1.0037..., five times in [0, 5)
    t0 = sorted(t for b, t, _ in log if b == 0)
    gold = [0.0, 1.003701787940065, 2.00740357588013, 3.011105363820195, 4.01480715176026]
    assert len(t0) == 5 and np.max(np.abs(np.array(t0) - gold)) < 5e-15
    assert len([1 for b, _, _ in log if b == 1]) == 5
    assert c_out is not None
    lo, hi = c_out.bounds
    assert np.all(lo == 0.0) and np.all(hi == 5.0)
    assert c_out.n_steps == int(ta.propagate_res_arrays[3].max())
    # the record reproduces the state along the way: compare with a plain run to the same times
    for tq in (0.7, 2.2, 4.9):
        ref = hy.taylor_adaptive_batch(W.pendulum_sys(), ic)
        ref.propagate_until(tq)
        assert _rel(c_out(tq), ref.state) < 1e-13
    assert np.max(np.abs(c_out(5.0) - ta.state)) < 1e-14


# ---------------------------------------------------------------- continuous output
def test_c_output_is_its_own_object_and_pool_is_recycled():
    sys_ = W.cr3bp_sys(0.01)
    ic = W.cr3bp_ensemble(33)
    ta = hy.taylor_adaptive_batch(sys_, ic)
    c1, _ = ta.propagate_until(3.0, c_output=True)
    s3 = ta.state.copy()
    c2, _ = ta.propagate_until(6.0, c_output=True)   # a second record: c1 must stay valid
    assert np.all(c1.bounds[0] == 0.0) and np.all(c1.bounds[1] == 3.0)
    assert np.all(c2.bounds[0] == 3.0) and np.all(c2.bounds[1] == 6.0)
    assert np.max(np.abs(c1(3.0) - s3)) < 1e-14
    assert np.max(np.abs(c2(6.0) - ta.state)) < 1e-14
    assert np.max(np.abs(c2(3.0) - s3)) < 1e-14
    del c1, c2
    c3, _ = ta.propagate_until(7.0, c_output=True)    # reuses a recycled pool
    assert np.max(np.abs(c3(7.0) - ta.state)) < 1e-14
    # outlives the integrator
    del ta
    assert np.all(np.isfinite(c3(6.5)))


def test_c_output_pool_growth_resumes_lanes_transparently():
    # 4 lanes x ~4600 steps need ~2300 chunks: the first pool segment (256 chunks) is exhausted
    # several times; lanes are paused, the pool grows, they resume - invisible to the caller.
    sys_ = W.pendulum_sys()
    ic = W.PEND_IC
    a = hy.taylor_adaptive_batch(sys_, ic)
    b = hy.taylor_adaptive_batch(sys_, ic)
    a.propagate_until(1000.0)
    c_out, _ = b.propagate_until(1000.0, c_output=True)
    assert np.array_equal(a.state, b.state)
    for x, y in zip(a.propagate_res_arrays, b.propagate_res_arrays):
        assert np.array_equal(x, y)
    ns = b.propagate_res_arrays[3]
    assert c_out.n_steps == ns.max() and ns.max() > 4000
    assert b._ctx.last_timing()[1] > 1           # more than one launch
    tq = np.array([[1.0, 2.0, 3.0, 4.0], [500.5, 600.25, 700.125, 999.0], [1000.0] * 4])
    out = c_out(tq)
    for q in range(2):
        ref = hy.taylor_adaptive_batch(sys_, ic)
        ref.propagate_until(list(tq[q]))
        assert _rel(out[q], ref.state) < 1e-10
    assert np.max(np.abs(out[2] - b.state)) < 1e-13
    tcs, times = c_out.tcs, c_out.times
    assert tcs.shape == (c_out.n_steps, 2, 21, 4) and times.shape == (c_out.n_steps + 1, 4)
    for l in range(4):
        assert np.all(np.isfinite(tcs[: ns[l], :, :, l])) and np.all(np.isnan(tcs[ns[l]:, :, :, l]))
        assert times[ns[l], l] == 1000.0 and np.all(np.diff(times[: ns[l] + 1, l]) > 0)


# ---------------------------------------------------------------- copies
def test_copies_carry_tc_last_h_and_cooldowns():
    # update_d_output() on a copied / unpickled integrator read zeros.
    x, v = hy.make_vars("x", "v")
    ta = hy.taylor_adaptive_batch(W.pendulum_sys(), W.PEND_IC, t_events=[hy.t_event_batch(v)])
    ta.propagate_until(2.0, write_tc=True)
    tq = list(ta.time - 0.5 * ta.last_h)
    want = np.array(ta.update_d_output(tq))
    cds = ta.te_cooldowns
    for mk in (copy.copy, copy.deepcopy, lambda t: pickle.loads(pickle.dumps(t))):
        tb = mk(ta)
        assert np.array_equal(np.array(tb.tc), np.array(ta.tc))
        assert np.array_equal(tb.last_h, ta.last_h)
        assert np.array_equal(np.array(tb.update_d_output(tq)), want)
        assert tb.te_cooldowns == cds
        tb.step()
        tc = mk(ta)
        tc.step()
        assert np.array_equal(tb.state, tc.state)


def test_hy_clone_is_a_deep_copy_without_rescheduling():
    sys_ = W.oss_sys()
    ic = W.oss_ensemble(8)
    ta = hy.taylor_adaptive_batch(sys_, ic)
    ta.propagate_until(5.0)
    ctx2 = ta._ctx.clone()
    assert ctx2.launch_info() == ta._ctx.launch_info()
    st = np.zeros_like(ic)
    th = np.zeros(8)
    ctx2.download(state=st, t_hi=th)
    assert np.array_equal(st, ta.state) and np.all(th == 5.0)
    oc = np.zeros(8, dtype=np.int64)
    ns = np.zeros(8, dtype=np.uint64)
    ctx2.propagate(np.full(8, 7.0), 0, 0, None, 0, 0, oc, None, None, ns)
    ctx2.download(state=st)
    ta.propagate_until(7.0)
    assert np.array_equal(st, ta.state)           # same program, same arithmetic


# ---------------------------------------------------------------- ensembles / multi-device
def test_ensemble_reuses_contexts_and_matches_serial():
    sys_ = W.cr3bp_sys(0.01)
    base = W.cr3bp_ensemble(16)
    ta = hy.taylor_adaptive_batch(sys_, base)
    ics = [W.cr3bp_ensemble(16, seed=100 + i) for i in range(40)]

    def gen(t, i):
        t.state[:] = ics[i]
        return t

    created = []
    orig = _cabi.lib().hy_create

    ret = hy.ensemble_propagate_until_batch(ta, 4.0, 40, gen)
    assert len(ret) == 40
    for i in (0, 7, 39):
        s = hy.taylor_adaptive_batch(sys_, ics[i])
        s.propagate_until(4.0)
        assert np.array_equal(ret[i][0].state, s.state)       # _test_ensemble.py:68-78: bit-exact
        assert ret[i][0].propagate_res == s.propagate_res
    # finished iterations hold no device context; the pool holds at most one per worker thread
    assert all(r[0]._ctx_obj is None for r in ret)
    n_pool = sum(len(v) for v in _devctx.POOL._free.values())
    assert 1 <= n_pool <= 32
    # an iteration object is a full integrator: it can go on
    r0 = ret[0][0]
    r0.propagate_until(5.0)
    s = hy.taylor_adaptive_batch(sys_, ics[0])
    s.propagate_until(4.0)
    s.propagate_until(5.0)
    assert np.array_equal(r0.state, s.state)


def test_ensemble_device_placement():
    ndev = _cabi.device_count()
    if ndev < 2:
        pytest.skip("needs 2 GPUs")
    sys_ = W.oss_sys()
    ta = hy.taylor_adaptive_batch(sys_, W.oss_ensemble(8))
    seen = {}

    def gen(t, i):
        t.state[:] = W.oss_ensemble(8, seed=i)
        seen[i] = t._device
        return t

    ret = hy.ensemble_propagate_until_batch(ta, 10.0, 2 * ndev, gen)
    assert [seen[i] for i in range(2 * ndev)] == [i % ndev for i in range(2 * ndev)]
    for i in range(2 * ndev):
        s = hy.taylor_adaptive_batch(sys_, W.oss_ensemble(8, seed=i))
        s.propagate_until(10.0)
        assert np.array_equal(ret[i][0].state, s.state)


@pytest.mark.parametrize("devs", [[0, 0], "all"])
def test_lane_sharded_integrator_matches_single_context(devs):
    # ONE integrator split by trajectory range over several contexts (two on one GPU also works,
    # so the path is covered on a 1-GPU box); results are bit-identical to the unsplit run.
    sys_ = W.oss_sys()
    ic = W.oss_ensemble(37)
    a = hy.taylor_adaptive_batch(sys_, ic)
    b = hy.taylor_adaptive_batch(sys_, ic, device=devs)
    if devs == "all" and _cabi.device_count() < 2:
        assert isinstance(b._ctx, _cabi.Context)
    else:
        assert isinstance(b._ctx, _devctx.MultiContext)
        assert len({id(p) for p in b._ctx.parts}) >= 2
    a.step(write_tc=True)
    b.step(write_tc=True)
    assert np.array_equal(a.state, b.state) and a.step_res == b.step_res
    assert np.array_equal(np.array(a.tc), np.array(b.tc))
    a.propagate_until(20.0)
    b.propagate_until(20.0)
    assert np.array_equal(a.state, b.state) and np.array_equal(a.time, b.time)
    for x, y in zip(a.propagate_res_arrays, b.propagate_res_arrays):
        assert np.array_equal(x, y)
    ca, _ = a.propagate_for(5.0, c_output=True)
    cb, _ = b.propagate_for(5.0, c_output=True)
    tq = np.repeat(np.linspace(20.0, 25.0, 5), 37).reshape(5, 37)
    assert np.array_equal(ca(tq), cb(tq))
    grid = np.repeat(np.linspace(25.0, 30.0, 7), 37).reshape(7, 37)
    _, ga = a.propagate_grid(grid)
    _, gb = b.propagate_grid(grid)
    assert np.array_equal(ga, gb)


def test_process_ensemble_spawn():
    # _ensemble_impl.py:102-138: spawn context, active serialization backend, chunksize
    sys_ = W.pendulum_sys()
    ta = hy.taylor_adaptive_batch(sys_, W.PEND_IC)
    ret = hy.ensemble_propagate_for_batch(ta, 2.0, 3, _proc_gen, algorithm="process", max_workers=2,
                                          chunksize=2)
    assert len(ret) == 3
    for i in range(3):
        s = hy.taylor_adaptive_batch(sys_, W.PEND_IC + 0.01 * i)
        s.propagate_for(2.0)
        assert np.array_equal(ret[i][0].state, s.state)
    with pytest.raises(TypeError):
        hy.ensemble_propagate_for_batch(ta, 2.0, 3, _proc_gen, chunksize=2)   # thread mode: no chunksize
    assert hy.get_serialization_backend().__name__ in ("cloudpickle", "pickle")
    with pytest.raises(ValueError):
        hy.set_serialization_backend("nope")


def _proc_gen(t, i):
    t.state[:] = W.PEND_IC + 0.01 * i
    return t


# ---------------------------------------------------------------- angle reducer
@pytest.mark.parametrize("fp", [np.float64, np.float32])
def test_angle_reducer_runs_on_the_device(fp):
    # expose_callbacks.cpp:67-72 / _test_batch_integrator.py: a rotating pendulum, x kept in [0, 2 pi)
    x, v = hy.make_vars("x", "v")
    sys_ = [(x, v), (v, -9.8 * hy.sin(x))]
    ic = np.array([[0.0, 0.1, 0.2, 0.3], [8.0, 8.5, 9.0, 9.5]], dtype=fp)   # over the top
    a = hy.taylor_adaptive_batch(sys_, ic, fp_type=fp)
    b = hy.taylor_adaptive_batch(sys_, ic, fp_type=fp)
    red = hy.callback.angle_reducer([x])
    _, ret = a.propagate_until(fp(20.0), callback=red)
    assert ret is red
    assert a._ctx.last_timing()[1] == 1                    # one launch: no Python per step

    class HostReducer:                                     # the same thing as a Python callback
        def __call__(self, ta):
            ta.state[0] -= 2 * np.pi * np.floor(ta.state[0] / (2 * np.pi))
            return True

    b.propagate_until(fp(20.0), callback=HostReducer())
    assert np.all(a.state[0] >= 0) and np.all(a.state[0] < 2 * np.pi)
    tol = 1e-9 if fp == np.float64 else 2e-3
    assert _rel(a.state.astype(np.float64), b.state.astype(np.float64)) < tol
    assert np.array_equal(a.propagate_res_arrays[3], b.propagate_res_arrays[3]) or fp == np.float32
    # the unreduced run winds x up far beyond 2 pi
    c = hy.taylor_adaptive_batch(sys_, ic, fp_type=fp)
    c.propagate_until(fp(20.0))
    assert np.all(c.state[0] > 20.0)
    d = c.state[0].astype(np.float64)
    assert _rel(np.sin(a.state[0].astype(np.float64)), np.sin(d)) < (1e-8 if fp == np.float64 else 5e-2)
    # mixed with a Python callback it falls back to the host loop, same result
    e = hy.taylor_adaptive_batch(sys_, ic, fp_type=fp)
    e.propagate_until(fp(20.0), callback=[hy.callback.angle_reducer([x]), _CountCb()])
    assert _rel(e.state.astype(np.float64), a.state.astype(np.float64)) < tol


def test_cout_eval_with_device_pointers_matches_host_path():
    # hy_cout_eval_dev: query times / output in device memory (what a GPU-resident consumer uses;
    # bench.py's device arm) - the same numbers as the host-pointer call
    import torch

    B = 300
    sys_, ic = W.cr3bp_sys(0.01), W.cr3bp_ensemble(B)
    ta = hy.taylor_adaptive_batch(sys_, ic)
    ta._push()
    oc = np.zeros(B, dtype=np.int64)
    ns = np.zeros(B, dtype=np.uint64)
    ta._ctx.propagate(np.full(B, 5.0), 0, 0, None, 0, 1, oc, None, None, ns)
    rec = ta._ctx.cout_detach()
    K = 7
    tq = np.repeat(np.linspace(0.3, 4.7, K), B).reshape(K, B)
    out_h = np.empty((K, 6, B))
    rec.eval(tq, K, out_h)
    d_tq = torch.from_numpy(tq).cuda()
    d_out = torch.empty((K, 6, B), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    rec.eval_dev(d_tq.data_ptr(), K, d_out.data_ptr())
    assert np.array_equal(d_out.cpu().numpy(), out_h)
    rec.close()


def test_large_dense_output_lands_in_recycled_pinned_buffers():
    # c_out(t) with a [k, B] time grid: results of a MiB and more are written into page-locked buffers of
    # _cabi.OUT_POOL; a buffer is reused only after the array that holds it is gone
    sys_ = W.cr3bp_sys(0.01)
    B, K = 2048, 16
    ta = hy.taylor_adaptive_batch(sys_, W.cr3bp_ensemble(B))
    c_out, _ = ta.propagate_until(2.0, c_output=True)
    tq = np.linspace(0.0, 2.0, K)[:, None] * np.ones((1, B))
    a = c_out(tq)
    pa = a.ctypes.data
    assert a.shape == (K, 6, B) and a.flags.writeable and type(a.base).__name__ == "_PinnedBlock"
    b = c_out(tq[::-1].copy())
    assert b.ctypes.data != pa                       # `a` is alive: its buffer is not handed out again
    assert np.array_equal(a, b[::-1])
    assert np.allclose(a[-1], ta.state, rtol=0, atol=1e-12)   # t = 2: the final state
    keep = a.copy()
    del a
    c = c_out(tq)
    assert c.ctypes.data == pa and np.array_equal(c, keep)   # the released buffer, reused
    small = c_out(tq[:1])                            # below a MiB: an ordinary array
    assert small.base is None or type(small.base).__name__ != "_PinnedBlock"
    assert np.array_equal(small[0], keep[0])


@pytest.mark.parametrize("devs", [[0, 0], "all"])
def test_lane_sharded_cr3bp_with_events_and_c_output(devs):
    # the CR3BP register kernel with terminal events (event workspace in its global slab, one per context)
    # and with continuous output (order-major records, TMA bulk copy), split over several contexts / cloned
    mu = 0.01
    x, y, z = hy.make_vars("x", "y", "z")
    evs = [(x - mu) ** 2 + y * y + z * z - 0.2 ** 2, x * x + y * y + z * z - 1.3 ** 2]
    B = 150
    rng = np.random.default_rng(3)
    ic = np.array([-0.80, 0.0, 0.0, 0.0, -0.6276410653920693, 0.0])[:, None] * np.ones((1, B))
    ic[0] += rng.uniform(-1e-2, 1e-2, B)
    ic[4] += rng.uniform(-1e-2, 1e-2, B)
    mk = lambda **kw: hy.taylor_adaptive_batch(W.cr3bp_sys(mu), ic, t_events=[hy.t_event_batch(e) for e in evs], **kw)
    a, b = mk(), mk(device=devs)
    assert a._ctx.launch_info()["kernel_variant"] == 203
    a.propagate_until(12.0)
    b.propagate_until(12.0)
    assert (a.propagate_res_arrays[0] > -10).sum() > 5          # some lanes stopped at an event
    assert np.array_equal(a.state, b.state) and np.array_equal(a.time, b.time)
    for u, v in zip(a.propagate_res_arrays, b.propagate_res_arrays):
        assert np.array_equal(u, v)
    c = copy.deepcopy(a)                                        # hy_clone: its own slab
    a.propagate_until(14.0)
    c.propagate_until(14.0)
    assert np.array_equal(a.state, c.state)
    # continuous output of the event-free system on the same split
    p, q = hy.taylor_adaptive_batch(W.cr3bp_sys(mu), ic), hy.taylor_adaptive_batch(W.cr3bp_sys(mu), ic, device=devs)
    cp, _ = p.propagate_until(3.0, c_output=True)
    cq, _ = q.propagate_until(3.0, c_output=True)
    tq = np.repeat(np.linspace(0.0, 3.0, 9), B).reshape(9, B)
    assert np.array_equal(cp(tq), cq(tq))
    assert np.array_equal(np.array(cp.tcs), np.array(cq.tcs), equal_nan=True)   # (NaN past a lane's own step count)
    assert np.array_equal(np.array(p.tc), np.array(q.tc))       # tc: the last step's coefficients
