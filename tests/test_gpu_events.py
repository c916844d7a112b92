"""GPU: device event detection (fast exclusion + Descartes/bisection root
isolation inside the propagate kernel) against the notebook golden event
times, the reference's own known-answer counts
(/root/reference/heyoka/test.py:693-760, _test_batch_integrator.py:555-600) and
the numpy oracle's independent root finder."""

import json
import os

import numpy as np
import pytest

import hy_b200 as hy
from oracle.np_oracle import NpTaylorBatch

import common

pytestmark = pytest.mark.gpu
G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "notebook_golden.json")))


def test_nt_event_times_golden_A10():
    g = G["pendulum_events"]
    x, v = hy.make_vars("x", "v")
    log = []

    def cb(ta, t, d_sgn, bidx):
        log.append((bidx, float(t), d_sgn))

    ic = np.array(g["ic"])[:, None] * np.ones((1, 3))
    ta = hy.taylor_adaptive_batch(common.pendulum_sys(), ic, nt_events=[hy.nt_event_batch(v, cb)])
    assert ta.with_events
    ta.propagate_until(5.0)
    for lane in range(3):
        times = [t for b, t, s in log if b == lane]
        assert len(times) == 5
        assert np.max(np.abs(np.array(times) - np.array(g["v_zero_times"]))) < 1e-15 * 5
        sg = [s for b, t, s in log if b == lane]
        assert sg == [1, -1, 1, -1, 1]
    # direction filter
    log.clear()
    ta = hy.taylor_adaptive_batch(
        common.pendulum_sys(), ic,
        nt_events=[hy.nt_event_batch(v, cb, direction=hy.event_direction.positive)])
    ta.propagate_until(5.0)
    times = [t for b, t, s in log if b == 0]
    assert np.max(np.abs(np.array(times) - np.array(g["v_zero_times_positive"]))) < 1e-14


def test_two_close_events_golden():
    g = G["pendulum_events"]
    x, v = hy.make_vars("x", "v")
    log = []
    ev0 = hy.nt_event_batch(v, lambda ta, t, d, b: log.append((b, 0, float(t))))
    ev1 = hy.nt_event_batch(v * v - 1e-12, lambda ta, t, d, b: log.append((b, 1, float(t))))
    ic = np.array(g["ic"])[:, None] * np.ones((1, 2))
    ta = hy.taylor_adaptive_batch(common.pendulum_sys(), ic, nt_events=[ev0, ev1])
    ta.propagate_until(5.0)
    gold = g["two_events_log"]
    for lane in range(2):
        mine = [(e, t) for b, e, t in log if b == lane]
        assert [e for e, t in mine] == [e for e, t in gold]
        # the v^2 - 1e-12 roots are ill-conditioned (nearly double): 1e-11 absolute
        assert max(abs(t - tg) for (e, t), (eg, tg) in zip(mine, gold)) < 1e-11
    n, mn, mx = g["two_events_propagate_until_5"]
    r = ta.propagate_res[0]
    assert r[3] == n and abs(r[1] - mn) < 1e-14 and abs(r[2] - mx) < 1e-14


def test_reference_kat_12_callbacks_per_lane():
    # /root/reference/heyoka/test.py:693-758
    x, v = hy.make_vars("x", "v")
    counter = [0] * 2
    cur_time = [0.0] * 2
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

    ta = hy.taylor_adaptive_batch(
        sys=common.pendulum_sys(), state=[[0.0, 0.001], [0.25, 0.2501]],
        nt_events=[hy.nt_event_batch(v * v - 1e-10, cb0), hy.nt_event_batch(v, cb1)])
    ta_id[0] = id(ta)
    ta.propagate_until([4.0, 4.0])
    assert all(r[0] == hy.taylor_outcome.time_limit for r in ta.propagate_res)
    assert counter == [12, 12]


@pytest.mark.parametrize("fp", [np.float64, np.float32])
def test_stopping_terminal_event_and_cooldowns(fp):
    # /root/reference/heyoka/_test_batch_integrator.py:555-600
    x, v = hy.make_vars("x", "v")
    ta = hy.taylor_adaptive_batch(
        sys=common.pendulum_sys(), state=np.array([[0.0, 0.001], [0.25, 0.2501]], dtype=fp),
        nt_events=[hy.nt_event_batch(v * v - 1e-6, lambda ta, t, d, b: None, fp_type=fp)],
        t_events=[hy.t_event_batch(v, fp_type=fp)], fp_type=fp)
    assert ta.with_events and len(ta.t_events) == 1 and len(ta.nt_events) == 1
    ta.propagate_until([fp(1e9), fp(1e9)])
    assert all(int(r[0]) == -1 for r in ta.propagate_res)
    assert np.all(np.abs(ta.state[1]) < (1e-12 if fp == np.float64 else 1e-5))  # stopped at v = 0
    assert ta.te_cooldowns[0][0] is not None and ta.te_cooldowns[1][0] is not None
    ta.reset_cooldowns(0)
    assert ta.te_cooldowns[0][0] is None and ta.te_cooldowns[1][0] is not None
    ta.reset_cooldowns()
    assert ta.te_cooldowns[0][0] is None and ta.te_cooldowns[1][0] is None


def test_device_resident_terminal_events_vs_oracle():
    # Stopping terminal events without callbacks run entirely on the device
    # (config 5 style: CR3BP with collision/escape spheres).
    mu = 0.01
    sys_ = common.cr3bp_sys(mu)
    x, y, z = hy.make_vars("x", "y", "z")
    evs = [(x - mu) ** 2 + y * y + z * z - 0.2 ** 2,
           (x - mu + 1.0) ** 2 + y * y + z * z - 0.2 ** 2,
           x * x + y * y + z * z - 1.3 ** 2]
    B = 16
    rng = np.random.default_rng(5)
    ic = np.array([-0.80, 0.0, 0.0, 0.0, -0.6276410653920693, 0.0])[:, None] + \
        np.concatenate([rng.uniform(-1e-2, 1e-2, (1, B)), np.zeros((3, B)),
                        rng.uniform(-1e-2, 1e-2, (1, B)), np.zeros((1, B))])
    ta = hy.taylor_adaptive_batch(sys_, ic, t_events=[hy.t_event_batch(e) for e in evs])
    ta.propagate_until(30.0)
    orc = NpTaylorBatch(sys_, ic, events=evs,
                        ev_spec=[{"dir": 0, "terminal": True, "cooldown": -1}] * 3)
    ro = orc.propagate_until(30.0)
    oc = np.array([int(r[0]) for r in ta.propagate_res])
    assert np.array_equal(oc, ro[0])
    assert np.any(oc > -10)  # some lanes did hit an event
    assert [r[3] for r in ta.propagate_res] == list(ro[3])
    assert np.max(np.abs(ta.time - orc.t_hi)) < 1e-11
    assert np.max(np.abs(ta.state - orc.state)) < 1e-9


def test_continuing_terminal_event_callback():
    # Event detection.ipynb cells 23-29: drag switched on/off at every v = 0.
    x, v = hy.make_vars("x", "v")
    times = []

    def t_cb(ta, d_sgn, bidx):
        ta.pars[0, bidx] = 1.0 if ta.pars[0, bidx] == 0 else 0.0
        times.append((bidx, float(ta.time[bidx])))
        return True

    sys_ = [(x, v), (v, -9.8 * hy.sin(x) - hy.par[0] * v)]
    ta = hy.taylor_adaptive_batch(sys_, np.array([[0.05, 0.05], [0.025, 0.025]]),
                                  t_events=[hy.t_event_batch(v, callback=t_cb)])
    # step by step until the event triggers: outcome is the event index (continuing)
    for _ in range(50):
        ta.step()
        if any(int(r[0]) == 0 for r in ta.step_res):
            break
    assert all(int(r[0]) == 0 for r in ta.step_res)
    assert np.all(ta.pars == 1.0)
    ta.propagate_until(10.0)
    assert np.all(ta.time == 10.0)
    assert all(r[0] == hy.taylor_outcome.time_limit for r in ta.propagate_res)
    assert len([t for b, t in times if b == 0]) >= 8

    def t_stop(ta, d_sgn, bidx):
        return False

    ta = hy.taylor_adaptive_batch(common.pendulum_sys(), np.array([[0.05, 0.05], [0.025, 0.025]]),
                                  t_events=[hy.t_event_batch(v, callback=t_stop)])
    ta.propagate_until(10.0)
    assert all(int(r[0]) == -1 for r in ta.propagate_res)

    def bad(ta, d_sgn, bidx):
        return "hello"

    ta = hy.taylor_adaptive_batch(common.pendulum_sys(), np.array([[0.05, 0.05], [0.025, 0.025]]),
                                  t_events=[hy.t_event_batch(v, callback=bad)])
    with pytest.raises(RuntimeError, match="in the construction of the return value of an event callback"):
        ta.propagate_until(10.0)
