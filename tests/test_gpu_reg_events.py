"""GPU: event detection on the register-resident kernels (csrc/hy_evtape.cuh) against the tape
interpreter (same integrator built with HY_CUDA_NO_REG_EVENTS=1) and the numpy oracle.
Reference semantics: /root/reference/heyoka/taylor_expose_events.cpp:185-317,
doc/notebooks/Event detection.ipynb."""

import os

import numpy as np
import pytest

import hy_b200 as hy
from hy_b200 import workloads as W
from oracle.np_oracle import NpTaylorBatch

pytestmark = pytest.mark.gpu


def _interp(make):
    os.environ["HY_CUDA_NO_REG_EVENTS"] = "1"
    try:
        ta = make()
        ta._ctx  # build the context while the switch is set
    finally:
        del os.environ["HY_CUDA_NO_REG_EVENTS"]
    return ta


def _cfg5(B, seed, R=0.012, Resc=5.0):
    mu = 0.01
    x, y, z = hy.make_vars("x", "y", "z")
    evs = [(x - mu) ** 2 + y * y + z * z - R ** 2,
           (x - mu + 1.0) ** 2 + y * y + z * z - R ** 2,
           x * x + y * y + z * z - Resc ** 2]
    rng = np.random.default_rng(seed)
    ic = np.array([-0.80, 0.0, 0.0, 0.0, -0.6276410653920693, 0.0])[:, None] * np.ones((1, B))
    ic[0] += rng.uniform(-1e-2, 1e-2, B)
    ic[4] += rng.uniform(-1e-2, 1e-2, B)
    return W.cr3bp_sys(mu), evs, ic


@pytest.mark.parametrize("R,Resc,T", [(0.012, 5.0, 40.0), (0.2, 1.3, 30.0)])
def test_cr3bp_terminal_events_register_kernel_vs_interpreter(R, Resc, T):
    B = 2048
    sys_, evs, ic = _cfg5(B, 77, R, Resc)
    mk = lambda: hy.taylor_adaptive_batch(sys_, ic, t_events=[hy.t_event_batch(e) for e in evs])
    a = mk()
    assert a._ctx.launch_info()["kernel_variant"] == 203       # events did not push it off the register kernel
    b = _interp(mk)
    assert b._ctx.launch_info()["kernel_variant"] == 0
    a.propagate_until(T)
    b.propagate_until(T)
    oa, ob = a.propagate_res_arrays[0], b.propagate_res_arrays[0]
    same = oa == ob
    # chaotic family: a handful of lanes may differ after many close encounters (the two paths sum the
    # event polynomials in a different order); everything else agrees tightly
    assert same.mean() > 0.995, same.mean()
    assert (oa > -10).sum() > 0
    dt = np.abs(a.time - b.time)[same]
    ds = np.abs(a.state - b.state).max(axis=0)[same]
    calm = ds < 1e-7
    assert calm.mean() > 0.97
    assert np.max(dt[calm]) < 1e-7
    # short horizon (no chaos yet): tight agreement on everything
    a2, b2 = mk(), _interp(mk)
    a2.propagate_until(3.0)
    b2.propagate_until(3.0)
    assert np.array_equal(a2.propagate_res_arrays[0], b2.propagate_res_arrays[0])
    assert np.array_equal(a2.propagate_res_arrays[3], b2.propagate_res_arrays[3])
    assert np.max(np.abs(a2.time - b2.time)) < 1e-13
    assert np.max(np.abs(a2.state - b2.state)) < 1e-11
    ma, mb = a2.propagate_res_arrays[1], b2.propagate_res_arrays[1]
    fin = np.isfinite(ma)
    assert np.array_equal(fin, np.isfinite(mb)) and np.max(np.abs(ma[fin] - mb[fin])) < 1e-13


def test_cr3bp_register_events_vs_numpy_oracle():
    B, T = 64, 30.0
    sys_, evs, ic = _cfg5(B, 5, 0.2, 1.3)
    ta = hy.taylor_adaptive_batch(sys_, ic, t_events=[hy.t_event_batch(e) for e in evs])
    assert ta._ctx.launch_info()["kernel_variant"] == 203
    ta.propagate_until(T)
    orc = NpTaylorBatch(sys_, ic, events=evs, ev_spec=[{"dir": 0, "terminal": True, "cooldown": -1}] * 3)
    ro = orc.propagate_until(T)
    oc = ta.propagate_res_arrays[0]
    assert np.array_equal(oc, ro[0]) and np.any(oc > -10)
    assert list(ta.propagate_res_arrays[3]) == list(ro[3])
    assert np.max(np.abs(ta.time - orc.t_hi)) < 1e-11
    assert np.max(np.abs(ta.state - orc.state)) < 1e-9
    # stopped lanes sit on the surface of the event that fired
    st = ta.state
    g = [(st[0] - 0.01) ** 2 + st[1] ** 2 + st[2] ** 2 - 0.04,
         (st[0] + 0.99) ** 2 + st[1] ** 2 + st[2] ** 2 - 0.04,
         st[0] ** 2 + st[1] ** 2 + st[2] ** 2 - 1.69]
    for e in range(3):
        m = oc == -e - 1
        if np.any(m):
            assert np.max(np.abs(g[e][m])) < 1e-13


def test_nbody_register_kernel_with_nt_events_and_callbacks():
    # non-terminal events with callbacks on the 6-body register kernel: Jupiter crossing y = 0
    sys_ = W.oss_sys()
    ic = W.oss_ensemble(6)
    y1 = hy.expression("y_1")
    vx5 = hy.expression("vx_5")

    def mk(log):
        def cb(ta, t, d_sgn, bidx):
            log.append((bidx, float(t), d_sgn))

        return hy.taylor_adaptive_batch(sys_, ic, nt_events=[
            hy.nt_event_batch(y1, cb), hy.nt_event_batch(vx5 * vx5 - 1e-2, cb, direction=hy.event_direction.positive)])

    la, lb = [], []
    a = mk(la)
    assert a._ctx.launch_info()["kernel_variant"] == 6
    b = _interp(lambda: mk(lb))
    assert b._ctx.launch_info()["kernel_variant"] == 0
    a.propagate_until(60.0)
    b.propagate_until(60.0)
    assert len(la) == len(lb) and len(la) >= 6 * 9          # ~5 Jupiter periods -> 10 crossings per lane
    ka = sorted(la, key=lambda r: (r[0], r[1]))
    kb = sorted(lb, key=lambda r: (r[0], r[1]))
    assert [(r[0], r[2]) for r in ka] == [(r[0], r[2]) for r in kb]
    assert np.max(np.abs(np.array([r[1] for r in ka]) - np.array([r[1] for r in kb]))) < 1e-11
    assert np.max(np.abs(a.state - b.state) / np.maximum(1, np.abs(b.state))) < 1e-12
    assert np.array_equal(a.propagate_res_arrays[3], b.propagate_res_arrays[3])
    # Jupiter's period: consecutive same-direction crossings are ~11.86 yr apart
    t0 = [r[1] for r in ka if r[0] == 0 and r[2] > 0 and abs(r[1]) > 0]
    ty = [t for t in t0]
    assert len(ty) >= 2


def test_register_events_with_recurrences_and_time():
    # event functions with sqrt / sin(time) (recurrent ops in the event tape) and a bare state variable
    mu = 0.01
    sys_ = W.cr3bp_sys(mu)
    x, y, z, px = hy.make_vars("x", "y", "z", "px")
    evs = [hy.sqrt(x * x + y * y) - (0.9 + 0.05 * hy.sin(hy.time)), px]
    ic = W.cr3bp_ensemble(32)
    hits = {}

    def mk(key):
        hits[key] = []

        def cb(ta, t, d_sgn, bidx):
            hits[key].append((bidx, float(t), d_sgn))

        return hy.taylor_adaptive_batch(sys_, ic, nt_events=[hy.nt_event_batch(e, cb) for e in evs])

    a = mk("a")
    assert a._ctx.launch_info()["kernel_variant"] == 203
    b = _interp(lambda: mk("b"))
    a.propagate_until(10.0)
    b.propagate_until(10.0)
    ka = sorted(hits["a"], key=lambda r: (r[0], r[1]))
    kb = sorted(hits["b"], key=lambda r: (r[0], r[1]))
    assert len(ka) == len(kb) and len(ka) > 32
    assert [(r[0], r[2]) for r in ka] == [(r[0], r[2]) for r in kb]
    assert np.max(np.abs(np.array([r[1] for r in ka]) - np.array([r[1] for r in kb]))) < 1e-11
    assert np.max(np.abs(a.state - b.state)) < 1e-11


def test_generated_event_code_vs_interpreted_event_tape():
    # Large batches get the event functions as generated code (csrc/hy_jit.hpp, EvtGen: NVRTC build of the
    # FX kernel); small ones the interpreted event tape.  Same trajectories, same events.  The generated code
    # chains every sum like the interpreted tape and its (different, cheaper) enclosure test is conservative
    # as well, so config 5's chaotic family comes out bit for bit - outcomes, event times, states, step counts.
    B = 1024
    sys_, evs, ic = _cfg5(B, 123, 0.2, 1.3)
    mk = lambda: hy.taylor_adaptive_batch(sys_, ic, t_events=[hy.t_event_batch(e) for e in evs])
    a = mk()                                    # B < 4096: interpreted event tape
    os.environ["HY_CUDA_JIT_EVT"] = "2"
    try:
        b = mk()
        b._ctx
    finally:
        del os.environ["HY_CUDA_JIT_EVT"]
    assert a._ctx.launch_info()["kernel_variant"] == 203 and b._ctx.launch_info()["kernel_variant"] == 203
    a.propagate_until(30.0)
    b.propagate_until(30.0)
    assert (a.propagate_res_arrays[0] > -10).sum() > 50
    for u, v in zip(a.propagate_res_arrays, b.propagate_res_arrays):
        assert np.array_equal(u, v)
    assert np.array_equal(a.time, b.time) and np.array_equal(a.state, b.state)


def _gen_vs_interp_events(sys_, ic, evs, variant, T):
    """The same event-carrying system with the event tape interpreted (HY_CUDA_JIT_EVT=0) and as generated
    code (=2: products spread over the lanes of the group, shared sub-expressions computed once, "x + c"
    folded into its readers).  The generated code chains every sum like the interpreter: bit for bit."""
    logs = {}

    def mk(key, mode):
        logs[key] = []

        def cb(ta, t, d_sgn, bidx):
            logs[key].append((bidx, float(t), d_sgn))

        os.environ["HY_CUDA_JIT_EVT"] = mode
        try:
            ta = hy.taylor_adaptive_batch(sys_, ic, nt_events=[hy.nt_event_batch(e, cb) for e in evs])
            ta._ctx
        finally:
            del os.environ["HY_CUDA_JIT_EVT"]
        assert ta._ctx.launch_info()["kernel_variant"] == variant
        return ta

    a, b = mk("a", "0"), mk("b", "2")
    a.propagate_until(T)
    b.propagate_until(T)
    ka = sorted(logs["a"], key=lambda r: (r[0], r[1]))
    kb = sorted(logs["b"], key=lambda r: (r[0], r[1]))
    assert len(ka) > 2 * ic.shape[1]
    assert ka == kb
    assert np.array_equal(a.state, b.state) and np.array_equal(a.time, b.time)
    assert np.array_equal(a.propagate_res_arrays[3], b.propagate_res_arrays[3])


def test_generated_events_nbody_lane_parallel_products():
    # 16 lanes per trajectory: squares of every-order differences (unit-stride operands), products of
    # state jets, products of folded "x + c" operands, a repeated square
    sys_ = W.oss_sys()
    ic = W.oss_ensemble(8)
    v = lambda s: hy.expression(s)
    dx, dy, dz = v("x_1") - v("x_2"), v("y_1") - v("y_2"), v("z_1") - v("z_2")
    evs = [dx * dx + dy * dy + dz * dz - 60.0,
           v("x_1") * v("vy_1") - v("y_1") * v("vx_1") - 2.0,
           (v("x_1") + 0.5) * (v("y_1") - 0.25),
           v("y_1"),
           v("vx_5") * v("vx_5") - 1e-2,
           dx * dx - 4.0]
    _gen_vs_interp_events(sys_, ic, evs, 6, 60.0)


def test_generated_events_cr3bp_recurrences_and_products():
    # recurrent ops (sqrt, sin(time)) keep the generic interval pass and the every-order pass on lane 0
    sys_ = W.cr3bp_sys(0.01)
    x, y, z, px = hy.make_vars("x", "y", "z", "px")
    evs = [hy.sqrt(x * x + y * y) - (0.9 + 0.05 * hy.sin(hy.time)), px, (x - 0.3) * (y + 0.1), (x - 0.3) ** 2 + y * y - 0.5]
    _gen_vs_interp_events(sys_, W.cr3bp_ensemble(32), evs, 203, 10.0)
