"""GPU: the run-time compiled kernels (csrc/hy_jit.hpp, kernel_variant 1000) against the tape
interpreter (same integrator built with compact_mode=True) - bit for bit, every op kind and every
API feature that runs inside the kernel - and against the C oracle.

Reference: the reference JIT-compiles every system in the constructor
(/root/reference/heyoka/expose_batch_integrators.cpp:166-208); `compact_mode` there trades
code-generation time for run time (:198), here it selects the interpreter.
"""

import os

import numpy as np
import pytest

import hy_b200 as hy
from hy_b200 import decompose as D
from hy_b200 import workloads as W
from oracle.c_oracle import COracle

pytestmark = pytest.mark.gpu

JIT = 1000


def _forced(make):
    """Build with HY_CUDA_JIT=1 (every tape, however small, gets a compiled kernel)."""
    old = os.environ.get("HY_CUDA_JIT")
    os.environ["HY_CUDA_JIT"] = "1"
    try:
        ta = make()
        ta._ctx
    finally:
        if old is None:
            del os.environ["HY_CUDA_JIT"]
        else:
            os.environ["HY_CUDA_JIT"] = old
    return ta


def _pair(sys_, ic, **kw):
    a = _forced(lambda: hy.taylor_adaptive_batch(sys_, ic, **kw))
    b = hy.taylor_adaptive_batch(sys_, ic, compact_mode=True, **kw)
    assert a._ctx.launch_info()["kernel_variant"] == JIT, a._ctx.launch_info()
    assert b._ctx.launch_info()["kernel_variant"] == 0, b._ctx.launch_info()
    return a, b


def _same(a, b):
    assert np.array_equal(a.state, b.state)
    assert np.array_equal(a.time, b.time)
    for x, y in zip(a.propagate_res_arrays, b.propagate_res_arrays):
        assert np.array_equal(x, y)


def test_config4_default_is_compiled_and_bitwise_equal_to_interpreter():
    # 42 state variables, 191 ops: the default build of a large tape is the compiled kernel
    vs = hy.var_ode_sys(W.kepler_j2_sys(), hy.var_args.vars)
    ic = W.kepler_j2_ensemble(96, seed=5)
    a = hy.taylor_adaptive_batch(vs, ic)
    b = hy.taylor_adaptive_batch(vs, ic, compact_mode=True)
    li = a._ctx.launch_info()
    assert li["kernel_variant"] == JIT and li["group"] == 1 and li["ws_in_smem"] == 0
    assert b._ctx.launch_info()["kernel_variant"] == 0
    a.propagate_until(4000.0)
    b.propagate_until(4000.0)
    _same(a, b)
    # and the oracle (step counts exact, state to 1e-11)
    orc = COracle(D.decompose(vs.sys, a.order), W_full(vs, ic))
    oc, mn, mx, ns, _ = orc.propagate_until(4000.0)
    assert list(a.propagate_res_arrays[3]) == list(ns)
    err = np.max(np.abs(a.state - orc.state) / np.maximum(1.0, np.abs(orc.state)))
    assert err < 1e-11, err


def W_full(vs, ic):
    """Initial conditions of a variational system: the state plus the identity sensitivities."""
    ta = hy.taylor_adaptive_batch(vs, ic, compact_mode=True)
    return ta.state.copy()


@pytest.mark.parametrize("fp", [np.float64, np.float32])
def test_every_op_kind_bitwise(fp):
    # sin/cos, exp, log, sqrt, pow, div, square, products, time, runtime parameters
    x, v, s = hy.make_vars("x", "v", "s")
    sys_ = [(x, v),
            (v, hy.cos(hy.time) - hy.par[0] * v - hy.sin(x) + 0.01 * hy.exp(-s * s) / (1.0 + x * x)),
            (s, hy.log(2.0 + hy.cos(x)) * hy.sqrt(1.0 + v * v) - hy.par[1] * s + (1.5 + x * x) ** -1.5 + (x * v) * (s * v))]
    rng = np.random.default_rng(3)
    ic = rng.uniform(-0.5, 0.5, (3, 70)).astype(fp)
    pars = np.stack([np.full(70, 0.1), np.linspace(0.2, 0.4, 70)]).astype(fp)
    a, b = _pair(sys_, ic, fp_type=fp, pars=pars)
    for ta in (a, b):
        ta.propagate_until(fp(6.0))
    _same(a, b)
    # single steps with the Taylor coefficients, backward too
    for ta in (a, b):
        ta.step(write_tc=True)
        ta.step_backward()
    assert np.array_equal(a.tc, b.tc)
    _same(a, b)
    orc = COracle(D.decompose(sys_, a.order), ic, pars=pars, fp_type=fp)
    orc.propagate_until(fp(6.0))


def test_ragged_times_grid_and_continuous_output_bitwise():
    sys_ = W.forced_pendulum_sys()
    B = 45
    ic = np.stack([np.linspace(0.0, 1.0, B), np.linspace(0.2, 0.3, B)])
    pars = np.full((1, B), 0.05)
    a, b = _pair(sys_, ic, pars=pars)
    tf = np.linspace(2.0, 9.0, B)
    for ta in (a, b):
        ta.propagate_until(tf, max_delta_t=0.7)
    _same(a, b)
    grid = np.linspace(9.0, 14.0, 11)[:, None] * np.ones((1, B)) + np.linspace(0, 0.3, B)[None, :]
    ga = a.propagate_grid(grid)[1]
    gb = b.propagate_grid(grid)[1]
    assert np.array_equal(ga, gb)
    _same(a, b)
    ca, _ = a.propagate_for(3.0, c_output=True)
    cb, _ = b.propagate_for(3.0, c_output=True)
    _same(a, b)
    tq = np.linspace(14.4, 17.0, 9)[:, None] * np.ones((1, B))
    assert np.array_equal(ca(tq), cb(tq))


def test_events_bitwise_and_high_accuracy():
    # terminal + non-terminal events share u-variables with the ODE; compensated update
    x, v = hy.make_vars("x", "v")
    sys_ = [(x, v), (v, -9.8 * hy.sin(x))]
    B = 40
    ic = np.stack([np.linspace(0.1, 1.2, B), np.linspace(-0.3, 0.3, B)])

    def mk(**kw):
        return dict(t_events=[hy.t_event_batch(v * v - 1.0, direction=hy.event_direction.positive)],
                    nt_events=[hy.nt_event_batch(x, lambda ta, t, d, i: None)], high_accuracy=True, **kw)

    a = _forced(lambda: hy.taylor_adaptive_batch(sys_, ic, **mk()))
    b = hy.taylor_adaptive_batch(sys_, ic, compact_mode=True, **mk())
    assert a._ctx.launch_info()["kernel_variant"] == JIT
    for ta in (a, b):
        ta.propagate_until(8.0)
    _same(a, b)
    assert (a.propagate_res_arrays[0] > -10).any()  # some lanes stopped on the terminal event


def test_parametric_masses_nbody_gets_the_register_kernel():
    # An N-body system with the masses as runtime parameters (model.nbody(masses=[par[i], ...])): the
    # matcher accepts parameter-scaled acceleration terms and hy_create BUILDS the register-resident
    # kernel for it (NVRTC, HY_NBR_PAR: the body lanes multiply their coefficients by the trajectory's
    # parameters).  Bit for bit the tape interpreter's result, the oracle's to 1e-12 - and every lane has
    # its own masses.
    from hy_b200 import model

    sys_ = model.nbody(6, masses=[hy.par[i] for i in range(6)], Gconst=W.OSS_G)
    B = 64
    ic = W.oss_ensemble(B)
    pars = W.OSS_MASSES[:, None] * (1.0 + 0.01 * np.linspace(-1, 1, B))[None, :]
    a = hy.taylor_adaptive_batch(sys_, ic, pars=pars)
    assert a._ctx.launch_info()["kernel_variant"] == 6, a._ctx.launch_info()
    os.environ["HY_CUDA_NO_NBODY_REG"] = "1"
    try:
        b = hy.taylor_adaptive_batch(sys_, ic, pars=pars)
        b._ctx
    finally:
        del os.environ["HY_CUDA_NO_NBODY_REG"]
    assert b._ctx.launch_info()["kernel_variant"] == 0 and b._ctx.launch_info()["group"] == 16
    for ta in (a, b):
        ta.propagate_until(30.0)
    _same(a, b)
    # the FX build too (continuous output) and a copy on the same kernel
    import copy

    c = copy.deepcopy(a)
    ca, _ = a.propagate_for(5.0, c_output=True)
    cb, _ = b.propagate_for(5.0, c_output=True)
    c.propagate_for(5.0)
    _same(a, b)
    _same(c, b)
    tq = np.linspace(30.5, 34.5, 5)[:, None] * np.ones((1, B))
    assert np.array_equal(ca(tq), cb(tq))
    orc = COracle(D.decompose(sys_, a.order), ic, pars=pars)
    oc, mn, mx, ns, _ = orc.propagate_until(35.0)
    err = np.max(np.abs(a.state - orc.state) / np.maximum(1.0, np.abs(orc.state)))
    assert err < 1e-12, err
    # changing a lane's masses changes that lane only
    a2 = hy.taylor_adaptive_batch(sys_, ic, pars=pars)
    a3 = hy.taylor_adaptive_batch(sys_, ic, pars=pars)
    a2.pars[1, 7] *= 1.5
    a2.propagate_until(35.0)
    a3.propagate_until(35.0)
    diff = np.abs(a2.state - a3.state).max(axis=0)
    assert diff[7] > 1e-9 and np.all(np.delete(diff, 7) == 0.0)


def _nbody_ic(nb, B, fp=np.float64):
    """The outer Solar System plus nb - 6 light bodies on wide orbits."""
    extra = nb - 6
    masses = list(W.OSS_MASSES) + [1e-9 * (1 + e) for e in range(extra)]
    add = []
    for e in range(extra):
        r = 45.0 + 7.0 * e
        add.append(np.array([r, 0.0, 0.3, 0.0, 2 * np.pi / np.sqrt(r), 0.0])[:, None] * np.ones((1, B)))
    ic = np.concatenate([W.oss_ensemble(B)] + add, axis=0).astype(fp)
    from hy_b200 import model

    return model.nbody(nb, masses=masses, Gconst=W.OSS_G), ic


@pytest.mark.parametrize("nb,fp", [(7, np.float64), (8, np.float64), (8, np.float32)])
def test_seven_and_eight_bodies_on_32_lane_groups(nb, fp):
    # 21 / 28 pairs do not fit a 16-lane group: hy_create builds the register-resident kernel on
    # 32-lane groups (one trajectory per warp; NVRTC, HY_NBR_G32).  Same arithmetic as the interpreter's
    # fused pair op + LINCOMB + SVD: bit for bit, every API feature that runs in the kernel included.
    B = 37
    sys_, ic = _nbody_ic(nb, B, fp)
    a = hy.taylor_adaptive_batch(sys_, ic, fp_type=fp)
    li = a._ctx.launch_info()
    assert li["kernel_variant"] == nb and li["group"] == 32, li
    os.environ["HY_CUDA_NO_NBODY_REG"] = "1"
    try:
        b = hy.taylor_adaptive_batch(sys_, ic, fp_type=fp)
        b._ctx
    finally:
        del os.environ["HY_CUDA_NO_NBODY_REG"]
    assert b._ctx.launch_info()["kernel_variant"] == 0
    for ta in (a, b):
        ta.step(write_tc=True)
    assert np.array_equal(a.tc, b.tc)
    tf = np.linspace(20.0, 40.0, B).astype(fp)
    for ta in (a, b):
        ta.propagate_until(tf)
    _same(a, b)
    grid = (np.linspace(40.0, 44.0, 5)[:, None] * np.ones((1, B))).astype(fp)
    assert np.array_equal(a.propagate_grid(grid)[1], b.propagate_grid(grid)[1])
    ca, _ = a.propagate_for(fp(3.0), c_output=True)
    cb, _ = b.propagate_for(fp(3.0), c_output=True)
    _same(a, b)
    tq = (np.linspace(44.2, 46.8, 4)[:, None] * np.ones((1, B))).astype(fp)
    assert np.array_equal(ca(tq), cb(tq))
    if fp == np.float64:
        orc = COracle(D.decompose(sys_, a.order), ic)
        c = hy.taylor_adaptive_batch(sys_, ic)
        c.propagate_until(25.0)
        oc, mn, mx, ns, _ = orc.propagate_until(25.0)
        assert list(c.propagate_res_arrays[3]) == list(ns)
        err = np.max(np.abs(c.state - orc.state) / np.maximum(1.0, np.abs(orc.state)))
        assert err < 1e-12, err


def test_eight_bodies_with_parametric_masses_and_an_event():
    # both build switches at once (HY_NBR_PAR + HY_NBR_G32), plus a terminal event on the event tape
    from hy_b200 import model

    B = 20
    _, ic = _nbody_ic(8, B)
    masses = list(W.OSS_MASSES) + [1e-9, 2e-9]
    sys_ = model.nbody(8, masses=[hy.par[i] for i in range(8)], Gconst=W.OSS_G)
    pars = np.array(masses)[:, None] * np.ones((1, B))
    x6 = sys_[36][0]  # x of body 6
    ev = lambda: [hy.t_event_batch(x6 - 44.0, direction=hy.event_direction.negative)]
    a = hy.taylor_adaptive_batch(sys_, ic, pars=pars, t_events=ev())
    assert a._ctx.launch_info()["kernel_variant"] == 8
    os.environ["HY_CUDA_NO_NBODY_REG"] = "1"
    try:
        b = hy.taylor_adaptive_batch(sys_, ic, pars=pars, t_events=ev())
        b._ctx
    finally:
        del os.environ["HY_CUDA_NO_NBODY_REG"]
    assert b._ctx.launch_info()["kernel_variant"] == 0
    for ta in (a, b):
        ta.propagate_until(60.0)
    assert np.array_equal(a.propagate_res_arrays[0], b.propagate_res_arrays[0])
    assert (a.propagate_res_arrays[0] > -10).all()          # every lane stops on the event
    assert np.max(np.abs(a.time - b.time)) < 1e-9
    assert np.max(np.abs(a.state - b.state) / np.maximum(1.0, np.abs(b.state))) < 1e-9


def test_compiled_kernel_in_shared_memory_and_clone():
    # a small tape: the interleaved workspace of a whole CTA fits in shared memory
    sys_ = W.forced_pendulum_sys()
    B = 300
    ic = np.stack([np.linspace(0.0, 1.0, B), np.linspace(0.2, 0.3, B)])
    pars = np.full((1, B), 0.05)
    a, b = _pair(sys_, ic, pars=pars)
    assert a._ctx.launch_info()["ws_in_smem"] == 1
    import copy

    c = copy.deepcopy(a)  # hy_clone: the compiled kernel is loaded again for the copy
    for ta in (a, b, c):
        ta.propagate_until(3.0)
    _same(a, b)
    _same(c, b)
    assert c._ctx.launch_info()["kernel_variant"] == JIT


def test_compiled_kernel_global_workspace_mid_size_tape():
    # CR3BP with the register kernel switched off: 24 ops, the jets stream from global memory
    sys_ = W.cr3bp_sys(0.01)
    ic = W.cr3bp_ensemble(200)
    os.environ["HY_CUDA_NO_CR3BP_REG"] = "1"
    try:
        a, b = _pair(sys_, ic)
    finally:
        del os.environ["HY_CUDA_NO_CR3BP_REG"]
    assert a._ctx.launch_info()["ws_in_smem"] == 0
    for ta in (a, b):
        ta.propagate_until(3.0)
    _same(a, b)


def test_without_nvrtc_the_interpreter_takes_over(tmp_path):
    # no compiler and an empty cache: hy_create must not fail - the tape interpreter runs the system
    # (a subprocess: the NVRTC handle is loaded once per process)
    import subprocess
    import sys

    code = """
import numpy as np, hy_b200 as hy
from hy_b200 import workloads as W
vs = hy.var_ode_sys(W.kepler_j2_sys(), hy.var_args.vars)
ta = hy.taylor_adaptive_batch(vs, W.kepler_j2_ensemble(8))
li = ta._ctx.launch_info()
assert li["kernel_variant"] == 0, li
ta.propagate_until(500.0)
print("fallback ok", li["group"])
"""
    env = dict(os.environ, HY_CUDA_NVRTC_LIB="/nonexistent/libnvrtc.so", HY_CUDA_JIT_CACHE=str(tmp_path),
               PYTHONPATH=os.pathsep.join(sys.path))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "fallback ok" in out.stdout, out.stderr[-2000:]


def test_thread_ensemble_on_a_compiled_kernel():
    # ensemble iterations clone the template's context (hy_clone loads the compiled image again) and
    # run concurrently from a thread pool: same results as serial runs of the same initial conditions
    vs = hy.var_ode_sys(W.kepler_j2_sys(), hy.var_args.vars)
    B, n_iter = 32, 6
    ics = [W.kepler_j2_ensemble(B, seed=100 + i) for i in range(n_iter)]
    tmpl = hy.taylor_adaptive_batch(vs, ics[0])
    assert tmpl._ctx.launch_info()["kernel_variant"] == JIT

    def gen(ta, i):
        ta.state[:6] = ics[i]
        return ta

    res = hy.ensemble_propagate_until_batch(tmpl, 2500.0, n_iter, gen)
    for i, r in enumerate(res):
        ref = hy.taylor_adaptive_batch(vs, ics[i])
        ref.propagate_until(2500.0)
        assert np.array_equal(r[0].state, ref.state), i
        assert np.array_equal(r[0].propagate_res_arrays[3], ref.propagate_res_arrays[3])
