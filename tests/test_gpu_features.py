"""GPU: tc / dense output / continuous output / propagate_grid through the
Python front end, against the notebook golden values and the numpy oracle.
Mirrors /root/reference/heyoka/_test_batch_integrator.py:340-413 (dense),
:692-857 (grid) and test.py:1414-1771 (continuous output)."""

import json
import os

import numpy as np
import pytest

import hy_b200 as hy
from oracle.np_oracle import NpTaylorBatch

import common

pytestmark = pytest.mark.gpu
G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "notebook_golden.json")))


def _forced():
    g = G["batch_forced_pendulum"]
    return g, hy.taylor_adaptive_batch(common.forced_pendulum_sys(), np.array(g["ic"]),
                                       pars=np.array(g["pars"]))


def test_tc_and_dense_output_golden_A9():
    g, ta = _forced()
    ta.step(write_tc=True)
    tc = ta.tc
    assert tc.shape == (2, 21, 4) and not tc.flags.writeable
    assert np.all(tc[:, 0, :] == np.array(g["ic"]))
    assert np.max(np.abs(tc[0, 2] - np.array(g["tc_x_order2"]))) < 1e-9
    assert np.max(np.abs(tc[0, 3] - np.array(g["tc_x_order3"]))) < 1e-9
    d = ta.update_d_output(g["d_output_at"])
    assert d.shape == (2, 4) and not d.flags.writeable
    assert np.max(np.abs(d - np.array(g["d_output"]))) < 1e-8
    assert ta.d_output is not None and np.all(ta.d_output == d)


@pytest.mark.parametrize("fp", [np.float64, np.float32])
def test_dense_output_self_consistency(fp):
    # _test_batch_integrator.py:393-413
    ic = common.PEND_IC.astype(fp)
    ta = hy.taylor_adaptive_batch(common.pendulum_sys(), ic, fp_type=fp)
    ta.step(write_tc=True)
    eps = np.finfo(fp).eps
    d = ta.update_d_output(ta.time.copy())
    assert np.all(np.abs(d - ta.state) <= 10 * eps * np.maximum(1, np.abs(ta.state)))
    d = ta.update_d_output(fp(0.0), rel_time=True)
    assert np.all(np.abs(d - ta.state) <= 10 * eps * np.maximum(1, np.abs(ta.state)))
    d = ta.update_d_output([fp(0.0)] * 4, rel_time=True)
    assert np.all(np.abs(d - ta.state) <= 10 * eps * np.maximum(1, np.abs(ta.state)))
    # start of the step: tc[:, 0]
    d = ta.update_d_output(-ta.last_h.copy(), rel_time=True)
    assert np.all(np.abs(d - ic) <= 10 * eps)


def test_continuous_output_harmonic_A6():
    g = G["harmonic"]
    x, v = hy.make_vars("x", "v")
    B = 3
    ic = np.array([[0.0] * B, [1.0] * B])
    ta = hy.taylor_adaptive_batch([(x, v), (v, -x)], ic)
    c_out, _ = ta.propagate_until(10.0, c_output=True)
    assert c_out.n_steps == g["c_output_steps_to_10"] and c_out.batch_size == B
    assert c_out.times.shape == (c_out.n_steps + 1, B)
    assert c_out.tcs.shape == (c_out.n_steps, 2, ta.order + 1, B)
    assert np.all(c_out.times[0] == 0.0) and np.all(c_out.times[-1] == 10.0)
    r = c_out(5.0)
    assert r.shape == (2, B)
    assert np.max(np.abs(r[:, 0] - np.array(g["c_out_5"]))) < 1e-8
    tq = np.repeat(np.arange(1.0, 6.0), B).reshape(5, B)
    rr = c_out(tq)
    assert rr.shape == (5, 2, B)
    assert np.max(np.abs(rr[:, :, 1] - np.array(g["c_out_1_to_5"]))) < 1e-8
    # exact solution over a fine grid
    tg = np.linspace(0, 10, 257)
    rr = c_out(np.repeat(tg, B).reshape(-1, B))
    assert np.max(np.abs(rr[:, 0, 0] - np.sin(tg))) < 1e-14
    assert np.max(np.abs(rr[:, 1, 2] - np.cos(tg))) < 1e-14
    b0, b1 = c_out.bounds
    assert np.all(b0 == 0.0) and np.all(b1 == 10.0)
    assert c_out(np.zeros((0, B))).shape == (0, 2, B)
    with pytest.raises(ValueError):
        c_out(np.zeros(B + 1))
    with pytest.raises(ValueError):
        hy.continuous_output_batch_dbl()(0.0)


def test_continuous_output_vs_oracle_ragged():
    # lanes with different step counts; compare against the numpy oracle's record
    sys_ = common.pendulum_sys()
    ic = common.PEND_IC
    ta = hy.taylor_adaptive_batch(sys_, ic)
    tf = [3.0, 5.0, 7.0, 9.0]
    c_out, _ = ta.propagate_until(tf, c_output=True)
    orc = NpTaylorBatch(sys_, ic)
    res, rec = orc.propagate_until_recorded(tf)
    ns = [len(r) for r in rec]
    assert [r[3] for r in ta.propagate_res] == ns
    assert c_out.n_steps == max(ns)
    times = c_out.times
    for l in range(4):
        assert np.all(times[ns[l] + 1:, l] == times[ns[l], l]) and np.all(np.isfinite(times[:, l]))   # end time repeated
        assert np.all(np.isnan(c_out.tcs[ns[l]:, :, :, l]))
    tq = np.array([[0.5, 0.5, 0.5, 0.5], [2.9, 4.9, 6.9, 8.9], [1.0, 2.0, 3.0, 4.0]])
    out = c_out(tq)
    for q in range(3):
        for l in range(4):
            ref = NpTaylorBatch.eval_record(rec[l], tq[q, l])
            assert np.max(np.abs(out[q, :, l] - ref)) < 1e-12


def test_propagate_grid_golden_and_oracle():
    g, ta = _forced()
    grid = np.repeat(np.linspace(0, 100, 1000), 4).reshape(1000, 4)
    cb, out = ta.propagate_grid(grid)
    assert cb is None and out.shape == (1000, 2, 4)
    res = np.array(g["grid_0_100_x1000"]["res"])
    pr = ta.propagate_res
    assert [r[3] for r in pr] == list(res[:, 2].astype(int))
    assert np.max(np.abs(np.array([r[1] for r in pr]) - res[:, 0]) / res[:, 0]) < 1e-10
    assert np.max(np.abs(np.array([r[2] for r in pr]) - res[:, 1]) / res[:, 1]) < 1e-10
    assert np.all(ta.time == 100.0)
    assert np.all(out[0] == np.array(g["ic"]))
    assert np.max(np.abs(out[-1] - ta.state)) < 1e-13
    # against the oracle on a short grid with per-lane different times
    ta2 = hy.taylor_adaptive_batch(common.forced_pendulum_sys(), np.array(g["ic"]),
                                   pars=np.array(g["pars"]))
    grid2 = np.linspace(0, 1, 11)[:, None] * np.array([5.0, 6.0, 7.0, 8.0])[None, :]
    _, out2 = ta2.propagate_grid(grid2)
    orc = NpTaylorBatch(common.forced_pendulum_sys(), np.array(g["ic"]), pars=np.array(g["pars"]))
    _, oref = orc.propagate_grid(grid2)
    assert np.max(np.abs(out2 - oref)) < 1e-12


def test_propagate_grid_scalar_pendulum_A4():
    g = G["pendulum_scalar"]
    ic = np.array(g["ic"])[:, None] * np.ones((1, 2))
    ta = hy.taylor_adaptive_batch(common.pendulum_sys(), ic)
    grid = np.repeat(np.linspace(0, 1, 11), 2).reshape(11, 2)
    _, out = ta.propagate_grid(grid)
    gg = g["grid_0_1"]
    r = ta.propagate_res[0]
    assert r[3] == gg["steps"]
    assert abs(r[1] - gg["min_h"]) < 1e-15 and abs(r[2] - gg["max_h"]) < 1e-15
    assert np.max(np.abs(out[:, :, 1] - np.array(gg["out"]))) < 1e-8


def test_propagate_grid_errors():
    ta = hy.taylor_adaptive_batch(common.pendulum_sys(), common.PEND_IC)
    with pytest.raises(ValueError, match="number of dimensions is 2"):
        ta.propagate_grid([0.0, 1.0])
    with pytest.raises(ValueError):
        ta.propagate_grid(np.zeros((3, 5)))
    with pytest.raises(ValueError):
        ta.propagate_grid(np.zeros((0, 4)))


@pytest.mark.parametrize("env", [{"HY_CUDA_GROUP": "1"}, {"HY_CUDA_GROUP": "4"}, {"HY_CUDA_GROUP": "16"},
                                 {"HY_CUDA_FORCE_GLOBAL_WS": "1"}])
def test_launch_geometries_agree_bitwise(env):
    # The scheduler spreads the tape over G lanes per trajectory (1, 4 or 16) and keeps the jets in shared
    # memory, or - for systems too large for it - in global memory: the arithmetic does not depend on
    # the geometry, so every choice must reproduce the default one bit for bit.
    import os
    from hy_b200 import workloads as W

    def make(sys_, ic, extra):
        old = {k: os.environ.get(k) for k in list(extra) + ["HY_CUDA_NO_CR3BP_REG"]}
        os.environ.update(extra)
        os.environ["HY_CUDA_NO_CR3BP_REG"] = "1"  # the tape interpreter is what this test is about
        try:
            return hy.taylor_adaptive_batch(sys_, ic)
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v

    B = 150
    for sys_, ic, t_end in ((W.cr3bp_sys(0.01), W.cr3bp_ensemble(B), 6.0),
                            (W.pendulum_sys(), np.linspace(-1.0, 1.0, 2 * B).reshape(2, B), 15.0)):
        a = make(sys_, ic, {})
        b = make(sys_, ic, env)
        lb = b._ctx.launch_info()
        if "HY_CUDA_GROUP" in env:
            assert lb["group"] == int(env["HY_CUDA_GROUP"]) and lb["kernel_variant"] == 0
        else:
            assert lb["ws_in_smem"] == 0 and lb["kernel_variant"] == 0
        a.step(write_tc=True)
        b.step(write_tc=True)
        assert np.array_equal(a.tc, b.tc)
        tf = np.linspace(0.5 * t_end, t_end, B)
        a.propagate_until(tf)
        b.propagate_until(tf)
        assert np.array_equal(a.state, b.state)
        assert a.propagate_res == b.propagate_res
