"""GPU tests of the register-resident N-body kernel (csrc/hy_nbody_reg.cuh).

hy_create recognises the tape of model.nbody(6) and runs the jets in registers;
HY_CUDA_NO_NBODY_REG=1 forces the tape interpreter on the same tape.  Both
implement the same recurrences in the same term order, so they must agree BIT
FOR BIT; the oracle comparison (1e-12 relative, BASELINE.json) pins both.
"""

import os

import numpy as np
import pytest

import hy_b200 as hy
from hy_b200 import decompose as D
from oracle.c_oracle import COracle

import common

pytestmark = pytest.mark.gpu


def _make(sys_, ic, interp=False, **kw):
    old = os.environ.get("HY_CUDA_NO_NBODY_REG")
    if interp:
        os.environ["HY_CUDA_NO_NBODY_REG"] = "1"
    else:
        os.environ.pop("HY_CUDA_NO_NBODY_REG", None)
    try:
        ta = hy.taylor_adaptive_batch(sys_, ic, **kw)
    finally:
        if old is None:
            os.environ.pop("HY_CUDA_NO_NBODY_REG", None)
        else:
            os.environ["HY_CUDA_NO_NBODY_REG"] = old
    return ta


def _rel(a, b):
    return np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))


def test_variant_selected():
    ta = _make(common.oss_sys(), common.oss_ensemble(32))
    assert ta._ctx.launch_info()["kernel_variant"] == 6
    tb = _make(common.oss_sys(), common.oss_ensemble(32), interp=True)
    assert tb._ctx.launch_info()["kernel_variant"] == 0
    # other systems / events / orders above 20 keep the interpreter
    tc = hy.taylor_adaptive_batch(common.pendulum_sys(), common.PEND_IC)
    assert tc._ctx.launch_info()["kernel_variant"] == 0
    # tol = 1e-18 (order 22, the reference's benchmark configuration): the 6-body FP64 order-22 build
    td = _make(common.oss_sys(), common.oss_ensemble(8), tol=1e-18)
    assert td.order == 22 and td._ctx.launch_info()["kernel_variant"] == 226
    te = _make(hy.model.nbody(5, masses=list(common.OSS_MASSES[:5]), Gconst=common.OSS_G),
               common.oss_ensemble(8)[:30].copy(), tol=1e-18)
    # (not served by a register kernel: the interpreter, or a run-time compiled kernel)
    assert te.order == 22 and te._ctx.launch_info()["kernel_variant"] in (0, 1000)


@pytest.mark.parametrize("fp", [np.float64, np.float32])
def test_bitwise_vs_interpreter(fp):
    B = 333  # ragged: not a multiple of the trajectories per CTA
    sys_ = common.oss_sys()
    ic = common.oss_ensemble(B, amp=1e-3).astype(fp)
    a = _make(sys_, ic, fp_type=fp)
    b = _make(sys_, ic, interp=True, fp_type=fp)
    # single steps with tc
    a.step(write_tc=True)
    b.step(write_tc=True)
    assert np.array_equal(a.tc, b.tc)
    assert np.array_equal(a.state, b.state)
    assert [r[1] for r in a.step_res] == [r[1] for r in b.step_res]
    # propagate, per-lane final times, forward then backward
    tf = np.linspace(40.0, 60.0, B).astype(fp)
    a.propagate_until(tf)
    b.propagate_until(tf)
    assert np.array_equal(a.state, b.state)
    assert a.propagate_res == b.propagate_res
    assert np.array_equal(a.time, tf)
    a.propagate_for(fp(-7.5))
    b.propagate_for(fp(-7.5))
    assert np.array_equal(a.state, b.state)
    assert a.propagate_res == b.propagate_res


@pytest.mark.parametrize("nb", [3, 4, 5])
def test_smaller_systems_bitwise_and_oracle(nb):
    # the first nb bodies of the outer Solar System (Sun + giant planets)
    B = 37
    sys_ = hy.model.nbody(nb, masses=list(common.OSS_MASSES[:nb]), Gconst=common.OSS_G)
    ic = common.oss_ensemble(B, amp=1e-4)[: 6 * nb].copy()
    a = _make(sys_, ic)
    b = _make(sys_, ic, interp=True)
    assert a._ctx.launch_info()["kernel_variant"] == nb
    assert b._ctx.launch_info()["kernel_variant"] == 0
    a.step(write_tc=True)
    b.step(write_tc=True)
    assert np.array_equal(a.tc, b.tc)
    assert np.array_equal(a.state, b.state)
    a.propagate_until(80.0)
    b.propagate_until(80.0)
    assert np.array_equal(a.state, b.state)
    assert a.propagate_res == b.propagate_res
    orc = COracle(D.decompose(sys_, a.order), ic)
    orc.step()
    oc, mn, mx, ns, _ = orc.propagate_until(80.0)
    assert [r[3] for r in a.propagate_res] == list(ns)
    assert _rel(a.state, orc.state) < 1e-11


@pytest.mark.parametrize("B", [1, 2, 3, 17])
def test_tiny_batches(B):
    # fewer trajectories than one warp / one CTA holds: the spare half-warps idle through the steps
    sys_ = common.oss_sys()
    ic = common.oss_ensemble(B, amp=1e-5)
    a = _make(sys_, ic)
    b = _make(sys_, ic, interp=True)
    assert a._ctx.launch_info()["kernel_variant"] == 6
    a.propagate_until(30.0)
    b.propagate_until(30.0)
    assert np.array_equal(a.state, b.state)
    assert a.propagate_res == b.propagate_res
    assert np.all(a.time == 30.0)


def test_oracle_parity_fp64():
    B = 48
    sys_ = common.oss_sys()
    ic = common.oss_ensemble(B)
    ta = _make(sys_, ic)
    assert ta._ctx.launch_info()["kernel_variant"] == 6
    orc = COracle(D.decompose(sys_, ta.order), ic)
    # accepted step sequence, step by step
    worst = 0.0
    for _ in range(60):
        ta.step()
        oc, h = orc.step()
        hg = np.array([r[1] for r in ta.step_res])
        worst = max(worst, float(np.max(np.abs(hg - h) / np.abs(h))))
        assert [int(r[0]) for r in ta.step_res] == list(oc)
    assert worst < 1e-12, worst
    assert _rel(ta.state, orc.state) < 1e-12
    # then a propagate_until: step counts, min/max h, final state
    ta.propagate_until(150.0)
    oc, mn, mx, ns, _ = orc.propagate_until(150.0)
    res = ta.propagate_res
    assert [r[3] for r in res] == list(ns)
    assert _rel(np.array([r[1] for r in res]), mn) < 1e-11
    assert _rel(np.array([r[2] for r in res]), mx) < 1e-11
    assert _rel(ta.state, orc.state) < 1e-11


def test_low_order_and_features():
    # tol = 1e-9 -> order 12 < NBR_PMAX: the unrolled order loop exits early
    B = 40
    sys_ = common.oss_sys()
    ic = common.oss_ensemble(B, amp=1e-6)
    a = _make(sys_, ic, tol=1e-9)
    b = _make(sys_, ic, interp=True, tol=1e-9)
    assert a.order == 12 and a._ctx.launch_info()["kernel_variant"] == 6
    grid = np.repeat(np.linspace(0.0, 30.0, 7)[:, None], B, axis=1)
    ra = a.propagate_grid(grid)
    rb = b.propagate_grid(grid)
    assert np.array_equal(ra[1], rb[1])
    assert np.array_equal(a.state, b.state)
    # continuous output + high accuracy + max_delta_t + max_steps
    a2 = _make(sys_, ic, high_accuracy=True)
    b2 = _make(sys_, ic, interp=True, high_accuracy=True)
    ca = a2.propagate_for(20.0, max_delta_t=0.5, c_output=True)[0]
    cb = b2.propagate_for(20.0, max_delta_t=0.5, c_output=True)[0]
    tq = np.repeat(np.array([[3.3], [11.7], [19.9]]), B, axis=1)
    assert np.array_equal(ca(tq), cb(tq))
    assert np.array_equal(a2.state, b2.state)
    a2.propagate_for(50.0, max_steps=9)
    b2.propagate_for(50.0, max_steps=9)
    assert a2.propagate_res == b2.propagate_res
    assert all(int(r[0]) == int(hy.taylor_outcome.step_limit) for r in a2.propagate_res)
    assert np.array_equal(a2.state, b2.state)


def test_energy_conservation_1000_steps():
    B = 64
    ic = common.oss_ensemble(B)
    ta = _make(common.oss_sys(), ic)
    e0 = common.oss_energy(ic)
    ta.propagate_for(700.0)
    assert min(r[3] for r in ta.propagate_res) >= 900
    e1 = common.oss_energy(ta.state)
    assert np.max(np.abs((e1 - e0) / e0)) < 5e-14


def test_warpgroup_rotation_variant_bitwise():
    # experimental HY_CUDA_WGX=1 kernel (setmaxnreg register hand-over between warpgroups):
    # slower than the default, kept for measurement - it must still agree bit for bit
    B = 24 * 148 + 77
    sys_ = common.oss_sys()
    ic = common.oss_ensemble(B, amp=1e-6)
    a = _make(sys_, ic)
    os.environ["HY_CUDA_WGX"] = "1"
    try:
        b = _make(sys_, ic)
    finally:
        os.environ.pop("HY_CUDA_WGX", None)
    assert a._ctx.launch_info()["kernel_variant"] == 6
    if b._ctx.launch_info()["kernel_variant"] != 106:
        pytest.skip("warpgroup-rotation geometry not available on this device")
    a.propagate_until(20.0)
    b.propagate_until(20.0)
    assert np.array_equal(a.state, b.state)
    assert a.propagate_res == b.propagate_res


@pytest.mark.parametrize("high_accuracy", [False, True])
def test_order22_build_bitwise_and_oracle(high_accuracy):
    # tol = 1e-18 -> order 22 (ensemble_batch_perf.ipynb:233 runs the outer Solar System with
    # high_accuracy=True, tol=1e-18): propagate_kernel<double,16,true,6,false,22>
    B = 77
    sys_ = common.oss_sys()
    ic = common.oss_ensemble(B, amp=1e-4)
    a = _make(sys_, ic, tol=1e-18, high_accuracy=high_accuracy)
    b = _make(sys_, ic, interp=True, tol=1e-18, high_accuracy=high_accuracy)
    assert a.order == 22 and a._ctx.launch_info()["kernel_variant"] == 226
    assert b._ctx.launch_info()["kernel_variant"] == 0
    a.step(write_tc=True)
    b.step(write_tc=True)
    assert np.array_equal(a.tc, b.tc)
    assert np.array_equal(a.state, b.state)
    tf = np.linspace(40.0, 60.0, B)
    a.propagate_until(tf)
    b.propagate_until(tf)
    assert np.array_equal(a.state, b.state)
    assert a.propagate_res == b.propagate_res
    a.propagate_for(-7.5)
    b.propagate_for(-7.5)
    assert np.array_equal(a.state, b.state)
    if not high_accuracy:
        orc = COracle(D.decompose(sys_, 22), ic, tol=1e-18)
        orc.step()
        orc.propagate_until(tf)
        orc.propagate_for(-7.5)
        assert _rel(a.state, orc.state) < 1e-11


def test_non_finite_lanes_stop_alone():
    # a NaN coordinate and two bodies on top of each other (r = 0): those trajectories end with
    # err_nf_state; the other trajectory of the same warp is unaffected (the finiteness ballot is per half-warp)
    B = 9
    sys_ = common.oss_sys()
    ic = common.oss_ensemble(B, amp=1e-6)
    ic[7, 2] = np.nan
    ic[6:9, 5] = ic[0:3, 5]  # body 1 placed on body 0
    a = _make(sys_, ic)
    b = _make(sys_, ic, interp=True)
    assert a._ctx.launch_info()["kernel_variant"] == 6
    a.propagate_until(30.0)
    b.propagate_until(30.0)
    oa = [int(r[0]) for r in a.propagate_res]
    assert oa == [int(r[0]) for r in b.propagate_res]
    nf = int(hy.taylor_outcome.err_nf_state)
    assert oa[2] == nf and oa[5] == nf
    ok = [i for i in range(B) if i not in (2, 5)]
    assert all(oa[i] == int(hy.taylor_outcome.time_limit) for i in ok)
    assert np.array_equal(a.state[:, ok], b.state[:, ok])
    assert np.all(a.time[ok] == 30.0)
