"""GPU tests of the register-resident CR3BP kernel (csrc/hy_cr3bp_reg.cuh).

hy_create recognises the tape of model.cr3bp (order 20 in FP64, 9 in FP32) and runs
the jets in registers, two lanes per trajectory; HY_CUDA_NO_CR3BP_REG=1 forces the
tape interpreter on the same tape.  Both implement the same recurrences in the same
term order, so they must agree BIT FOR BIT; the oracle comparison (1e-12 relative in
FP64, 1e-5 in FP32 on a short horizon: the system is chaotic) pins both.
"""

import os

import numpy as np
import pytest

import hy_b200 as hy
from hy_b200 import decompose as D
from hy_b200 import workloads as W
from oracle.c_oracle import COracle

pytestmark = pytest.mark.gpu

CRB = 203  # hy_launch_info.kernel_variant of the register-resident CR3BP kernel


def _make(sys_, ic, interp=False, **kw):
    old = os.environ.get("HY_CUDA_NO_CR3BP_REG")
    if interp:
        os.environ["HY_CUDA_NO_CR3BP_REG"] = "1"
    else:
        os.environ.pop("HY_CUDA_NO_CR3BP_REG", None)
    try:
        ta = hy.taylor_adaptive_batch(sys_, ic, **kw)
    finally:
        if old is None:
            os.environ.pop("HY_CUDA_NO_CR3BP_REG", None)
        else:
            os.environ["HY_CUDA_NO_CR3BP_REG"] = old
    return ta


def _rel(a, b):
    return np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))


def test_variant_selected():
    sys_ = W.cr3bp_sys(0.01)
    ic = W.cr3bp_ensemble(64)
    assert _make(sys_, ic)._ctx.launch_info()["kernel_variant"] == CRB
    assert _make(sys_, ic.astype(np.float32), fp_type=np.float32)._ctx.launch_info()["kernel_variant"] == CRB
    assert _make(sys_, ic, interp=True)._ctx.launch_info()["kernel_variant"] == 0
    # lower orders take the order-checked path of the same kernel
    assert _make(sys_, ic, tol=1e-9)._ctx.launch_info()["kernel_variant"] == CRB
    # tol = 1e-18 (order 22, the setting of the reference's CR3BP notebook): the FP64 order-22 build
    t22 = _make(sys_, ic, tol=1e-18)
    assert t22.order == 22 and t22._ctx.launch_info()["kernel_variant"] == 222
    # orders above that are not served by the register kernel (interpreter, or a run-time compiled kernel)
    assert _make(sys_, ic, tol=1e-21)._ctx.launch_info()["kernel_variant"] in (0, 1000)
    x = hy.make_vars("x")
    ev = hy.t_event_batch(x - 5.0)
    # (round 2: an event-carrying system keeps the register kernel - the events run from the event tape)
    assert hy.taylor_adaptive_batch(sys_, ic, t_events=[ev])._ctx.launch_info()["kernel_variant"] == CRB


@pytest.mark.parametrize("fp", [np.float64, np.float32])
@pytest.mark.parametrize("mu", [0.01, 0.3])
def test_bitwise_vs_interpreter(fp, mu):
    B = 333  # ragged: not a multiple of the trajectories per warp / CTA
    sys_ = W.cr3bp_sys(mu)
    ic = W.cr3bp_ensemble(B).astype(fp)
    a = _make(sys_, ic, fp_type=fp)
    b = _make(sys_, ic, interp=True, fp_type=fp)
    assert a._ctx.launch_info()["kernel_variant"] == CRB and b._ctx.launch_info()["kernel_variant"] == 0
    # single steps with tc
    a.step(write_tc=True)
    b.step(write_tc=True)
    assert np.array_equal(a.tc, b.tc)
    assert np.array_equal(a.state, b.state)
    assert [r[1] for r in a.step_res] == [r[1] for r in b.step_res]
    # propagate, per-lane final times (ragged step counts), forward then backward
    tf = np.linspace(3.0, 9.0, B).astype(fp)
    a.propagate_until(tf)
    b.propagate_until(tf)
    assert np.array_equal(a.state, b.state)
    assert a.propagate_res == b.propagate_res
    assert np.array_equal(a.time, tf)
    a.propagate_for(fp(-2.5))
    b.propagate_for(fp(-2.5))
    assert np.array_equal(a.state, b.state)
    assert a.propagate_res == b.propagate_res


@pytest.mark.parametrize("fp,tol,order", [(np.float64, 1e-9, 12), (np.float64, 1e-4, 6), (np.float32, 1e-4, 6),
                                          (np.float64, 1e-18, 22), (np.float64, 1e-17, 21)])
def test_lower_orders_bitwise(fp, tol, order):
    # orders below the unrolled maximum: run-time order checks in the register kernel
    B = 75
    sys_ = W.cr3bp_sys(0.01)
    ic = W.cr3bp_ensemble(B).astype(fp)
    a = _make(sys_, ic, fp_type=fp, tol=fp(tol))
    b = _make(sys_, ic, interp=True, fp_type=fp, tol=fp(tol))
    assert a.order == order and b.order == order
    assert a._ctx.launch_info()["kernel_variant"] == (222 if order > 20 else CRB)
    assert b._ctx.launch_info()["kernel_variant"] == 0
    a.step(write_tc=True)
    b.step(write_tc=True)
    assert np.array_equal(a.tc, b.tc)
    a.propagate_until(fp(6.0))
    b.propagate_until(fp(6.0))
    assert np.array_equal(a.state, b.state)
    assert a.propagate_res == b.propagate_res


@pytest.mark.parametrize("B", [1, 2, 3, 17, 33])
def test_tiny_batches(B):
    # fewer trajectories than one warp holds: the spare lane pairs idle through the steps
    sys_ = W.cr3bp_sys(0.01)
    ic = W.cr3bp_ensemble(B)
    a = _make(sys_, ic)
    b = _make(sys_, ic, interp=True)
    assert a._ctx.launch_info()["kernel_variant"] == CRB
    a.propagate_until(5.0)
    b.propagate_until(5.0)
    assert np.array_equal(a.state, b.state)
    assert a.propagate_res == b.propagate_res
    assert np.all(a.time == 5.0)


def test_features_bitwise():
    # continuous output, grid, max_delta_t, max_steps, high accuracy: the tail of the step is shared code
    B = 100
    sys_ = W.cr3bp_sys(0.01)
    ic = W.cr3bp_ensemble(B)
    a = _make(sys_, ic)
    b = _make(sys_, ic, interp=True)
    ca, _ = a.propagate_until(4.0, c_output=True)
    cb, _ = b.propagate_until(4.0, c_output=True)
    tq = np.repeat(np.linspace(0.0, 4.0, 9), B).reshape(9, B)
    assert np.array_equal(ca(tq), cb(tq))
    assert np.array_equal(a.state, b.state)
    grid = np.repeat(np.linspace(4.0, 6.0, 11), B).reshape(11, B)
    ga = a.propagate_grid(grid)[1]
    gb = b.propagate_grid(grid)[1]
    assert np.array_equal(ga, gb)
    a.propagate_for(1.0, max_delta_t=0.01, max_steps=50)
    b.propagate_for(1.0, max_delta_t=0.01, max_steps=50)
    assert np.array_equal(a.state, b.state)
    assert a.propagate_res == b.propagate_res
    assert all(int(r[0]) == int(hy.taylor_outcome.step_limit) for r in a.propagate_res)
    ha = _make(sys_, ic, high_accuracy=True)
    hb = _make(sys_, ic, interp=True, high_accuracy=True)
    ha.propagate_until(3.0)
    hb.propagate_until(3.0)
    assert np.array_equal(ha.state, hb.state)


def test_oracle_parity_fp64():
    B = 48
    sys_ = W.cr3bp_sys(0.01)
    ic = W.cr3bp_ensemble(B)
    ta = _make(sys_, ic)
    assert ta._ctx.launch_info()["kernel_variant"] == CRB
    orc = COracle(D.decompose(sys_, ta.order), ic)
    worst = 0.0
    for _ in range(40):
        ta.step()
        oc, h = orc.step()
        hg = np.array([r[1] for r in ta.step_res])
        worst = max(worst, float(np.max(np.abs(hg - h) / np.abs(h))))
        assert [int(r[0]) for r in ta.step_res] == list(oc)
    assert worst < 1e-12, worst
    assert _rel(ta.state, orc.state) < 1e-12


def test_oracle_parity_fp32():
    B = 48
    sys_ = W.cr3bp_sys(0.01)
    ic = W.cr3bp_ensemble(B).astype(np.float32)
    ta = _make(sys_, ic, fp_type=np.float32)
    assert ta._ctx.launch_info()["kernel_variant"] == CRB and ta.order == 9
    orc = COracle(D.decompose(sys_, ta.order), ic, fp_type=np.float32)
    for _ in range(20):
        ta.step()
        oc, h = orc.step()
        hg = np.array([r[1] for r in ta.step_res])
        assert np.max(np.abs(hg - h) / np.abs(h)) < 1e-5
    assert _rel(ta.state.astype(np.float64), orc.state.astype(np.float64)) < 1e-5


def test_jacobi_constant_large_batch():
    # size-independent property at a batch that fills the device several times over
    B = 200000
    sys_ = W.cr3bp_sys(0.01)
    ic = W.cr3bp_ensemble(B)
    ta = _make(sys_, ic)
    ta.propagate_until(20.0)
    assert np.all(ta.time == 20.0)
    j0, j1 = W.cr3bp_jacobi(ic), W.cr3bp_jacobi(ta.state)
    assert np.max(np.abs((j1 - j0) / j0)) < 1e-13


def test_ensemble_copy_and_batch_composition():
    # the ensemble driver deep-copies the integrator per iteration (hy_clone keeps the kernel choice);
    # a trajectory gives bit-identical results alone, inside a big batch, and through the ensemble
    from copy import deepcopy

    sys_ = W.cr3bp_sys(0.01)
    ics = W.cr3bp_ensemble(40).reshape(6, 10, 4).transpose(1, 0, 2).copy()  # 10 iterations x [6, 4]
    ta = hy.taylor_adaptive_batch(sys_, np.ascontiguousarray(ics[0]))
    assert ta._ctx.launch_info()["kernel_variant"] == CRB

    def gen(t, i):
        t.state[:] = ics[i]
        return t

    ret = hy.ensemble_propagate_until_batch(ta, 6.0, 10, gen, algorithm="thread", max_workers=4)
    assert len(ret) == 10
    big = hy.taylor_adaptive_batch(sys_, np.ascontiguousarray(np.concatenate(list(ics), axis=1)))
    big.propagate_until(6.0)
    for i in range(10):
        assert ret[i][0]._ctx.launch_info()["kernel_variant"] == CRB
        ser = gen(deepcopy(ta), i)
        ser.propagate_until(6.0)
        assert np.array_equal(ret[i][0].state, ser.state)
        assert ret[i][0].propagate_res == ser.propagate_res
        assert np.array_equal(big.state[:, 4 * i:4 * i + 4], ser.state)


@pytest.mark.parametrize("fp", [np.float64, np.float32])
def test_non_finite_lanes_stop_alone(fp):
    # a NaN initial condition and a start exactly on a primary (r = 0): those lanes end with
    # err_nf_state, their neighbours in the warp are unaffected (outcomes and states as on the interpreter)
    B = 40
    sys_ = W.cr3bp_sys(0.01)
    ic = W.cr3bp_ensemble(B).astype(fp)
    ic[1, 3] = np.nan
    ic[:, 18] = [0.01, 0.0, 0.0, 0.1, 0.2, 0.0]  # x = mu, y = z = 0: on the first primary
    a = _make(sys_, ic, fp_type=fp)
    b = _make(sys_, ic, interp=True, fp_type=fp)
    assert a._ctx.launch_info()["kernel_variant"] == CRB
    a.propagate_until(fp(3.0))
    b.propagate_until(fp(3.0))
    oa = [int(r[0]) for r in a.propagate_res]
    assert oa == [int(r[0]) for r in b.propagate_res]
    nf = int(hy.taylor_outcome.err_nf_state)
    assert oa[3] == nf and oa[18] == nf
    ok = [i for i in range(B) if i not in (3, 18)]
    assert all(oa[i] == int(hy.taylor_outcome.time_limit) for i in ok)
    assert np.array_equal(a.state[:, ok], b.state[:, ok])
    assert np.all(a.time[ok] == fp(3.0))


@pytest.mark.parametrize("interp", [False, True])
def test_notebook_golden_A13(interp):
    # The reference's own numbers (The restricted three-body problem.ipynb:110-112, SURVEY.md App. B A13;
    # tests/golden/notebook_golden.json): tol = 1e-18 -> order 22, 753 steps to t = 200, min / max step size
    # and final state as printed - reproduced by the order-22 register build (and by the interpreter).
    import json

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "notebook_golden.json")) as f:
        g = json.load(f)["cr3bp"]
    ic = np.repeat(np.array(g["ic"], dtype=float)[:, None], 4, axis=1)
    ta = _make(W.cr3bp_sys(g["mu"]), ic, interp=interp, tol=g["tol"])
    assert ta.order == 22 and ta._ctx.launch_info()["kernel_variant"] == (0 if interp else 222)
    ta.propagate_until(g["t_end"])
    for oc, mn, mx, ns in ta.propagate_res:
        assert int(ns) == g["steps"]
        assert abs(mn - g["min_h"]) / g["min_h"] < 1e-12
        assert abs(mx - g["max_h"]) / g["max_h"] < 1e-12
    assert np.max(np.abs(ta.state - np.array(g["final_state"])[:, None])) < 1e-8
