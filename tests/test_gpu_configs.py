"""GPU: BASELINE.json's configurations 2-5 at sizes beyond what the oracle can
follow, checked through size-independent properties (conserved quantities,
lane independence, symplecticity, event-surface residuals) plus oracle spot
checks on a few lanes."""

import numpy as np
import pytest

import hy_b200 as hy
from hy_b200 import decompose as D
from oracle.c_oracle import COracle

import common
from hy_b200 import workloads as W

pytestmark = pytest.mark.gpu


def test_config2_outer_solar_system_20k():
    B = 20000
    sys_ = W.oss_sys()
    ic = W.oss_ensemble(B)
    ta = hy.taylor_adaptive_batch(sys_, ic)
    ta.propagate_until(200.0)
    oc, mn, mx, ns = ta.propagate_res_arrays
    assert np.all(oc == int(hy.taylor_outcome.time_limit)) and np.all(ta.time == 200.0)
    assert ns.min() > 250 and ns.max() < 320
    e0, e1 = W.oss_energy(ic), W.oss_energy(ta.state)
    assert np.max(np.abs((e1 - e0) / e0)) < 2e-14      # conserved-energy drift
    # oracle spot check on 8 lanes spread over the batch (step counts + state)
    idx = np.linspace(0, B - 1, 8).astype(int)
    orc = COracle(D.decompose(sys_, ta.order), ic[:, idx])
    r = orc.propagate_until(200.0)
    assert list(ns[idx]) == list(r[3])
    assert np.max(np.abs(ta.state[:, idx] - orc.state) / np.maximum(1, np.abs(orc.state))) < 1e-12 * 20


@pytest.mark.parametrize("fp,tol", [(np.float64, 1e-13), (np.float32, 2e-5)])
def test_config3_cr3bp_c_output(fp, tol):
    B = 4096
    sys_ = W.cr3bp_sys(0.01)
    ic = W.cr3bp_ensemble(B).astype(fp)
    ta = hy.taylor_adaptive_batch(sys_, ic, fp_type=fp)
    c_out, _ = ta.propagate_until(fp(20.0), c_output=True)
    assert c_out.n_steps >= 20 and c_out.batch_size == B
    tq = np.repeat(np.linspace(0.0, 20.0, 16), B).reshape(16, B).astype(fp)
    out = c_out(tq)
    assert out.shape == (16, 6, B)
    eps = np.finfo(fp).eps
    assert np.max(np.abs(out[0] - ic)) <= 8 * eps
    assert np.max(np.abs(out[-1] - ta.state) / np.maximum(1, np.abs(ta.state))) < 100 * eps
    j0 = W.cr3bp_jacobi(ic.astype(np.float64))
    for q in range(16):
        jq = W.cr3bp_jacobi(out[q].astype(np.float64))
        assert np.max(np.abs((jq - j0) / j0)) < tol
    # ragged per-lane step counts are reported consistently
    ns = ta.propagate_res_arrays[3]
    assert c_out.n_steps == ns.max()


def test_config4_kepler_j2_variational():
    B = 2048
    sys_ = W.kepler_j2_sys()
    vs = hy.var_ode_sys(sys_, hy.var_args.vars)
    ic = W.kepler_j2_ensemble(B)
    ta = hy.taylor_adaptive_batch(vs, ic)
    assert ta.dim == 42 and np.all(ta.state[6:].reshape(6, 6, B)[:, :, 0] == np.eye(6))
    T = 5000.0   # s, a bit less than one orbital period
    ta.propagate_until(T)
    assert np.all(ta.propagate_res_arrays[0] == int(hy.taylor_outcome.time_limit))
    e0, e1 = W.kepler_j2_energy(ic), W.kepler_j2_energy(ta.state[:6])
    assert np.max(np.abs((e1 - e0) / e0)) < 1e-13
    Phi = ta.state[6:].reshape(6, 6, B)
    # symplecticity: det Phi = 1
    det = np.linalg.det(np.moveaxis(Phi, 2, 0))
    assert np.max(np.abs(det - 1)) < 1e-9
    # finite differences of the plain flow on 16 lanes
    idx = np.arange(0, B, B // 16)[:16]
    sub = ic[:, idx]
    for j in range(6):
        h = 1e-3 if j < 3 else 1e-6
        d = np.zeros_like(sub)
        d[j] = h
        tp = hy.taylor_adaptive_batch(sys_, sub + d)
        tm = hy.taylor_adaptive_batch(sys_, sub - d)
        tp.propagate_until(T)
        tm.propagate_until(T)
        fd = (tp.state - tm.state) / (2 * h)
        got = Phi[:, j, :][:, idx]
        assert np.max(np.abs(fd - got) / (1 + np.abs(got))) < 1e-5


def test_config5_cr3bp_terminal_events_4k():
    mu, B = 0.01, 4096
    sys_ = W.cr3bp_sys(mu)
    x, y, z = hy.make_vars("x", "y", "z")
    R1 = R2 = 0.012
    Resc = 5.0
    evs = [(x - mu) ** 2 + y * y + z * z - R1 ** 2,
           (x - mu + 1.0) ** 2 + y * y + z * z - R2 ** 2,
           x * x + y * y + z * z - Resc ** 2]
    rng = np.random.default_rng(20251022)
    ic = np.array([-0.80, 0.0, 0.0, 0.0, -0.6276410653920693, 0.0])[:, None] * np.ones((1, B))
    ic[0] += rng.uniform(-1e-2, 1e-2, B)
    ic[4] += rng.uniform(-1e-2, 1e-2, B)
    ta = hy.taylor_adaptive_batch(sys_, ic, t_events=[hy.t_event_batch(e) for e in evs])
    ta.propagate_until(100.0)
    oc = ta.propagate_res_arrays[0]
    done = oc == int(hy.taylor_outcome.time_limit)
    hit = ~done
    assert np.all((oc[hit] >= -3) & (oc[hit] <= -1))
    assert np.all(ta.time[done] == 100.0) and np.all(ta.time[hit] < 100.0)
    # stopped lanes sit on the surface of the event that fired
    st = ta.state
    g = [(st[0] - mu) ** 2 + st[1] ** 2 + st[2] ** 2 - R1 ** 2,
         (st[0] - mu + 1) ** 2 + st[1] ** 2 + st[2] ** 2 - R2 ** 2,
         st[0] ** 2 + st[1] ** 2 + st[2] ** 2 - Resc ** 2]
    for e in range(3):
        m = oc == -e - 1
        if np.any(m):
            assert np.max(np.abs(g[e][m])) < 1e-12
    # the Jacobi constant is conserved up to the event / final time
    j0, j1 = W.cr3bp_jacobi(ic), W.cr3bp_jacobi(st)
    assert np.max(np.abs((j1 - j0) / j0)) < 1e-9
    assert hit.sum() > 0 and done.sum() > 0
