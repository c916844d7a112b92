"""Diagnostic: GPU vs C oracle step sequences, free-running and re-synchronised every step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "heyoka.py_b200")); sys.path.insert(0, ROOT)
import numpy as np
import hy_b200 as hy
from hy_b200 import decompose as D, workloads as W
from oracle.c_oracle import COracle

def run(name, sys_, ic, fp, nsteps, sync):
    ta = hy.taylor_adaptive_batch(sys_, ic.astype(fp), fp_type=fp)
    orc = COracle(D.decompose(ta._sys, ta.order), ta.state.copy(), fp_type=fp)
    worst = 0.0; trace = []
    for i in range(nsteps):
        if sync:
            ta.state[:] = orc.state
            ta.set_dtime(orc.t_hi.copy(), orc.t_lo.copy())
        ta.step(); oc, h = orc.step()
        hg = np.array([r[1] for r in ta.step_res], dtype=np.float64)
        e = float(np.max(np.abs(hg - h) / np.abs(h)))
        worst = max(worst, e)
        if (i + 1) % max(1, nsteps // 8) == 0:
            trace.append("%d:%.1e" % (i + 1, worst))
    ds = np.max(np.abs(ta.state.astype(float) - orc.state.astype(float)) / np.maximum(1, np.abs(orc.state.astype(float))))
    print(name, fp.__name__, "sync" if sync else "free", "worst dh/h", "%.2e" % worst, "state", "%.2e" % ds, " ".join(trace), flush=True)

vs = hy.var_ode_sys(W.kepler_j2_sys(), hy.var_args.vars)
for sync in (True, False):
    run("cfg4", vs, W.kepler_j2_ensemble(8, seed=7), np.float64, 200, sync)
    run("cfg2", W.oss_sys(), W.oss_ensemble(8), np.float64, 1000, sync)
    run("cfg2", W.oss_sys(), W.oss_ensemble(8), np.float32, 1000, sync)
    run("cfg3", W.cr3bp_sys(0.01), W.cr3bp_ensemble(8), np.float64, 1000, sync)
    run("cfg1", W.pendulum_sys(), W.PEND_IC, np.float64, 1000, sync)
