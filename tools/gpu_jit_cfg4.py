"""Run-time compiled kernel (csrc/hy_jit.hpp) on config 4 (Kepler+J2 with first-order variational
equations): bit-for-bit comparison with the tape interpreter, then throughput (developer tool)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "heyoka.py_b200")); sys.path.insert(0, ROOT)
import numpy as np
import hy_b200 as hy
from hy_b200 import workloads as W

B = int(os.environ.get("QB", 125000))
T_END = float(os.environ.get("QT", 3000.0))
vs = hy.var_ode_sys(W.kepler_j2_sys(), hy.var_args.vars, order=1)

if os.environ.get("QCHECK", "1") == "1":
    ic = W.kepler_j2_ensemble(64)
    t0 = time.time()
    tj = hy.taylor_adaptive_batch(vs, ic)
    print("jit ctor %.1f s" % (time.time() - t0), tj._ctx.launch_info(), flush=True)
    ti = hy.taylor_adaptive_batch(vs, ic, compact_mode=True)
    print("interp", ti._ctx.launch_info(), flush=True)
    for ta in (tj, ti):
        ta.propagate_until(2000.0)
    print("steps equal", np.array_equal(tj.propagate_res_arrays[3], ti.propagate_res_arrays[3]),
          "state bitwise", np.array_equal(tj.state, ti.state),
          "max rel diff", float(np.max(np.abs(tj.state - ti.state) / np.maximum(1e-300, np.abs(ti.state)))), flush=True)

for mode in os.environ.get("QMODES", "jit,interp").split(","):
    ic = W.kepler_j2_ensemble(B)
    ta = hy.taylor_adaptive_batch(vs, ic, compact_mode=(mode == "interp"))
    fl, lo = ta._dc.flops_per_step()
    st0 = ta.state.copy()
    for rep in range(2):
        ta.state[:] = st0
        ta.set_time(0.0)
        ta.propagate_until(T_END)
        ms, _ = ta._ctx.last_timing()
        ns = int(ta.propagate_res_arrays[3].sum())
    li = ta._ctx.launch_info()
    print("%s: variant %d T %d threads %d ctas %d smem %d ws_in_smem %d regs %d: %.3e steps/s, %.2f TFLOP/s (%.1f ms, %d steps)" % (
        mode, li["kernel_variant"], li["traj_per_cta"], li["threads"], li["ctas"], li["smem_bytes"], li["ws_in_smem"],
        li["regs_per_thread"], ns / (ms * 1e-3), ns * fl / (ms * 1e-3) / 1e12, ms, ns), flush=True)
