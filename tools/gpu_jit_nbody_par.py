"""6-body outer Solar System with the masses as runtime parameters (not matched by the register
kernel): run-time compiled kernel vs tape interpreter (developer tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "heyoka.py_b200")); sys.path.insert(0, ROOT)
import numpy as np
import hy_b200 as hy
from hy_b200 import workloads as W, model

B = int(os.environ.get("QB", 125000))
sys_ = model.nbody(6, masses=[hy.par[i] for i in range(6)], Gconst=W.OSS_G)
ic = W.oss_ensemble(B)
pars = W.OSS_MASSES[:, None] * np.ones((1, B))
for mode in ("jit", "interp"):
    ta = hy.taylor_adaptive_batch(sys_, ic, pars=pars, compact_mode=(mode == "interp"))
    fl, lo = ta._dc.flops_per_step()
    for rep in range(2):
        ta.state[:] = ic
        ta.set_time(0.0)
        ta.propagate_until(float(os.environ.get("QT", 300.0)))
        ms, _ = ta._ctx.last_timing()
        ns = int(ta.propagate_res_arrays[3].sum())
    li = ta._ctx.launch_info()
    print("%s: variant %d G %d T %d threads %d smem %d regs %d: %.3e steps/s, %.2f TFLOP/s (%.1f ms)" % (
        mode, li["kernel_variant"], li["group"], li["traj_per_cta"], li["threads"], li["smem_bytes"], li["regs_per_thread"],
        ns / (ms * 1e-3), ns * fl / (ms * 1e-3) / 1e12, ms), flush=True)
