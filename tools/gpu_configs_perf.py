"""Throughput of the tape interpreter on the other BASELINE configs (developer tool):
config 3 (CR3BP, FP64/FP32), config 4 (Kepler+J2 with first-order variational equations)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "heyoka.py_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import hy_b200 as hy
from hy_b200 import _cabi
import common

B = int(os.environ.get("QB", 200000))


def run(name, sys_, ic, t_end, fp=np.float64, **kw):
    ta = hy.taylor_adaptive_batch(sys_, ic.astype(fp), fp_type=fp, **kw)
    fl, lo = ta._dc.flops_per_step()
    best = None
    for rep in range(2):
        ta.state[:] = ic.astype(fp) if ta.state.shape == ic.shape else ta.state
        ta.set_time(fp(0.0))
        ta.propagate_until(fp(t_end))
        ms, _ = ta._ctx.last_timing()
        ns = int(ta.propagate_res_arrays[3].sum())
        best = (ns / (ms * 1e-3), ns * fl / (ms * 1e-3) / 1e12)
    li = ta._ctx.launch_info()
    print("%-28s B=%d order=%d G=%d T=%d smem=%d: %.3g steps/s, %.2f TFLOP/s (flops/step %d)" % (
        name, ic.shape[1], ta.order, li["group"], li["traj_per_cta"], li["smem_bytes"], best[0], best[1], fl), flush=True)


run("cfg3 CR3BP f64", common.cr3bp_sys(), common.cr3bp_ensemble(B), 20.0)
run("cfg3 CR3BP f32", common.cr3bp_sys(), common.cr3bp_ensemble(B), 20.0, fp=np.float32)
vs = hy.var_ode_sys(common.kepler_j2_sys(), hy.var_args.vars, order=1)
ic4 = common.kepler_j2_ensemble(B // 4)
ta = hy.taylor_adaptive_batch(vs, ic4)
fl, lo = ta._dc.flops_per_step()
st0 = ta.state.copy()
for rep in range(2):  # (the first call pays the lazy load of the kernel)
    ta.state[:] = st0
    ta.set_time(0.0)
    ta.propagate_until(3000.0)
    ms, _ = ta._ctx.last_timing()
    ns = int(ta.propagate_res_arrays[3].sum())
    print("  cfg4 rep", rep, "steps", ns, "ms %.2f" % ms, flush=True)
li = ta._ctx.launch_info()
print("cfg4 Kepler+J2 variational   B=%d order=%d G=%d T=%d: %.3g steps/s, %.2f TFLOP/s (flops/step %d)" % (
    ic4.shape[1], ta.order, li["group"], li["traj_per_cta"], ns / (ms * 1e-3), ns * fl / (ms * 1e-3) / 1e12, fl))
