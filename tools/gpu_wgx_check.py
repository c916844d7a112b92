"""Warpgroup-rotation variant (HY_CUDA_WGX=1) against the default register-resident kernel:
bitwise agreement and throughput (developer tool)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "heyoka.py_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import hy_b200 as hy
import common

B = int(os.environ.get("QB", 125000))
T_END = float(os.environ.get("QT", 100.0))
ic = common.oss_ensemble(B)
res = {}
for wgx in (0, 1):
    os.environ["HY_CUDA_WGX"] = str(wgx)
    ta = hy.taylor_adaptive_batch(common.oss_sys(), ic)
    print("wgx", wgx, ta._ctx.launch_info(), flush=True)
    for rep in range(2):
        ta.state[:] = ic
        ta.set_time(0.0)
        ta.propagate_until(T_END)
        ms, _ = ta._ctx.last_timing()
        ns = int(ta.propagate_res_arrays[3].sum())
        print("  rep", rep, "steps", ns, "ms", ms, "steps/s %.4g" % (ns / (ms * 1e-3)), flush=True)
    res[wgx] = (ta.state.copy(), ta.propagate_res_arrays[3].copy())
print("bitwise equal:", np.array_equal(res[0][0], res[1][0]), np.array_equal(res[0][1], res[1][1]))
