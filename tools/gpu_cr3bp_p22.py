import os, sys
sys.path.insert(0, "heyoka.py_b200"); sys.path.insert(0, ".")
import numpy as np
import hy_b200 as hy
from hy_b200 import workloads as W
B = 1000000
ic = W.cr3bp_ensemble(B)
for interp in (1, 0):
    os.environ["HY_CUDA_NO_CR3BP_REG"] = str(interp)
    ta = hy.taylor_adaptive_batch(W.cr3bp_sys(0.01), ic, tol=1e-18)
    for rep in range(2):
        ta.state[:] = ic; ta.set_time(0.0); ta.propagate_until(20.0)
        ms, _ = ta._ctx.last_timing(); ns = int(ta.propagate_res_arrays[3].sum())
    print("interp" if interp else "register", ta.order, ta._ctx.launch_info()["kernel_variant"], "steps/s %.4g" % (ns / (ms * 1e-3)), flush=True)
