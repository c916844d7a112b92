"""Small run of the register-resident N-body kernel for compute-sanitizer (memcheck / racecheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "heyoka.py_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import hy_b200 as hy
import common
ic = common.oss_ensemble(37, amp=1e-3)
ta = hy.taylor_adaptive_batch(common.oss_sys(), ic)
print(ta._ctx.launch_info())
ta.step(write_tc=True)
ta.propagate_until(3.0)
print("ok", ta.propagate_res[0])
# the register-resident CR3BP kernel (FP64 and FP32), with continuous output and a grid
from hy_b200 import workloads as W
for fp in (np.float64, np.float32):
    ic3 = W.cr3bp_ensemble(37).astype(fp)
    tb = hy.taylor_adaptive_batch(W.cr3bp_sys(0.01), ic3, fp_type=fp)
    print(tb._ctx.launch_info())
    tb.step(write_tc=True)
    tb.propagate_until(fp(1.5), c_output=True)
    tb.propagate_grid(np.repeat(np.linspace(1.5, 2.0, 5), 37).reshape(5, 37).astype(fp))
    print("ok", tb.propagate_res[0])
