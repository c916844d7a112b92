import os, sys
sys.path.insert(0, "heyoka.py_b200"); sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import hy_b200 as hy
import common
ic = common.oss_ensemble(40, amp=1e-3)
ta = hy.taylor_adaptive_batch(common.oss_sys(), ic)
print(ta._ctx.launch_info())
ta.step(write_tc=True)
print("step ok", ta.step_res[0])
ta.propagate_until(10.0)
print("prop ok", ta.propagate_res[0])
