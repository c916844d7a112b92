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
