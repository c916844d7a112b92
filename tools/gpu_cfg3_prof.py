import os, sys
sys.path.insert(0, "heyoka.py_b200"); sys.path.insert(0, ".")
import numpy as np
import hy_b200 as hy
from hy_b200 import workloads as W
B = 200000
ic = W.cr3bp_ensemble(B)
ta = hy.taylor_adaptive_batch(W.cr3bp_sys(0.01), ic)
for rep in range(2):
    ta.state[:] = ic; ta.set_time(0.0)
    c, _ = ta.propagate_until(20.0, c_output=True)
    del c
print("done")
