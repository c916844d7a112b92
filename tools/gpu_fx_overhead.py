"""Cost of the FX build of the register-resident CR3BP kernel (developer tool): the same propagation
through the plain build, the FX build with no feature in use (all-ones active mask), and the FX
build recording the continuous output."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "heyoka.py_b200")); sys.path.insert(0, ROOT)
import numpy as np
import hy_b200 as hy
from hy_b200 import workloads as W

B = int(os.environ.get("QB", 500000))
ic = W.cr3bp_ensemble(B)
ta = hy.taylor_adaptive_batch(W.cr3bp_sys(0.01), ic)
ctx = ta._ctx
oc = np.zeros(B, dtype=np.int64); ns = np.zeros(B, dtype=np.uint64)
tf = np.full(B, 20.0)
act = np.ones(B, dtype=np.uint8)
for name, kw in (("plain", {}), ("FX, active mask only", {"active": act}), ("FX, recording", {"c_output": 1})):
    for rep in range(2):
        ta.state[:] = ic; ta.set_time(0.0); ta._push()
        ctx.propagate_ex(oc, None, None, ns, t=tf, **kw)
        ms, nl = ctx.last_timing()
        if kw.get("c_output"):
            r = ctx.cout_detach(); r.close()
    print("%-24s %.1f ms  %.3e steps/s  (%d launches)" % (name, ms, ns.sum() / ms * 1e3, nl), flush=True)
