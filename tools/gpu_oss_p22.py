"""The reference's own benchmark configuration (ensemble_batch_perf.ipynb: outer Solar System, high_accuracy=True,
tol=1e-18 -> order 22): order-22 register build vs the tape interpreter (developer tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "heyoka.py_b200")); sys.path.insert(0, ROOT)
import numpy as np
import hy_b200 as hy
from hy_b200 import workloads as W

B = int(os.environ.get("QB", 125000))
T_END = float(os.environ.get("QT", 100.0))
ic = W.oss_ensemble(B)
for interp in (1, 0):
    os.environ["HY_CUDA_NO_NBODY_REG"] = str(interp)
    ta = hy.taylor_adaptive_batch(W.oss_sys(), ic, tol=1e-18, high_accuracy=True)
    fl = ta._dc.flops_per_step()[0]
    for rep in range(2):
        ta.state[:] = ic
        ta.set_time(0.0)
        ta.propagate_until(T_END)
        ms, _ = ta._ctx.last_timing()
        ns = int(ta.propagate_res_arrays[3].sum())
    li = ta._ctx.launch_info()
    print("interpreter" if interp else "register   ", "order", ta.order, "variant", li["kernel_variant"], "regs", li["regs_per_thread"],
          "steps", ns, "ms %.2f" % ms, "steps/s %.4g" % (ns / (ms * 1e-3)), "TFLOP/s %.2f" % (ns * fl / (ms * 1e-3) / 1e12), flush=True)
