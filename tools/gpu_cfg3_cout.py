"""Config 3 at the per-GPU shard of the 8-GPU configuration (4M / 8 = 500 000 trajectories): CR3BP,
propagate_until(20) with c_output=True, continuous output evaluated on a [16, B] time grid (developer tool)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "heyoka.py_b200")); sys.path.insert(0, ROOT)
import numpy as np
import hy_b200 as hy
from hy_b200 import workloads as W

B = int(os.environ.get("QB", 500000))
for fp in (np.float64, np.float32):
    ic = W.cr3bp_ensemble(B).astype(fp)
    ta = hy.taylor_adaptive_batch(W.cr3bp_sys(0.01), ic, fp_type=fp)
    t0 = time.perf_counter()
    c_out, _ = ta.propagate_until(fp(20.0), c_output=True)
    t1 = time.perf_counter()
    ns = int(ta.propagate_res_arrays[3].sum())
    tq = np.repeat(np.linspace(0.0, 20.0, 16), B).reshape(16, B).astype(fp)
    out = c_out(tq)
    t2 = time.perf_counter()
    j0 = W.cr3bp_jacobi(ic.astype(np.float64))
    drift = max(float(np.max(np.abs((W.cr3bp_jacobi(out[q].astype(np.float64)) - j0) / j0))) for q in range(16))
    print(fp.__name__, ta._ctx.launch_info()["kernel_variant"], "B", B, "steps", ns, "n_steps(max)", c_out.n_steps,
          "propagate+record wall %.3f s (%.3g steps/s)" % (t1 - t0, ns / (t1 - t0)),
          "eval 16 x B wall %.3f s" % (t2 - t1), "max Jacobi drift over the grid %.2e" % drift, flush=True)
    del c_out, ta, out
