"""Register-resident CR3BP kernel vs the tape interpreter on config 3 (developer tool):
bitwise agreement and throughput, FP64 and FP32."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "heyoka.py_b200")); sys.path.insert(0, ROOT)
import numpy as np
import hy_b200 as hy
from hy_b200 import workloads as W

B = int(os.environ.get("QB", 1000000))
T_END = float(os.environ.get("QT", 20.0))
for fp in (np.float64, np.float32):
    ic = W.cr3bp_ensemble(B).astype(fp)
    res = {}
    for interp in ((0,) if os.environ.get("QSKIP_INTERP") else (1, 0)):
        os.environ["HY_CUDA_NO_CR3BP_REG"] = str(interp)
        ta = hy.taylor_adaptive_batch(W.cr3bp_sys(0.01), ic, fp_type=fp)
        print(fp.__name__, "interpreter" if interp else "register", ta._ctx.launch_info(), flush=True)
        for rep in range(2):
            ta.state[:] = ic
            ta.set_time(fp(0.0))
            ta.propagate_until(fp(T_END))
            ms, _ = ta._ctx.last_timing()
            ns = int(ta.propagate_res_arrays[3].sum())
            print("  rep", rep, "steps", ns, "ms %.2f" % ms, "steps/s %.4g" % (ns / (ms * 1e-3)), flush=True)
        res[interp] = (ta.state.copy(), ta.propagate_res_arrays[3].copy())
    if 1 in res:
        print("  bitwise equal:", np.array_equal(res[0][0], res[1][0]), np.array_equal(res[0][1], res[1][1]), flush=True)
