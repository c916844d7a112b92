"""Quick GPU sanity + throughput probe (developer tool, not the bench)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "heyoka.py_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import hy_b200 as hy
from hy_b200 import _cabi, decompose as D
import common

B = int(os.environ.get("QB", 148 * 11 * 4))
T_END = float(os.environ.get("QT", 100.0))
fp = np.float32 if os.environ.get("QFP", "64") == "32" else np.float64
print("fma peak f64 TF/s", _cabi.measure_fma_peak(0, 64), "f32", _cabi.measure_fma_peak(0, 32))
sys_ = common.oss_sys()
ic = common.oss_ensemble(B).astype(fp)
t0 = time.time()
ta = hy.taylor_adaptive_batch(sys_, ic, fp_type=fp)
print("ctor s", time.time() - t0, ta._ctx.launch_info())
fl, lo = ta._dc.flops_per_step()
for rep in range(3):
    ta.state[:] = ic; ta.set_time(fp(0.0))
    t0 = time.time()
    ta.propagate_until(fp(T_END))
    wall = time.time() - t0
    ms, nl = ta._ctx.last_timing()
    oc, mn, mx, ns = ta.propagate_res_arrays
    tot = int(ns.sum())
    print("rep", rep, "steps", tot, "kernel ms", ms, "wall s", wall, "steps/s (kernel)", tot / (ms * 1e-3),
          "TF/s", tot * fl / (ms * 1e-3) / 1e12, "smem load GB/s", tot * lo * np.dtype(fp).itemsize / (ms * 1e-3) / 1e9)
e0 = common.oss_energy(ic.astype(np.float64)); e1 = common.oss_energy(ta.state.astype(np.float64))
print("max energy drift", np.max(np.abs((e1 - e0) / e0)), "outcomes", set(oc.tolist()))
