"""Throughput of config 5 (CR3BP with three terminal events) on the tape interpreter (developer tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "heyoka.py_b200")); sys.path.insert(0, ROOT)
import numpy as np
import hy_b200 as hy
from hy_b200 import workloads as W

mu, B = 0.01, int(os.environ.get("QB", 200000))
T_END = float(os.environ.get("QT", 50.0))
x, y, z = hy.make_vars("x", "y", "z")
evs = [(x - mu) ** 2 + y * y + z * z - 0.012 ** 2, (x - mu + 1.0) ** 2 + y * y + z * z - 0.012 ** 2,
       x * x + y * y + z * z - 25.0]
rng = np.random.default_rng(20251022)
ic = np.array([-0.80, 0.0, 0.0, 0.0, -0.6276410653920693, 0.0])[:, None] * np.ones((1, B))
ic[0] += rng.uniform(-1e-2, 1e-2, B)
ic[4] += rng.uniform(-1e-2, 1e-2, B)
for with_ev in (0, 1):
    kw = dict(t_events=[hy.t_event_batch(e) for e in evs]) if with_ev else {}
    ta = hy.taylor_adaptive_batch(W.cr3bp_sys(mu), ic, **kw)
    print("events" if with_ev else "no events", ta._ctx.launch_info(), flush=True)
    for rep in range(2):
        ta.state[:] = ic
        ta.set_time(0.0)
        if with_ev:
            ta.reset_cooldowns()
        ta.propagate_until(T_END)
        ms, _ = ta._ctx.last_timing()
        ns = int(ta.propagate_res_arrays[3].sum())
        print("  rep", rep, "steps", ns, "max/lane", int(ta.propagate_res_arrays[3].max()), "ms %.2f" % ms, "steps/s %.4g" % (ns / (ms * 1e-3)),
              "stopped by an event:", int((ta.propagate_res_arrays[0] > -10).sum()), flush=True)
