"""Config 5 throughput: CR3BP ensemble with three stopping terminal events, register-resident
kernel + event tape against the tape interpreter (HY_CUDA_NO_REG_EVENTS=1)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "heyoka.py_b200")); sys.path.insert(0, ROOT)
import numpy as np
import hy_b200 as hy
from hy_b200 import workloads as W

B = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
T = float(sys.argv[2]) if len(sys.argv) > 2 else 50.0
mu = 0.01
x, y, z = hy.make_vars("x", "y", "z")
evs = [(x - mu) ** 2 + y * y + z * z - 0.012 ** 2, (x - mu + 1.0) ** 2 + y * y + z * z - 0.012 ** 2,
       x * x + y * y + z * z - 25.0]
rng = np.random.default_rng(20251022)
ic = np.array([-0.80, 0, 0, 0, -0.6276410653920693, 0])[:, None] * np.ones((1, B))
ic[0] += rng.uniform(-1e-2, 1e-2, B); ic[4] += rng.uniform(-1e-2, 1e-2, B)
os.environ["HY_CUDA_EVENT_STATS"] = "1"
for mode in ("reg", "interp", "noevents"):
    if mode == "interp":
        os.environ["HY_CUDA_NO_REG_EVENTS"] = "1"
    kw = {} if mode == "noevents" else {"t_events": [hy.t_event_batch(e) for e in evs]}
    ta = hy.taylor_adaptive_batch(W.cr3bp_sys(mu), ic, **kw)
    li = ta._ctx.launch_info()
    os.environ.pop("HY_CUDA_NO_REG_EVENTS", None)
    for rep in range(2):
        ta.state[:] = ic; ta.set_time(0.0)
        if kw: ta.reset_cooldowns()
        ta.propagate_until(T)
        ms, nl = ta._ctx.last_timing()
        ns = ta.propagate_res_arrays[3].sum()
    oc = ta.propagate_res_arrays[0]
    print(mode, "variant", li["kernel_variant"], "T", li["traj_per_cta"], "threads", li["threads"], "smem", li["smem_bytes"],
          "stats", ta._ctx.event_stats() if mode == "reg" else "", "steps %.3e" % ns, "kernel ms %.1f" % ms, "steps/s %.3e" % (ns / ms * 1e3), "hits", int((oc > -10).sum()), flush=True)
