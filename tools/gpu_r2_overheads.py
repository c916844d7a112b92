"""Two overhead figures the round-1 review asked for:
 (a) a 10^5-lane run with callback.angle_reducer (a device-side post-step op) against the callback-free run;
 (b) a 10 000-iteration B = 32 thread ensemble: wall time and the share of it spent creating contexts
     (hy_create / hy_clone; the context pool hands idle contexts back to new iterations)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "heyoka.py_b200")); sys.path.insert(0, ROOT)
import numpy as np
import hy_b200 as hy
from hy_b200 import _cabi, workloads as W

# (a)
x, v = hy.make_vars("x", "v")
pend = [(x, v), (v, -9.8 * hy.sin(x))]
B = 100000
rng = np.random.default_rng(1)
ic = np.stack([rng.uniform(2.9, 3.1, B), rng.uniform(2.0, 3.0, B)])   # circulating pendulums: the angle grows
res = {}
for name, cb in (("plain", None), ("angle_reducer", hy.callback.angle_reducer([x]))):
    ta = hy.taylor_adaptive_batch(pend, ic)
    best = 1e9
    for rep in range(3):
        ta.state[:] = ic; ta.set_time(0.0)
        t0 = time.perf_counter()
        ta.propagate_until(50.0, **({"callback": cb} if cb is not None else {}))
        best = min(best, time.perf_counter() - t0)
    ms, _ = ta._ctx.last_timing()
    res[name] = (best, ms, int(ta.propagate_res_arrays[3].sum()), float(np.abs(ta.state[0]).max()))
    print("(a) %-14s wall %.3f s  kernel %.1f ms  steps %.3e  max |x| %.2f" % ((name,) + res[name]), flush=True)
print("(a) angle_reducer / plain: wall %.3f x, kernel %.3f x" % (res["angle_reducer"][0] / res["plain"][0],
                                                                res["angle_reducer"][1] / res["plain"][1]), flush=True)

# (b)
sys_ = W.oss_sys()
tmpl = hy.taylor_adaptive_batch(sys_, W.oss_ensemble(32))
t_make = [0.0, 0]
orig_clone, orig_init = _cabi.Context.clone, _cabi.Context.__init__
def timed_clone(self, *a, **k):
    t0 = time.perf_counter(); r = orig_clone(self, *a, **k); t_make[0] += time.perf_counter() - t0; t_make[1] += 1; return r
def timed_init(self, *a, **k):
    t0 = time.perf_counter(); orig_init(self, *a, **k); t_make[0] += time.perf_counter() - t0; t_make[1] += 1
_cabi.Context.clone, _cabi.Context.__init__ = timed_clone, timed_init
def gen(ta, i):
    ta.state[0, :] += 1e-9 * i
    return ta
n_iter = 10000
t0 = time.perf_counter()
ret = hy.ensemble_propagate_until_batch(tmpl, 100.0, n_iter, gen)
wall = time.perf_counter() - t0
print("(b) %d iterations x B = 32: wall %.2f s (%.2f ms / iteration); contexts created %d, %.3f s in hy_create / hy_clone = %.1f %% of the wall time (summed over worker threads)"
      % (n_iter, wall, 1e3 * wall / n_iter, t_make[1], t_make[0], 100.0 * t_make[0] / wall), flush=True)
assert len(ret) == n_iter
