"""The product's own multi-GPU paths on all visible devices:
 (a) hy.ensemble_propagate_until_batch: iteration i -> device i mod G, one shard of config 2 each;
 (b) ONE hy.taylor_adaptive_batch(..., device="all") holding the whole ensemble, lanes split
     over the devices.
Prints trajectory-steps/s by wall clock (host gather included) and checks the results against a
single-device run of shard 0."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "heyoka.py_b200")); sys.path.insert(0, ROOT)
import numpy as np
import hy_b200 as hy
from hy_b200 import workloads as W, _cabi

G = _cabi.device_count()
PER = int(os.environ.get("QB", 125000))
T_END = float(os.environ.get("QT", 2000.0))
sys_ = W.oss_sys()
ic_all = W.oss_ensemble(PER * G)
print("devices", G, "trajectories", PER * G, "t_end", T_END, flush=True)

# reference: shard 0 on device 0
ta0 = hy.taylor_adaptive_batch(sys_, ic_all[:, :PER].copy())
ta0.propagate_until(T_END)
ms0, _ = ta0._ctx.last_timing()
ns0 = int(ta0.propagate_res_arrays[3].sum())
print("1 device, shard 0: %.3e steps/s (kernel)" % (ns0 / ms0 * 1e3), flush=True)

# (a) ensemble over the devices
tmpl = hy.taylor_adaptive_batch(sys_, ic_all[:, :PER].copy())

def gen(ta, i):
    ta.state[:] = ic_all[:, i * PER:(i + 1) * PER]
    return ta

for rep in range(2):
    t0 = time.perf_counter()
    res = hy.ensemble_propagate_until_batch(tmpl, T_END, G, gen)
    dt = time.perf_counter() - t0
    ns = sum(int(r[0].propagate_res_arrays[3].sum()) for r in res)
    devs = [getattr(r[0], "_device", None) for r in res]
    print("ensemble_propagate_until_batch rep %d: n_iter %d, %.3e steps in %.3f s = %.3e steps/s (wall)" % (
        rep, G, ns, dt, ns / dt), flush=True)
assert np.array_equal(res[0][0].state, ta0.state), "iteration 0 differs from the single-device run"
print("iteration 0 bit-identical to the single-device run; per-iteration devices:", devs, flush=True)

# (b) one integrator over all devices
tb = hy.taylor_adaptive_batch(sys_, ic_all, device="all")
for rep in range(2):
    tb.state[:] = ic_all
    tb.set_time(0.0)
    t0 = time.perf_counter()
    tb.propagate_until(T_END)
    dt = time.perf_counter() - t0
    ns = int(tb.propagate_res_arrays[3].sum())
    print("taylor_adaptive_batch(device='all') rep %d: %.3e steps in %.3f s = %.3e steps/s (wall, host gather included)" % (
        rep, ns, dt, ns / dt), flush=True)
assert np.array_equal(tb.state[:, :PER], ta0.state), "lanes of shard 0 differ from the single-device run"
print("device='all': shard 0 bit-identical to the single-device run", flush=True)
