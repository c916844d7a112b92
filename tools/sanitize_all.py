"""Small runs of EVERY kernel path for compute-sanitizer (memcheck / racecheck / initcheck):
 1. tape interpreter with group sizes 1, 4, 16 (retire / refill from the work counter, spilled state
    jets in the global scratch `gjet`, event detection with the atomic event log), and its
    global-memory workspace fallback;
 2. register-resident N-body and CR3BP kernels (plain and FX builds: continuous output, grid,
    events on the register kernels);
 3. run-time compiled kernels (global and shared-memory workspace, events).
Usage: compute-sanitizer --tool memcheck|racecheck python tools/sanitize_all.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "heyoka.py_b200")); sys.path.insert(0, ROOT)
import numpy as np
import hy_b200 as hy
from hy_b200 import workloads as W

B = 37


def env(**kw):
    class _E:
        def __enter__(self):
            self.old = {k: os.environ.get(k) for k in kw}
            os.environ.update({k: str(v) for k, v in kw.items()})
        def __exit__(self, *a):
            for k, v in self.old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
    return _E()


def drive(name, ta, t1, grid_to):
    li = ta._ctx.launch_info()
    ta.step(write_tc=True)
    ta.propagate_until(t1, c_output=True)
    b = ta.state.shape[1]
    g = np.repeat(np.linspace(t1, grid_to, 4), b).reshape(4, b).astype(ta.state.dtype)
    ta.propagate_grid(g)
    print("%-46s variant %4d G %2d T %3d smem %6d ws_in_smem %d: ok, outcome[0] %d" % (
        name, li["kernel_variant"], li["group"], li["traj_per_cta"], li["smem_bytes"], li["ws_in_smem"],
        int(ta.propagate_res_arrays[0][0])), flush=True)


x, v = hy.make_vars("x", "v")
pend = [(x, v), (v, -9.8 * hy.sin(x))]
pic = np.stack([np.linspace(0.05, 1.0, B), np.linspace(-0.2, 0.2, B)])
evs = lambda: dict(t_events=[hy.t_event_batch(v * v - 1.0, direction=hy.event_direction.positive)],
                   nt_events=[hy.nt_event_batch(x, lambda ta, t, d, i: None)])

# 1. interpreter
for G in (1, 4, 16):
    with env(HY_CUDA_GROUP=G):
        drive("interpreter G=%d pendulum + events" % G, hy.taylor_adaptive_batch(pend, pic, compact_mode=True, **evs()), 2.0, 3.0)
        vs = hy.var_ode_sys(W.kepler_j2_sys(), hy.var_args.vars)
        drive("interpreter G=%d Kepler+J2 variational" % G,
              hy.taylor_adaptive_batch(vs, W.kepler_j2_ensemble(B), compact_mode=True), 900.0, 1300.0)
with env(HY_CUDA_GROUP=16, HY_CUDA_NO_NBODY_REG=1):
    drive("interpreter G=16 N-body (spilled state jets)", hy.taylor_adaptive_batch(W.oss_sys(), W.oss_ensemble(B, amp=1e-3)), 3.0, 4.0)
with env(HY_CUDA_FORCE_GLOBAL_WS=1):
    drive("interpreter global-workspace fallback", hy.taylor_adaptive_batch(pend, pic, compact_mode=True, **evs()), 2.0, 3.0)
# 2. register-resident kernels
drive("N-body register kernel", hy.taylor_adaptive_batch(W.oss_sys(), W.oss_ensemble(B, amp=1e-3)), 3.0, 4.0)
for fp in (np.float64, np.float32):
    drive("CR3BP register kernel %s" % fp.__name__,
          hy.taylor_adaptive_batch(W.cr3bp_sys(0.01), W.cr3bp_ensemble(B).astype(fp), fp_type=fp), fp(1.5), fp(2.0))
xx, yy, zz = hy.make_vars("x", "y", "z")
cev = [hy.t_event_batch((xx - 0.01) ** 2 + yy * yy + zz * zz - 0.2 ** 2), hy.t_event_batch(xx * xx + yy * yy + zz * zz - 1.3 ** 2)]
drive("CR3BP register kernel + event tape", hy.taylor_adaptive_batch(W.cr3bp_sys(0.01), W.cr3bp_ensemble(B), t_events=cev), 1.5, 2.0)
with env(HY_CUDA_JIT_EVT=2):   # events as generated code on both lanes, workspace in the global slab (T = 128 layout)
    drive("CR3BP register kernel + generated event code", hy.taylor_adaptive_batch(W.cr3bp_sys(0.01), W.cr3bp_ensemble(B), t_events=cev), 1.5, 2.0)
    with env(HY_CUDA_EVT_GLOBAL_WS=0):
        drive("CR3BP register kernel + generated event code, workspace in the column",
              hy.taylor_adaptive_batch(W.cr3bp_sys(0.01), W.cr3bp_ensemble(B), t_events=cev), 1.5, 2.0)
    v_ = lambda s_: hy.expression(s_)
    nev = [hy.nt_event_batch((v_("x_1") - v_("x_2")) ** 2 + (v_("y_1") - v_("y_2")) ** 2 - 60.0, lambda ta, t, d, i: None),
           hy.nt_event_batch((v_("x_1") + 0.5) * (v_("y_1") - 0.25), lambda ta, t, d, i: None)]
    drive("N-body register kernel + generated event code (16 lanes)",
          hy.taylor_adaptive_batch(W.oss_sys(), W.oss_ensemble(B, amp=1e-3), nt_events=nev), 3.0, 4.0)
# 3. run-time compiled kernels
with env(HY_CUDA_JIT=1):
    drive("compiled kernel, shared-memory workspace + events", hy.taylor_adaptive_batch(pend, pic, **evs()), 2.0, 3.0)
vs = hy.var_ode_sys(W.kepler_j2_sys(), hy.var_args.vars)
drive("compiled kernel, global workspace (config 4)", hy.taylor_adaptive_batch(vs, W.kepler_j2_ensemble(B)), 900.0, 1300.0)
print("all paths done")
