"""7- and 8-body problems (no register kernel: more than 15 pairs): tape interpreter vs run-time compiled
kernel (developer tool; calibrates the automatic choice)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "heyoka.py_b200")); sys.path.insert(0, ROOT)
import numpy as np
import hy_b200 as hy
from hy_b200 import workloads as W, model

B = int(os.environ.get("QB", 75776))
for nb in (7, 8):
    extra = nb - 6
    masses = list(W.OSS_MASSES) + [1e-9] * extra
    ic6 = W.oss_ensemble(B)
    add = []
    for e in range(extra):
        r = 45.0 + 7.0 * e
        add.append(np.array([r, 0.0, 0.3, 0.0, 2 * np.pi / np.sqrt(r), 0.0])[:, None] * np.ones((1, B)))
    ic = np.concatenate([ic6] + add, axis=0)
    sys_ = model.nbody(nb, masses=masses, Gconst=W.OSS_G)
    for mode in ("auto", "jit", "interp"):
        if mode == "jit":
            os.environ["HY_CUDA_JIT"] = "1"
        ta = hy.taylor_adaptive_batch(sys_, ic, compact_mode=(mode == "interp"))
        os.environ.pop("HY_CUDA_JIT", None)
        fl, lo = ta._dc.flops_per_step()
        for rep in range(2):
            ta.state[:] = ic; ta.set_time(0.0)
            ta.propagate_until(200.0)
            ms, _ = ta._ctx.last_timing()
            ns = int(ta.propagate_res_arrays[3].sum())
        li = ta._ctx.launch_info()
        print("%d bodies %-6s variant %4d G %2d T %3d threads %3d: %.3e steps/s, %.2f TFLOP/s" % (
            nb, mode, li["kernel_variant"], li["group"], li["traj_per_cta"], li["threads"], ns / (ms * 1e-3),
            ns * fl / (ms * 1e-3) / 1e12), flush=True)
