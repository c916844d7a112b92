"""Throughput of the register-resident N-body kernels for 3..6 bodies, FP64 and FP32
(developer tool): the first nb bodies of the outer Solar System."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "heyoka.py_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import hy_b200 as hy
import common

B = int(os.environ.get("QB", 125000))
for fp in (np.float64, np.float32):
    for nb in (3, 4, 5, 6):
        sys_ = hy.model.nbody(nb, masses=list(common.OSS_MASSES[:nb]), Gconst=common.OSS_G)
        ic = common.oss_ensemble(B, amp=1e-6)[: 6 * nb].astype(fp)
        for interp in (False, True):
            if interp:
                os.environ["HY_CUDA_NO_NBODY_REG"] = "1"
            else:
                os.environ.pop("HY_CUDA_NO_NBODY_REG", None)
            ta = hy.taylor_adaptive_batch(sys_, ic, fp_type=fp)
            fl, _ = ta._dc.flops_per_step()
            for rep in range(2):
                ta.state[:] = ic
                ta.set_time(fp(0.0))
                ta.propagate_until(fp(60.0))
            ms, _ = ta._ctx.last_timing()
            ns = int(ta.propagate_res_arrays[3].sum())
            li = ta._ctx.launch_info()
            print("%s nb=%d order=%d variant=%d G=%d T=%d: %.3g steps/s, %.2f TFLOP/s" % (
                np.dtype(fp).name, nb, ta.order, li["kernel_variant"], li["group"], li["traj_per_cta"],
                ns / (ms * 1e-3), ns * fl / (ms * 1e-3) / 1e12), flush=True)
os.environ.pop("HY_CUDA_NO_NBODY_REG", None)
