#!/usr/bin/env python
"""bench.py - ensemble trajectory-steps/s of the batch Taylor integrator.

Workload (BASELINE.json configs[1]): outer Solar System 6-body ensemble
(model.nbody(6), masses/ICs of doc/notebooks/Outer Solar System.ipynb), FP64,
tol = eps (order 20), ICs perturbed by 1e-12 and recentred, propagate_until
1e4 yr.  The 1M-trajectory ensemble is sharded 125 000 trajectories per GPU
(weak scaling: per-GPU work fixed, no data-path collective).

One "step" = one pass of the hot path over the whole shard on every GPU: a
propagate_for(horizon/10) segment, continuing from the previous step, so the
default W=3 + K=7 steps cover exactly the configuration's 1e4 yr horizon.

  value : whole-job trajectory-steps/s with the ICs resident in HBM
          (device->device reset of the state, then the persistent kernel).
  e2e   : the same through the public Python API (hy.taylor_adaptive_batch),
          host buffers, H2D + D2H inside the timed region.
  roofline : the propagate kernel against the measured FP64 FMA peak
          (hy_measure_fma_peak) and HBM peak (MEASURED_PEAKS.json).
  cpu_baseline : the C oracle (port of the reference algorithm) on the host
          cores, on a bounded sample of the same workload.

`--impl reference` times the CPU implementation only (the reference's own
arithmetic, heyoka C++ 7.11, is not installable here: DESIGN.md).
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "heyoka.py_b200"))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "ensemble trajectory-steps/s (FP64 N-body)"
UNIT = "trajectory-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=7)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5],
                    help="BASELINE.json configuration (2 = the headline outer Solar System ensemble)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: --total trajectories split over the GPUs (default: every N runs THE "
                         "configuration); weak: --traj-per-gpu on every GPU")
    ap.add_argument("--total", type=int, default=0, help="trajectories in total (strong scaling)")
    ap.add_argument("--traj-per-gpu", type=int, default=0, help="trajectories per GPU (weak scaling)")
    ap.add_argument("--horizon", type=float, default=0.0, help="final time of the configuration (0: its own)")
    ap.add_argument("--segment", type=float, default=0.0,
                    help="time per bench step (default horizon/10)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0,
                    help="target duration of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-order22", action="store_true",
                    help="config 2: skip the extra measurement of the order-22 high-accuracy build")
    return ap.parse_args()


# trajectories of the configuration as BASELINE.json states it, and the largest shard one GPU takes
# (cfg 3: the continuous output of 4M lanes is ~400 GB - it exists only sharded over 8 GPUs;
#  cfgs 4 and 5 ran an eighth of their 10^6 lanes per GPU while they were on the tape interpreter; with
#  the generated / register-resident kernels of round 2 one GPU takes the whole configuration)
CFG_TOTAL = {2: 1000000, 3: 4000000, 4: 1000000, 5: 1000000}
CFG_MAX_PER_GPU = {2: 1000000, 3: 500000, 4: 1000000, 5: 1000000}


def shard_size(args, world, rank):
    from hy_b200.shard import shard_bounds

    if args.scaling == "weak":
        return args.traj_per_gpu or CFG_MAX_PER_GPU[args.config] // (1 if args.config == 3 else 8)
    total = args.total or min(CFG_TOTAL[args.config], CFG_MAX_PER_GPU[args.config] * world)
    lo, hi = shard_bounds(total, rank, world)
    return hi - lo


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return rank, local, world


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(np.max(mx)) if mx else None,
            "power_w_max": float(np.max(pw)) if pw else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


def ncu_traffic(trajectories):
    """dram__bytes_read.sum + dram__bytes_write.sum of the propagate kernel, from the committed
    `ncu --set full` capture of the SAME kernel at the bench geometry (profiles/r02_ncu_full_cfg2.txt:
    10^6 trajectories, 148 CTAs x 256 threads, one 400-yr segment) - a file read, not a measurement of the
    timed launch.  The kernel's HBM traffic is the per-lane state / time / result vectors in and out: it
    does not grow with the number of steps and is proportional to the trajectories of the launch (scaled
    here when a rank holds fewer than 10^6).  None if the summary is missing."""
    try:
        tot, mult = 0.0, {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for ln in open(os.path.join(ROOT, "profiles", "r02_ncu_full_cfg2.txt")):
            if ln.startswith("dram__bytes_read.sum [") or ln.startswith("dram__bytes_write.sum ["):
                unit = ln.split("[")[1].split("]")[0]
                tot += float(ln.split("=")[1]) * mult[unit]
        return tot * trajectories / 1e6 if tot else None
    except Exception:
        return None


# --------------------------------------------------------------------------------------
# The configurations of BASELINE.json: system, seeded ICs, horizon,
# what one bench "step" is.  --config 2 (default) is the headline.
# --------------------------------------------------------------------------------------
def make_config(cfg, traj, rank, horizon, segment):
    import hy_b200 as hy
    from hy_b200 import workloads as W

    c = {"cfg": cfg, "fp": np.float64, "events": None, "pars": None, "c_output": False,
         "grid_eval": 0, "reset_each_step": False}
    seed = 20251017 + cfg + 7919 * rank
    if cfg == 2:
        c.update(name="outer Solar System 6-body (model.nbody(6)) ensemble, FP64, tol=eps (order 20), "
                      "ICs x(1+U(-1e-12,1e-12)) recentred",
                 sys=W.oss_sys(), ic=W.oss_ensemble(traj, seed=seed), horizon=horizon or 1e4,
                 unit_t="yr")
        c["segment"] = segment or c["horizon"] / 10.0
        c["step_desc"] = "propagate_for({:g} yr) over the whole shard, continuing from the previous step".format(c["segment"])
        c["check"] = lambda ic, st: float(np.max(np.abs((W.oss_energy(st) - W.oss_energy(ic)) / W.oss_energy(ic))))
        c["check_name"], c["check_tol"] = "max_rel_energy_drift", 1e-10
    elif cfg == 3:
        c.update(name="CR3BP (model.cr3bp, mu=0.01) ensemble, FP64, tol=eps (order 20), ICs base + U(-1e-3,1e-3), "
                      "propagate_until(20) with c_output=True, then c_out on a [16, B] time grid",
                 sys=W.cr3bp_sys(0.01), ic=W.cr3bp_ensemble(traj, seed=seed), horizon=horizon or 20.0,
                 unit_t="", c_output=True, grid_eval=16, reset_each_step=True)
        c["segment"] = c["horizon"]
        c["step_desc"] = "ICs reset on the device, propagate_until(20, c_output=True), 16 x B dense evaluations"
        c["check"] = lambda ic, st: float(np.max(np.abs((W.cr3bp_jacobi(st) - W.cr3bp_jacobi(ic)) / W.cr3bp_jacobi(ic))))
        c["check_name"], c["check_tol"] = "max_rel_jacobi_drift", 1e-10
    elif cfg == 4:
        vs = hy.var_ode_sys(W.kepler_j2_sys(), hy.var_args.vars)
        ic6 = W.kepler_j2_ensemble(traj, seed=seed)
        ic = np.zeros((42, traj))
        ic[:6] = ic6
        ic[6:] = vs._initial_var_state(np.float64)[:, None]
        c.update(name="Kepler+J2 orbit ensemble with first-order variational equations (var_ode_sys, 42 state "
                      "variables), FP64, tol=eps (order 20), LEO ICs",
                 sys=vs.sys, ic=ic, horizon=horizon or 6e4, unit_t="s")
        c["segment"] = segment or c["horizon"] / 10.0
        c["step_desc"] = "propagate_for({:g} s) over the whole shard, continuing from the previous step".format(c["segment"])
        c["check"] = lambda ic, st: float(np.max(np.abs(
            (W.kepler_j2_energy(st[:6]) - W.kepler_j2_energy(ic[:6])) / W.kepler_j2_energy(ic[:6]))))
        c["check_name"], c["check_tol"] = "max_rel_energy_drift", 1e-10
    elif cfg == 5:
        mu = 0.01
        x, y, z = hy.make_vars("x", "y", "z")
        evs = [(x - mu) ** 2 + y * y + z * z - 0.012 ** 2,
               (x - mu + 1.0) ** 2 + y * y + z * z - 0.012 ** 2,
               x * x + y * y + z * z - 5.0 ** 2]
        rng = np.random.default_rng(seed)
        ic = np.array([-0.80, 0.0, 0.0, 0.0, -0.6276410653920693, 0.0])[:, None] * np.ones((1, traj))
        ic[0] += rng.uniform(-1e-2, 1e-2, traj)
        ic[4] += rng.uniform(-1e-2, 1e-2, traj)
        c.update(name="CR3BP (mu=0.01) ensemble with three stopping terminal events (collision spheres R=0.012 "
                      "around both primaries, escape sphere R=5), FP64, tol=eps (order 20), planar chaotic family",
                 sys=W.cr3bp_sys(mu), ic=ic, horizon=horizon or 2000.0, unit_t="", events=evs,
                 reset_each_step=True)
        c["segment"] = segment or c["horizon"] / 10.0
        c["horizon_step"] = c["segment"]
        c["step_desc"] = ("ICs reset on the device, propagate_until({:g}) with event detection "
                          "(1/10 of the configuration's t = {:g})".format(c["segment"], c["horizon"]))
        c["check"] = lambda ic, st: float(np.nanmax(np.abs((W.cr3bp_jacobi(st) - W.cr3bp_jacobi(ic)) / W.cr3bp_jacobi(ic))))
        c["check_name"], c["check_tol"] = "max_rel_jacobi_drift", 1e-6
    else:
        raise SystemExit("unknown --config {}".format(cfg))
    return c


def cpu_baseline(cfgd, order, target_s, rank_seed=0):
    """Time the CPU port on all host cores on a bounded sample: the SIMD-batched,
    multithreaded restatement (oracle/hy_baseline_simd.c: 8 lanes in lock-step per
    thread, OpenMP over batches) - the shape of the reference's own CPU ensemble."""
    from hy_b200 import decompose as D
    from oracle.c_oracle import simd_propagate_until

    sys_, horizon = cfgd["sys"], cfgd["segment"] if cfgd["reset_each_step"] else cfgd["horizon"]
    dc = D.decompose(sys_, order, events=cfgd["events"] or ())
    # torchrun exports OMP_NUM_THREADS=1: ask for every host core explicitly.
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    full = cfgd["ic"]
    ntot = full.shape[1]

    def sample(k, off):
        idx = (np.arange(k) + off) % ntot
        return np.ascontiguousarray(full[:, idx])

    # Calibrate on tiny runs (the first one also pays thread start-up), then size the
    # sample for ~target_s seconds; grow it once more if it still came out short.
    cal_traj, cal_h = 8 * cores, min(horizon, horizon / 25.0 if cfgd["cfg"] == 2 else horizon)
    rate, steps_per_traj = 0.0, 1.0
    for _ in range(2):
        t0 = time.perf_counter()
        _, ns = simd_propagate_until(dc, sample(cal_traj, 11 * rank_seed), cal_h, nthreads=cores)
        dt = time.perf_counter() - t0
        rate = float(ns.sum()) / max(dt, 1e-9)
        steps_per_traj = float(ns.mean()) * horizon / cal_h
    for _ in range(3):
        want = rate * target_s
        traj = max(8 * cores, int(want / steps_per_traj) // 8 * 8)
        h = horizon
        if traj * steps_per_traj > 1.5 * want:
            # Keep at least one full SIMD batch per thread: shorten the horizon instead.
            h = max(cal_h, horizon * want / (steps_per_traj * traj))
        t0 = time.perf_counter()
        _, ns = simd_propagate_until(dc, sample(traj, 13 * rank_seed + 5), h, nthreads=cores)
        dt = time.perf_counter() - t0
        rate = float(ns.sum()) / max(dt, 1e-9)
        if dt >= 0.5 * target_s:
            break
    val = float(ns.sum()) / dt
    extra = ""
    if cfgd["events"]:
        extra = "; the event functions are integrated (they enter the step-size norms) but the CPU port runs no root finding"
    if cfgd["c_output"]:
        extra = "; the CPU port does not record the continuous output"
    return {
        "value": val,
        "unit": UNIT,
        "cores": cores,
        "kind": "port",
        "sample": "{} trajectories x t = {:g} = {} steps in {:.1f} s (SIMD-batched C port of the reference "
                  "algorithm: 8 lanes/thread in lock-step, OpenMP over batches, built {}{})".format(
                      traj, h, int(ns.sum()), dt, _c_oracle_flags(), extra),
    }, dt, int(ns.sum())


def _c_oracle_flags():
    from oracle import c_oracle

    return c_oracle.SIMD_FLAGS


def run_reference(args):
    rank, local, world = dist_env()
    if rank != 0:
        return
    from hy_b200 import decompose as D

    cfgd = make_config(args.config, 4096, 0, args.horizon, args.segment)
    order = D.taylor_order(float(np.finfo(np.float64).eps))
    tot_steps, tot_t = 0, 0.0
    per = max(2.0, min(args.cpu_seconds, 120.0 / max(1, args.steps + args.warmup)))
    info = None
    for i in range(args.warmup + args.steps):
        info, dt, ns = cpu_baseline(cfgd, order, per, rank_seed=i)
        if i >= args.warmup:
            tot_steps += ns
            tot_t += dt
    val = tot_steps / tot_t
    info["value"] = val
    line = {
        "impl": "reference",
        "metric": metric_name(args.config), "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(1, args.steps),
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {
            "workload": cfgd["name"] + "; bounded CPU sample",
            "config_id": args.config,
            "note": "reference arithmetic (heyoka C++ 7.11) not installable here: C oracle port "
                    "of the same algorithm on all host threads",
        },
        "cpu_baseline": info,
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


def metric_name(cfg):
    return {2: METRIC, 3: "ensemble trajectory-steps/s (FP64 CR3BP, c_output)",
            4: "ensemble trajectory-steps/s (FP64 Kepler+J2 variational)",
            5: "ensemble trajectory-steps/s (FP64 CR3BP, terminal events)"}[cfg]


_REAL_STDOUT = None


def _quiet_stdout():
    """Route everything libraries print on stdout (e.g. NCCL's version banner) to
    stderr: stdout carries exactly ONE line, the bench JSON."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


def main():
    args = parse()
    _quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
        return
    rank, local, world = dist_env()
    import torch
    import torch.distributed as dist

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    import ctypes as C
    import hy_b200 as hy
    from hy_b200 import _cabi, decompose as D
    from hy_b200.shard import reduce_throughput

    B = shard_size(args, world, rank)
    cfgd = make_config(args.config, B, rank, args.horizon, args.segment)
    sys_, ic, seg = cfgd["sys"], cfgd["ic"], cfgd["segment"]
    fp = np.float64
    eps = float(np.finfo(fp).eps)
    order = D.taylor_order(eps)
    n = ic.shape[0]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm ("value") ----------------
    evs = cfgd["events"] or []
    dc = D.decompose(sys_, order, events=evs)
    # (event-carrying systems: the ODE-only tape and the event tape go along, as the front end does -
    #  a matched ODE then keeps its register-resident kernel, hy_create2)
    evt = D.decompose_event_tape(evs, [l.name for l, _ in sys_], order) if evs else None
    dc_ode = D.decompose(sys_, order) if evt is not None else None
    ctx = _cabi.Context(dc, 64, B, eps, False, device=local, n_tevents=len(evs),
                        ev_dir=[0] * len(evs) if evs else None, ev_cooldown=[-1.0] * len(evs) if evs else None,
                        dc_ode=dc_ode, evt=evt)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    li = ctx.launch_info()
    d_ic = torch.from_numpy(ic).to(dev)
    d_zero = torch.zeros(B, dtype=torch.float64, device=dev)
    tf = np.full(B, seg, dtype=fp)
    oc = np.zeros(B, dtype=np.int64)
    nst = np.zeros(B, dtype=np.uint64)
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    K = cfgd["grid_eval"]
    tq = np.repeat(np.linspace(0.0, cfgd["horizon"], K), B).reshape(K, B) if K else None
    gout = np.empty((K, n, B)) if K else None
    # (device-resident arm: the query times and the evaluated grid stay in HBM)
    d_tq = torch.from_numpy(tq).to(dev) if K else None
    d_gout = torch.empty((K, n, B), dtype=torch.float64, device=dev) if K else None

    def reset_dev():
        _cabi.check(_cabi.lib().hy_upload_dev(
            ctx._ctx, C.c_void_p(d_ic.data_ptr()), None, C.c_void_p(d_zero.data_ptr()),
            C.c_void_p(d_zero.data_ptr())))
        if evs:
            ctx.reset_cooldowns(-1)

    launches = [0]

    def one_step():
        l2_flush.zero_()
        if cfgd["reset_each_step"]:
            reset_dev()
            ctx.propagate(tf, 0, 0, None, 0, 1 if cfgd["c_output"] else 0, oc, None, None, nst)
        else:
            ctx.propagate(tf, 1, 0, None, 0, 0, oc, None, None, nst)  # propagate_for(seg)
        ms, nl = ctx.last_timing()
        launches[0] += nl
        if cfgd["c_output"]:
            rec = ctx.cout_detach()
            rec.eval_dev(d_tq.data_ptr(), K, d_gout.data_ptr())  # K x B dense evaluations (one more launch)
            launches[0] += 2               # chunk directory + evaluation kernels
            rec.close()                    # the pool goes back to the context
        return int(nst.sum()), ms

    reset_dev()
    for _ in range(args.warmup):
        one_step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    launches[0] = 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    t0 = time.perf_counter()
    tot_steps, kern_ms = 0, 0.0
    for _ in range(args.steps):
        s_, ms = one_step()
        tot_steps += s_
        kern_ms += ms
    ev1.record(stream)
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    gpu_launches = launches[0]
    if not evs:
        assert np.all(oc == int(hy.taylor_outcome.time_limit)), "some trajectories did not finish"
    # Sanity of what was timed (outside the timed region): the final state of the device arm
    # conserves the invariant of every trajectory.
    st_end = np.empty_like(ic)
    ctx.download(state=st_end)
    drift = cfgd["check"](ic, st_end)
    assert drift < cfgd["check_tol"], drift

    value, t_max, steps_all = reduce_throughput(dev_ms * 1e-3, tot_steps, dist if world > 1 else None, dev)

    # ---------------- the reference's own benchmark setting (config 2 only) ----------------
    # doc/notebooks/ensemble_batch_perf.ipynb:233 integrates the outer Solar System with
    # high_accuracy=True, tol=1e-18 (order 22): the order-22 build of the same kernel, timed on the same
    # shard for one segment (reported beside the headline, not instead of it).
    hi_acc = None
    if args.config == 2 and not args.no_order22:
        o22 = D.taylor_order(1e-18)
        dc22 = D.decompose(sys_, o22)
        c22 = _cabi.Context(dc22, 64, B, 1e-18, True, device=local)
        c22.set_stream(stream.cuda_stream)
        nst22 = np.zeros(B, dtype=np.uint64)
        best = None
        for rep in range(2):  # (first call: lazy module load)
            _cabi.check(_cabi.lib().hy_upload_dev(c22._ctx, C.c_void_p(d_ic.data_ptr()), None,
                                                  C.c_void_p(d_zero.data_ptr()), C.c_void_p(d_zero.data_ptr())))
            c22.propagate(tf, 1, 0, None, 0, 0, oc, None, None, nst22)
            ms22, _ = c22.last_timing()
            best = (int(nst22.sum()), ms22)
        fl22, _ = dc22.flops_per_step()
        hi_acc = {
            "workload": "same shard, tol=1e-18 (order {}), high_accuracy=True, one segment of {:g}".format(o22, seg),
            "value": best[0] / (best[1] * 1e-3), "unit": UNIT, "kernel_variant": c22.launch_info()["kernel_variant"],
            "flops_per_trajectory_step": fl22, "achieved_tflops": best[0] * fl22 / (best[1] * 1e-3) / 1e12,
        }
        c22.close()
        barrier()

    # ---------------- end-to-end arm through the public API ----------------
    e2e = None
    if not args.no_e2e:
        kw = {}
        if evs:
            kw["t_events"] = [hy.t_event_batch(e) for e in evs]
        ta = hy.taylor_adaptive_batch(sys_, ic, device=local, **kw)
        ta._ctx.set_stream(stream.cuda_stream)

        def e2e_step():
            if cfgd["reset_each_step"]:
                ta.state[:] = ic
                ta.set_time(0.0)
                if evs:
                    ta.reset_cooldowns()
                c_out, _ = ta.propagate_until(seg, c_output=cfgd["c_output"])
                if cfgd["c_output"]:
                    c_out(tq)
                    del c_out
            else:
                ta.propagate_for(seg)
            return int(ta.propagate_res_arrays[3].sum())

        ta.state[:] = ic
        ta.set_time(0.0)
        # warm-up (the device arm above already warmed the GPU; config 3's short steps get the full count: the
        # first ones allocate the page-locked result buffer and grow the recorder pool)
        for _ in range(min(args.warmup, 3 if cfgd["c_output"] else 1)):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        e_steps = 0
        for _ in range(args.steps):
            # host (pinned) state -> H2D, kernel, D2H of state/time/results
            e_steps += e2e_step()
        barrier()
        e_wall = time.perf_counter() - t0
        e_val, _, _ = reduce_throughput(e_wall, e_steps, dist if world > 1 else None, dev)
        m = dc.n_par
        e2e = {
            "value": e_val,
            "unit": UNIT,
            "h2d_bytes_per_step": int(B * 8 * (n + m + 2 + 1 + K)),      # state, pars, t_hi, t_lo, t_final (+ query times)
            "d2h_bytes_per_step": int(B * 8 * (n + 3 + 4 + K * n)),     # state, t_hi, t_lo, last_h, results (+ dense output)
        }
        del ta

    if rank == 0:
        fl, lo = dc.flops_per_step()
        steps_rank = tot_steps
        k_s = kern_ms * 1e-3
        fma_peak = _cabi.measure_fma_peak(local, 64)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback"
        achieved_tf = steps_rank * fl / k_s / 1e12
        rec_bytes = 8.0 * (n * (order + 1) + 2) * (steps_rank / args.steps) if cfgd["c_output"] else 0.0
        alg_bytes = 8.0 * (2 * dc.n_state + dc.n_par + 4 + 4) * B + rec_bytes     # per launch
        variant = li.get("kernel_variant", 0)
        kname = {0: " (tape interpreter)", 203: " (register-resident CR3BP jets, hy_cr3bp_reg.cuh)",
                 1000: " (generated from the tape and compiled at hy_create time with NVRTC, hy_jit.hpp)"}.get(
            variant, " (register-resident N-body jets, hy_nbody_reg.cuh)")
        roof = {
            "bound": "fp64-fma",
            "achieved": achieved_tf,
            "peak": fma_peak,
            "unit": "TFLOP/s",
            "frac": achieved_tf / fma_peak if fma_peak else None,
            "peak_source": "measured DFMA microbenchmark (hy_measure_fma_peak) on this GPU; "
                           "MEASURED_PEAKS.json holds no FP64 figure",
            "kernel": "hy::propagate_kernel<double,{},true,{}>{}".format(li["group"], variant, kname),
            "kernel_ms_per_launch": kern_ms / args.steps,
            "flops_per_trajectory_step": fl,
            "traffic": ncu_traffic(B) if args.config == 2 else None,
            "traffic_note": "ncu --set full capture of the same kernel at the bench geometry (10^6 trajectories, "
                            "profiles/r02_ncu_full_cfg2.txt: 691 MB per launch), scaled to this rank's trajectories; "
                            "the kernel's HBM traffic is the per-lane vectors in/out and does not depend on the "
                            "number of steps" if args.config == 2 else None,
            "hbm": {
                "algorithmic_bytes_per_launch": alg_bytes,
                "achieved_gbs": alg_bytes * args.steps / k_s / 1e9,
                "peak_gbs": hbm_peak, "peak_source": hbm_src,
                "note": "jets stay on chip: HBM carries the initial/final state"
                        + (" and the continuous-output record" if cfgd["c_output"] else ""),
            },
            "smem": {
                "operand_loads_per_trajectory_step": lo,
                "achieved_gbs": steps_rank * lo * 8 / k_s / 1e9,
                "peak_gbs": 128.0 * li["n_sm"] * (clocks.get("sm_mhz") or 1965.0) * 1e6 / 1e9,
                "note": "operand loads of the tape (what the tape INTERPRETER reads from shared memory; "
                        "with kernel_variant > 0 the convolution operands are registers and shared memory "
                        "carries only the state jets)",
            },
        }
        if variant == 1000 and not li["ws_in_smem"]:
            # Run-time compiled kernel with the jets in global memory: the operands of every recurrence
            # stream through L1 / L2 / HBM, so memory - not the FP64 pipe - is the roof.  Algorithmic bytes per
            # trajectory-step: every jet that is read with history is fetched once per order (k + 1 rows at
            # order k, however many ops share it) and every workspace row is written once:
            # 8 x (n_jets x p (p + 1) / 2 + n_rows).  ncu (profiles/r02_ncu_jit_cfg4.txt) measures 277 kB read +
            # 54 kB written per step at the DRAM pins for config 4 (229 kB algorithmic): the caches do not hold a
            # jet from one op that uses it to the next.
            n_jets = sum(1 for u in dc.uvars if u.jet and u.row is not None)
            abytes = 8.0 * (n_jets * order * (order + 1) / 2 + dc.n_rows)
            roof = {
                "bound": "hbm", "achieved": steps_rank * abytes / k_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": steps_rank * abytes / k_s / 1e9 / hbm_peak, "peak_source": hbm_src,
                "kernel": roof["kernel"], "kernel_ms_per_launch": roof["kernel_ms_per_launch"],
                "algorithmic_bytes_per_trajectory_step": abytes,
                "traffic": 331e3 * steps_rank / args.steps if args.config == 4 else None,
                "traffic_note": "DRAM bytes per launch = 331 kB per trajectory-step (ncu --set full capture of the same "
                                "kernel, profiles/r02_ncu_jit_cfg4.txt) x the steps of one launch; operand loads of the "
                                "tape: {:.0f} kB per step (L1 serves 38 % of them, L2 20 %)".format(8e-3 * (lo + dc.n_rows))
                if args.config == 4 else None,
                "fp64": {"achieved_tflops": achieved_tf, "peak_tflops": fma_peak,
                         "frac": achieved_tf / fma_peak if fma_peak else None, "flops_per_trajectory_step": fl},
            }
        if hi_acc is not None:
            hi_acc["frac_of_fp64_peak"] = hi_acc["achieved_tflops"] / fma_peak if fma_peak else None
            roof["order22_high_accuracy"] = hi_acc
        cpu = None
        if not args.no_cpu_baseline:
            cpu, _, _ = cpu_baseline(cfgd, order, args.cpu_seconds)
        total = int(round(steps_all / max(1, tot_steps) * B)) if world > 1 else B
        line = {
            "metric": metric_name(args.config), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_max / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {
                "workload": cfgd["name"] + ", final time {:g}".format(cfgd["horizon"]),
                "config_id": args.config,
                "trajectories_this_gpu": B,
                "trajectories_total": (args.total or min(CFG_TOTAL[args.config], CFG_MAX_PER_GPU[args.config] * world))
                if args.scaling == "strong" else B * world,
                "trajectories_in_BASELINE_config": CFG_TOTAL[args.config],
                "scaling_note": ("strong: the total is split over the GPUs by trajectory range; "
                                 "at N = 1 one GPU runs the largest shard it can hold ({} trajectories)".format(
                                     CFG_MAX_PER_GPU[args.config]))
                if args.scaling == "strong" else "weak: every GPU runs --traj-per-gpu trajectories",
                "horizon": cfgd["horizon"],
                "step": cfgd["step_desc"] + " (default W+K = 10 steps cover the horizon once)",
                "parallelism": "trajectory-range shards, no collective",
                "l2": "flushed between iterations (256 MiB memset)",
                "launch": li,
                "wall_s_timed_region": wall,
                "check": {cfgd["check_name"] + "_after_timed_steps": drift},
            },
            "roofline": roof,
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": gpu_launches,
            "clocks": clocks,
            "trajectory_steps_per_step": steps_all / args.steps,
        }
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
