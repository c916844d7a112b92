"""Device-context management of the batch integrator (host side).

 * ``host_array``: the numpy mirrors of state/pars/time live in page-locked memory when they
   are large enough for the DMA rate to matter, in plain numpy memory otherwise (an ensemble
   makes thousands of small integrator copies: page-locking each would cost more than it saves).
 * ``ContextPool``: contexts released by finished ensemble iterations are reused by the next
   ones (same tape, precision, batch size, device) instead of being created and destroyed
   per iteration (reference: one deepcopy(ta) per iteration, _ensemble_impl.py:47).
 * ``MultiContext``: ONE integrator whose lanes are split by contiguous trajectory range over
   several GPUs (shards only, no inter-GPU communication, final host
   gather).  It exposes the interface of ``_cabi.Context``; every call fans out to one host
   thread per device (ctypes releases the GIL) and gathers into the caller's arrays.
"""

import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _cabi
from .shard import shard_bounds

PIN_THRESHOLD = 256 * 1024  # bytes


class host_array:
    """numpy array, page-locked if it is large."""

    def __init__(self, shape, dtype, pinned=None):
        shape = tuple(int(s) for s in shape)
        nbytes = int(np.prod(shape, dtype=np.int64)) * np.dtype(dtype).itemsize
        if pinned is None:
            pinned = nbytes >= PIN_THRESHOLD
        self._pin = _cabi.PinnedArray(shape, dtype) if pinned else None
        self.array = self._pin.array if pinned else np.zeros(shape, dtype=dtype)


class ContextPool:
    """Free-list of idle contexts keyed by (tape identity, precision, batch, device, ...)."""

    def __init__(self):
        self._lock = threading.Lock()
        self._free = {}

    def acquire(self, key):
        with self._lock:
            lst = self._free.get(key)
            if lst:
                return lst.pop()
        return None

    def release(self, key, ctx):
        with self._lock:
            self._free.setdefault(key, []).append(ctx)

    def clear(self):
        with self._lock:
            free, self._free = self._free, {}
        for lst in free.values():
            for c in lst:
                c.close()


POOL = ContextPool()


class _MultiRecord:
    """Continuous output of a lane-sharded integrator: one record per shard."""

    def __init__(self, parts, bounds, n, B, fp):
        self.parts, self.bounds, self.n, self.B, self.fp = parts, bounds, n, B, fp

    def info(self, n_steps):
        mx = 0
        for r, (lo, hi) in zip(self.parts, self.bounds):
            if r is None:
                n_steps[lo:hi] = 0
                continue
            ns = np.zeros(hi - lo, dtype=np.uint64)
            mx = max(mx, r.info(ns))
            n_steps[lo:hi] = ns
        return mx

    def get(self, tcs, thi, tlo, S):
        for r, (lo, hi) in zip(self.parts, self.bounds):
            b = hi - lo
            t = np.full((S,) + tcs.shape[1:3] + (b,), np.nan, dtype=self.fp)
            h = np.full((S + 1, b), np.nan, dtype=self.fp)
            l = np.full((S + 1, b), np.nan, dtype=self.fp)
            if r is not None:
                r.get(t, h, l, S)
            tcs[..., lo:hi] = t
            thi[:, lo:hi] = h
            tlo[:, lo:hi] = l

    def eval(self, t, k, out):
        for r, (lo, hi) in zip(self.parts, self.bounds):
            b = hi - lo
            o = np.full((k, self.n, b), np.nan, dtype=self.fp)
            if r is not None:
                r.eval(np.ascontiguousarray(t[:, lo:hi]), k, o)
            out[:, :, lo:hi] = o

    def close(self):
        for r in self.parts:
            if r is not None:
                r.close()


class MultiContext:
    """Lane-sharded context over several devices with the interface of ``_cabi.Context``."""

    def __init__(self, dc, fp_bits, batch, tol, high_accuracy, devices, n_tevents=0, ev_dir=None,
                 ev_cooldown=None, dc_ode=None, evt=None, compact_mode=False):
        self.devices = list(devices)
        G = len(self.devices)
        self.batch = batch
        self.fp_bits = fp_bits
        self.fp = np.float64 if fp_bits == 64 else np.float32
        self.n = dc.n_state
        self.m = dc.n_par
        self.order = dc.order
        self.n_tevents = n_tevents
        self.bounds = [shard_bounds(batch, g, G) for g in range(G)]
        self.bounds = [b for b in self.bounds if b[1] > b[0]]
        self.devices = self.devices[: len(self.bounds)]
        self._pool = ThreadPoolExecutor(max_workers=len(self.bounds))
        first = _cabi.Context(dc, fp_bits, self.bounds[0][1] - self.bounds[0][0], tol, high_accuracy,
                              device=self.devices[0], n_tevents=n_tevents, ev_dir=ev_dir,
                              ev_cooldown=ev_cooldown, dc_ode=dc_ode, evt=evt, compact_mode=compact_mode)
        self.parts = [first]
        for (lo, hi), dev in zip(self.bounds[1:], self.devices[1:]):
            self.parts.append(_cabi.Context(dc, fp_bits, hi - lo, tol, high_accuracy, device=dev,
                                            n_tevents=n_tevents, ev_dir=ev_dir,
                                            ev_cooldown=ev_cooldown, dc_ode=dc_ode, evt=evt,
                                            compact_mode=compact_mode))
        # per-shard contiguous staging buffers (page-locked): the caller's arrays are [rows, B]
        # with the lane index fastest, so a shard is a strided slice of them
        self._st = [dict() for _ in self.parts]
        self.device = self.devices[0]

    # ---- helpers ----
    def _buf(self, g, name, shape, dtype):
        d = self._st[g]
        b = d.get(name)
        if b is None or b.array.shape != tuple(shape) or b.array.dtype != np.dtype(dtype):
            b = host_array(shape, dtype)
            d[name] = b
        return b.array

    def _fan(self, fn):
        futs = [self._pool.submit(fn, g, c, lo, hi) for g, (c, (lo, hi)) in
                enumerate(zip(self.parts, self.bounds))]
        return [f.result() for f in futs]

    def _in(self, g, name, a, lo, hi):
        if a is None:
            return None
        b = self._buf(g, name, a.shape[:-1] + (hi - lo,), a.dtype)
        b[...] = a[..., lo:hi]
        return b

    def close(self):
        for c in self.parts:
            c.close()
        self._pool.shutdown(wait=False)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        raise _cabi.HyCudaError("set_stream() is not available on a multi-device integrator")

    def sync(self):
        self._fan(lambda g, c, lo, hi: c.sync())

    # ---- state transfer ----
    def upload(self, state=None, pars=None, t_hi=None, t_lo=None):
        def f(g, c, lo, hi):
            c.upload(self._in(g, "state", state, lo, hi), self._in(g, "pars", pars, lo, hi),
                     self._in(g, "thi", t_hi, lo, hi), self._in(g, "tlo", t_lo, lo, hi))
            c.sync()  # the staging buffers are reused by the next call

        self._fan(f)

    def download(self, state=None, t_hi=None, t_lo=None, last_h=None):
        def f(g, c, lo, hi):
            b = hi - lo
            s = None if state is None else self._buf(g, "state", (state.shape[0], b), state.dtype)
            h = None if t_hi is None else self._buf(g, "thi", (b,), t_hi.dtype)
            l = None if t_lo is None else self._buf(g, "tlo", (b,), t_lo.dtype)
            lh = None if last_h is None else self._buf(g, "lasth", (b,), last_h.dtype)
            c.download(s, h, l, lh)
            for dst, src in ((state, s), (t_hi, h), (t_lo, l), (last_h, lh)):
                if dst is not None:
                    dst[..., lo:hi] = src

        self._fan(f)

    def set_tc(self, tc):
        self._fan(lambda g, c, lo, hi: c.set_tc(np.ascontiguousarray(tc[..., lo:hi])))

    def set_last_h(self, last_h):
        self._fan(lambda g, c, lo, hi: c.set_last_h(np.ascontiguousarray(last_h[lo:hi])))

    def set_angle_reducer(self, idx):
        self._fan(lambda g, c, lo, hi: c.set_angle_reducer(idx))

    # ---- stepping ----
    def step(self, max_delta_t, backward, write_tc, outcome, h):
        def f(g, c, lo, hi):
            b = hi - lo
            oc = np.zeros(b, dtype=np.int64)
            hh = np.zeros(b, dtype=h.dtype)
            c.step(self._in(g, "mdt", max_delta_t, lo, hi), backward, write_tc, oc, hh)
            outcome[lo:hi] = oc
            h[lo:hi] = hh

        self._fan(f)

    def propagate_ex(self, outcome, min_h, max_h, n_steps, t=None, is_delta=False, max_steps=0,
                     max_delta_t=None, write_tc=False, c_output=0, active=None, resume=False,
                     launch_steps=0, pause_on_nt=False, grid=None, grid_out=None):
        def f(g, c, lo, hi):
            b = hi - lo
            oc = np.zeros(b, dtype=np.int64)
            mn = np.zeros(b, dtype=self.fp)
            mx = np.zeros(b, dtype=self.fp)
            ns = np.zeros(b, dtype=np.uint64)
            go = None
            if grid is not None:
                go = self._buf(g, "gout", (grid.shape[0], self.n, b), self.fp)
            c.propagate_ex(oc, mn, mx, ns, t=self._in(g, "t", t, lo, hi), is_delta=is_delta,
                           max_steps=max_steps, max_delta_t=self._in(g, "mdt", max_delta_t, lo, hi),
                           write_tc=write_tc, c_output=c_output,
                           active=self._in(g, "active", active, lo, hi), resume=resume,
                           launch_steps=launch_steps, pause_on_nt=pause_on_nt,
                           grid=self._in(g, "grid", grid, lo, hi), grid_out=go)
            if outcome is not None:
                outcome[lo:hi] = oc
            if min_h is not None:
                min_h[lo:hi] = mn
            if max_h is not None:
                max_h[lo:hi] = mx
            if n_steps is not None:
                n_steps[lo:hi] = ns
            if go is not None:
                grid_out[:, :, lo:hi] = go

        self._fan(f)

    def propagate(self, t, is_delta, max_steps, max_delta_t, write_tc, c_output, outcome,
                  min_h, max_h, n_steps):
        self.propagate_ex(outcome, min_h, max_h, n_steps, t=t, is_delta=is_delta,
                          max_steps=max_steps, max_delta_t=max_delta_t, write_tc=write_tc,
                          c_output=1 if c_output else 0)

    def propagate_grid(self, grid, k, max_steps, max_delta_t, out, outcome, min_h, max_h, n_steps):
        self.propagate_ex(outcome, min_h, max_h, n_steps, max_steps=max_steps,
                          max_delta_t=max_delta_t, grid=grid, grid_out=out)

    def last_timing(self):
        r = [c.last_timing() for c in self.parts]
        return max(x[0] for x in r), sum(x[1] for x in r)

    def get_tc(self, tc):
        def f(g, c, lo, hi):
            b = self._buf(g, "tc", tc.shape[:-1] + (hi - lo,), tc.dtype)
            c.get_tc(b)
            tc[..., lo:hi] = b

        self._fan(f)

    def dense_eval(self, t, rel_time, out):
        def f(g, c, lo, hi):
            o = self._buf(g, "dout", (out.shape[0], hi - lo), out.dtype)
            c.dense_eval(self._in(g, "t", t, lo, hi), rel_time, o)
            out[:, lo:hi] = o

        self._fan(f)

    def cout_detach(self):
        parts = [c.cout_detach() for c in self.parts]
        if all(p is None for p in parts):
            return None
        return _MultiRecord(parts, self.bounds, self.n, self.batch, self.fp)

    # ---- events ----
    def events_drain(self):
        out = []
        for c, (lo, hi) in zip(self.parts, self.bounds):
            r = c.events_drain()
            if len(r):
                r = r.copy()
                r["lane"] += lo
                out.append(r)
        if not out:
            return np.zeros(0, dtype=_cabi.event_rec_dtype)
        return np.concatenate(out)

    def get_cooldowns(self, elapsed, total):
        for c, (lo, hi) in zip(self.parts, self.bounds):
            e = np.zeros((hi - lo,) + elapsed.shape[1:], dtype=elapsed.dtype)
            t = np.zeros_like(e)
            c.get_cooldowns(e, t)
            elapsed[lo:hi] = e
            total[lo:hi] = t

    def set_cooldowns(self, elapsed, total):
        for c, (lo, hi) in zip(self.parts, self.bounds):
            c.set_cooldowns(np.ascontiguousarray(elapsed[lo:hi]), np.ascontiguousarray(total[lo:hi]))

    def reset_cooldowns(self, lane=-1):
        if lane < 0:
            for c in self.parts:
                c.reset_cooldowns(-1)
            return
        for c, (lo, hi) in zip(self.parts, self.bounds):
            if lo <= lane < hi:
                c.reset_cooldowns(lane - lo)

    def launch_info(self):
        li = dict(self.parts[0].launch_info())
        li["devices"] = list(self.devices)
        li["shards"] = [hi - lo for lo, hi in self.bounds]
        li["ctas"] = sum(c.launch_info()["ctas"] for c in self.parts)
        return li
