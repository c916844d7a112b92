"""ctypes binding of libhy_cuda (include/hy_cuda.h).

This is the reference-side binding a maintainer would add in place of the
heyoka C++ calls made from expose_batch_integrators.cpp.
There is NO CPU fallback: if the shared library or a CUDA device is missing,
every entry point raises.
"""

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HY_CUDA_LIB") or os.path.join(_HERE, "..", "csrc", "libhy_cuda.so")

_lib = None


class HyCudaError(RuntimeError):
    pass


class dims_t(C.Structure):
    _fields_ = [
        (k, C.c_uint32)
        for k in (
            "n_state",
            "n_par",
            "order",
            "n_rows",
            "n_ops",
            "n_terms",
            "n_levels",
            "n_events",
            "n_tevents",
        )
    ]


class launch_info_t(C.Structure):
    _fields_ = [
        (k, C.c_uint32)
        for k in (
            "group",
            "traj_per_cta",
            "threads",
            "ctas",
            "smem_bytes",
            "ws_in_smem",
            "n_sm",
            "regs_per_thread",
            "kernel_variant",
        )
    ]


class event_rec_t(C.Structure):
    _fields_ = [
        ("lane", C.c_uint32),
        ("ev_idx", C.c_uint32),
        ("d_sgn", C.c_int32),
        ("step", C.c_uint32),
        ("t", C.c_double),
    ]


class tape_t(C.Structure):
    """hy_tape (include/hy_cuda.h)."""

    _fields_ = [("dims", C.c_void_p), ("ops", C.c_void_p), ("terms", C.c_void_p),
                ("level_start", C.c_void_p), ("ev_ref", C.c_void_p)]


class event_tape_t(C.Structure):
    """hy_event_tape (include/hy_cuda.h)."""

    _fields_ = [("n_ops", C.c_uint32), ("n_terms", C.c_uint32), ("n_rows", C.c_uint32),
                ("n_events", C.c_uint32), ("ops", C.c_void_p), ("terms", C.c_void_p),
                ("ev_ref", C.c_void_p), ("op_start", C.c_void_p)]


def _dims_of(dc, n_tevents, n_events=None):
    return dims_t(dc.n_state, dc.n_par, dc.order, dc.n_rows, len(dc.ops), len(dc.terms),
                  len(dc.level_start) - 1, dc.n_events if n_events is None else n_events, n_tevents)


class prop_args_t(C.Structure):
    """hy_prop_args (include/hy_cuda.h)."""

    _fields_ = [
        ("t", C.c_void_p),
        ("is_delta", C.c_int),
        ("max_steps", C.c_uint64),
        ("max_delta_t", C.c_void_p),
        ("write_tc", C.c_int),
        ("c_output", C.c_int),
        ("active", C.c_void_p),
        ("resume", C.c_int),
        ("launch_steps", C.c_uint64),
        ("pause_on_nt", C.c_int),
        ("grid", C.c_void_p),
        ("grid_k", C.c_size_t),
        ("grid_out", C.c_void_p),
    ]


OUTCOME_PAUSED = -4294967400

event_rec_dtype = np.dtype(
    [("lane", "<u4"), ("ev_idx", "<u4"), ("d_sgn", "<i4"), ("step", "<u4"), ("t", "<f8")]
)

# Every symbol include/hy_cuda.h declares (checked by tests/test_cabi_symbols.py).
SYMBOLS = [
    "hy_last_error",
    "hy_device_count",
    "hy_create",
    "hy_create2",
    "hy_destroy",
    "hy_clone",
    "hy_get_device",
    "hy_get_event_stats",
    "hy_sync",
    "hy_set_tc",
    "hy_set_last_h",
    "hy_propagate_ex",
    "hy_set_angle_reducer",
    "hy_cout_detach",
    "hy_cout_from_host",
    "hy_cout_free",
    "hy_host_alloc",
    "hy_host_free",
    "hy_set_stream",
    "hy_upload",
    "hy_download",
    "hy_upload_dev",
    "hy_state_dev",
    "hy_step",
    "hy_propagate",
    "hy_propagate_grid",
    "hy_last_timing",
    "hy_get_tc",
    "hy_dense_eval",
    "hy_cout_info",
    "hy_cout_get",
    "hy_cout_eval",
    "hy_cout_eval_dev",
    "hy_events_count",
    "hy_events_drain",
    "hy_get_cooldowns",
    "hy_set_cooldowns",
    "hy_reset_cooldowns",
    "hy_get_launch_info",
    "hy_tape_kernel_variant",
    "hy_jit_precompile",
    "hy_jit_precompile_events",
    "hy_measure_fma_peak",
]


def lib():
    """Load libhy_cuda.so (built in-tree by __graft_entry__.build())."""
    global _lib
    if _lib is None:
        path = os.path.abspath(LIB_PATH)
        if not os.path.exists(path):
            raise HyCudaError(
                "libhy_cuda.so not found at {}: build it with `python -c 'import "
                "__graft_entry__ as g; g.build()'` (there is no CPU fallback)".format(path)
            )
        _lib = C.CDLL(path)
        _lib.hy_last_error.restype = C.c_char_p
        for s in SYMBOLS:
            if s != "hy_last_error":
                getattr(_lib, s).restype = C.c_int
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().hy_last_error()
        raise HyCudaError(msg.decode() if msg else "libhy_cuda failure")


def ptr(a):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def jit_precompile(dc, fp_bits=64, batch=1 << 20, n_tevents=0):
    """hy_jit_precompile: generate + compile the run-time kernel of a tape without a device (warms
    the on-disk kernel cache).  Returns (from_cache, compile_seconds); from_cache is -1 when the
    tape is served by the interpreter."""
    dims = _dims_of(dc, n_tevents)
    full = tape_t(C.addressof(dims), _vp(dc.ops), _vp(dc.terms), _vp(dc.level_start), _vp(dc.ev_ref))
    fc, cs = C.c_int(0), C.c_double(0.0)
    check(lib().hy_jit_precompile(C.c_int(fp_bits), C.byref(full), C.c_uint32(batch), C.byref(fc), C.byref(cs)))
    return fc.value, cs.value


def jit_precompile_events(dc, dc_ode, evt, fp_bits=64, n_tevents=0):
    """hy_jit_precompile_events: the generated event functions of a system that a register-resident
    kernel serves (same tapes as hy_create2), compiled without a device.  Returns (from_cache,
    compile_seconds); from_cache is -1 when the system gets no such kernel."""
    dims = _dims_of(dc, n_tevents)
    d_ode = _dims_of(dc_ode, 0, 0)
    full = tape_t(C.addressof(dims), _vp(dc.ops), _vp(dc.terms), _vp(dc.level_start), _vp(dc.ev_ref))
    ode = tape_t(C.addressof(d_ode), _vp(dc_ode.ops), _vp(dc_ode.terms), _vp(dc_ode.level_start), None)
    et = event_tape_t(len(evt.ops), len(evt.terms), evt.n_rows, evt.n_events, _vp(evt.ops), _vp(evt.terms),
                      _vp(evt.ev_ref), _vp(evt.op_start))
    fc, cs = C.c_int(0), C.c_double(0.0)
    check(lib().hy_jit_precompile_events(C.c_int(fp_bits), C.byref(full), C.byref(ode), C.byref(et), C.byref(fc),
                                         C.byref(cs)))
    return fc.value, cs.value


def device_count():
    n = C.c_int(0)
    rc = lib().hy_device_count(C.byref(n))
    if rc != 0:
        return 0
    return n.value


class PinnedArray:
    """numpy array over cudaHostAlloc'ed memory."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(int(s) for s in shape)
        nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        self._p = C.c_void_p()
        check(lib().hy_host_alloc(C.byref(self._p), C.c_size_t(nbytes)))
        buf = (C.c_char * max(nbytes, 1)).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(
            self.shape
        )

    def __del__(self):
        try:
            if self._p and self._p.value:
                lib().hy_host_free(self._p)
                self._p = None
        except Exception:
            pass


class _PinnedBlock:
    """One pinned buffer of the output pool handed out as an ndarray (np.asarray(block)): the array's
    base keeps the block alive, the block goes back to the pool with the last view."""

    def __init__(self, pool, p, cap, shape, dtype):
        self._pool, self._p, self._cap = pool, p, cap
        self.__array_interface__ = {"shape": tuple(shape), "typestr": np.dtype(dtype).str, "data": (p, False),
                                    "version": 3}

    def __del__(self):
        try:
            self._pool._give(self._p, self._cap)
        except Exception:
            pass


class PinnedPool:
    """Recycled page-locked result buffers.  A large result (the [k, n, B] array of a continuous-output
    evaluation) written into fresh pageable memory costs a page fault per 4 KiB and a staged copy - several
    times the kernel that produced it; a pinned buffer that a previous, dropped result has released takes
    the DMA directly.  Arrays the caller keeps stay valid (their buffer is not reused while referenced)."""

    MIN_BYTES = 1 << 20

    def __init__(self, keep_bytes=4 << 30):
        import threading

        self._lock = threading.Lock()
        self._free = []  # (cap, ptr)
        self._kept = 0
        self._keep_bytes = keep_bytes

    def array(self, shape, dtype):
        nbytes = int(np.prod(shape, dtype=np.int64)) * np.dtype(dtype).itemsize
        if nbytes < self.MIN_BYTES:
            return np.zeros(shape, dtype=dtype)
        p = None
        with self._lock:
            best = None
            for i, (cap, _) in enumerate(self._free):
                if nbytes <= cap <= 2 * nbytes and (best is None or cap < self._free[best][0]):
                    best = i
            if best is not None:
                cap, p = self._free.pop(best)
                self._kept -= cap
        if p is None:
            cap = (nbytes + (1 << 21) - 1) & ~((1 << 21) - 1)
            q = C.c_void_p()
            if lib().hy_host_alloc(C.byref(q), C.c_size_t(cap)) != 0 or not q.value:
                return np.zeros(shape, dtype=dtype)  # (no page-locked memory left: an ordinary array)
            p = q.value
        return np.asarray(_PinnedBlock(self, p, cap, shape, dtype))

    def _give(self, p, cap):
        with self._lock:
            if self._kept + cap <= self._keep_bytes:
                self._free.append((cap, p))
                self._kept += cap
                return
        lib().hy_host_free(C.c_void_p(p))


OUT_POOL = PinnedPool()


def _vp(a):
    return None if a is None else a.ctypes.data


class Context:
    """Owner of one hy_ctx."""

    def __init__(self, dc, fp_bits, batch, tol, high_accuracy, device=0, n_tevents=0,
                 ev_dir=None, ev_cooldown=None, _handle=None, dc_ode=None, evt=None, compact_mode=False):
        import threading

        # hy_create's flag word: HY_CREATE_HIGH_ACCURACY | HY_CREATE_COMPACT (include/hy_cuda.h)
        high_accuracy = (1 if high_accuracy else 0) | (2 if compact_mode else 0)

        self._ctx = C.c_void_p()
        self._lock = threading.Lock()  # guards the context's recorder against a recycling __del__
        self.batch = batch
        self.fp_bits = fp_bits
        self.device = device
        self._dc = dc
        if _handle is not None:
            self._ctx = _handle
            return
        self.dims = dims_t(
            dc.n_state,
            dc.n_par,
            dc.order,
            dc.n_rows,
            len(dc.ops),
            len(dc.terms),
            len(dc.level_start) - 1,
            dc.n_events,
            n_tevents,
        )
        self._keep = (dc.ops, dc.terms, dc.level_start, dc.ev_ref)
        evd = None if ev_dir is None else np.ascontiguousarray(ev_dir, dtype=np.int32)
        evc = None if ev_cooldown is None else np.ascontiguousarray(ev_cooldown, dtype=np.float64)
        if dc_ode is not None and evt is not None and dc.n_events:
            # event-carrying system: hand over the ODE-only tape and the event tape as well, so that
            # a matched ODE runs on its register-resident kernel (hy_create2)
            self._keep += (dc_ode.ops, dc_ode.terms, dc_ode.level_start, evt.ops, evt.terms, evt.ev_ref,
                           evt.op_start)
            d_ode = _dims_of(dc_ode, 0, 0)
            full = tape_t(C.addressof(self.dims), _vp(dc.ops), _vp(dc.terms), _vp(dc.level_start), _vp(dc.ev_ref))
            ode = tape_t(C.addressof(d_ode), _vp(dc_ode.ops), _vp(dc_ode.terms), _vp(dc_ode.level_start), None)
            et = event_tape_t(len(evt.ops), len(evt.terms), evt.n_rows, evt.n_events, _vp(evt.ops),
                              _vp(evt.terms), _vp(evt.ev_ref), _vp(evt.op_start))
            check(lib().hy_create2(C.byref(self._ctx), C.c_int(device), C.c_int(fp_bits), C.byref(full),
                                   C.byref(ode), C.byref(et), ptr(evd), ptr(evc), C.c_double(tol),
                                   C.c_int(int(high_accuracy)), C.c_uint32(batch)))
            return
        check(
            lib().hy_create(
                C.byref(self._ctx),
                C.c_int(device),
                C.c_int(fp_bits),
                C.byref(self.dims),
                ptr(dc.ops),
                ptr(dc.terms),
                ptr(dc.level_start),
                ptr(dc.ev_ref),
                ptr(evd),
                ptr(evc),
                C.c_double(tol),
                C.c_int(int(high_accuracy)),
                C.c_uint32(batch),
            )
        )

    def clone(self, device=-1):
        """hy_clone: deep copy onto `device` (no re-scheduling of the tape)."""
        h = C.c_void_p()
        rc = lib().hy_clone(self._ctx, C.byref(h), C.c_int(int(device)))
        if rc != 0:
            if h.value:
                lib().hy_destroy(h)
            check(rc)
        dev = self.device if device < 0 else int(device)
        return Context(self._dc, self.fp_bits, self.batch, 0.0, False, device=dev, _handle=h)

    def close(self):
        if self._ctx is not None and self._ctx.value:
            with self._lock:
                h, self._ctx = self._ctx, None
            lib().hy_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        check(lib().hy_set_stream(self._ctx, C.c_void_p(cuda_stream)))

    def sync(self):
        check(lib().hy_sync(self._ctx))

    def upload(self, state=None, pars=None, t_hi=None, t_lo=None):
        check(lib().hy_upload(self._ctx, ptr(state), ptr(pars), ptr(t_hi), ptr(t_lo)))

    def download(self, state=None, t_hi=None, t_lo=None, last_h=None):
        check(lib().hy_download(self._ctx, ptr(state), ptr(t_hi), ptr(t_lo), ptr(last_h)))

    def set_tc(self, tc):
        check(lib().hy_set_tc(self._ctx, ptr(tc)))

    def set_last_h(self, last_h):
        check(lib().hy_set_last_h(self._ctx, ptr(last_h)))

    def set_angle_reducer(self, idx):
        arr = np.ascontiguousarray(idx, dtype=np.uint32)
        check(lib().hy_set_angle_reducer(self._ctx, ptr(arr) if arr.size else None, C.c_uint32(arr.size)))

    def step(self, max_delta_t, backward, write_tc, outcome, h):
        check(
            lib().hy_step(
                self._ctx, ptr(max_delta_t), C.c_int(int(backward)), C.c_int(int(write_tc)),
                ptr(outcome), ptr(h),
            )
        )

    def propagate(self, t, is_delta, max_steps, max_delta_t, write_tc, c_output, outcome,
                  min_h, max_h, n_steps):
        with self._lock:
            rc = lib().hy_propagate(
                self._ctx, ptr(t), C.c_int(int(is_delta)), C.c_uint64(int(max_steps)),
                ptr(max_delta_t), C.c_int(int(write_tc)), C.c_int(int(c_output)), ptr(outcome),
                ptr(min_h), ptr(max_h), ptr(n_steps),
            )
        check(rc)

    def propagate_ex(self, outcome, min_h, max_h, n_steps, t=None, is_delta=False, max_steps=0,
                     max_delta_t=None, write_tc=False, c_output=0, active=None, resume=False,
                     launch_steps=0, pause_on_nt=False, grid=None, grid_out=None):
        """hy_propagate_ex: propagate_until/for/grid with an active-lane mask, resumable
        launches and an appendable continuous output (include/hy_cuda.h)."""
        a = prop_args_t(
            _vp(t), int(is_delta), int(max_steps), _vp(max_delta_t), int(write_tc), int(c_output),
            _vp(active), int(resume), int(launch_steps), int(pause_on_nt), _vp(grid),
            0 if grid is None else int(grid.shape[0]), _vp(grid_out),
        )
        with self._lock:
            rc = lib().hy_propagate_ex(self._ctx, C.byref(a), ptr(outcome), ptr(min_h), ptr(max_h),
                                       ptr(n_steps))
        check(rc)

    def propagate_grid(self, grid, k, max_steps, max_delta_t, out, outcome, min_h, max_h, n_steps):
        check(
            lib().hy_propagate_grid(
                self._ctx, ptr(grid), C.c_size_t(int(k)), C.c_uint64(int(max_steps)),
                ptr(max_delta_t), ptr(out), ptr(outcome), ptr(min_h), ptr(max_h), ptr(n_steps),
            )
        )

    def last_timing(self):
        ms = C.c_double(0)
        n = C.c_uint64(0)
        check(lib().hy_last_timing(self._ctx, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def get_tc(self, tc):
        check(lib().hy_get_tc(self._ctx, ptr(tc)))

    def dense_eval(self, t, rel_time, out):
        check(lib().hy_dense_eval(self._ctx, ptr(t), C.c_int(int(rel_time)), ptr(out)))

    def cout_detach(self):
        """The continuous output recorded by the last propagate(c_output) call, as an object of
        its own (None if nothing was recorded)."""
        h = C.c_void_p()
        with self._lock:
            rc = lib().hy_cout_detach(self._ctx, C.byref(h))
        check(rc)
        if not h.value:
            return None
        return CoutRecord(h, self)

    def events_drain(self):
        n = C.c_uint64(0)
        check(lib().hy_events_count(self._ctx, C.byref(n)))
        recs = np.zeros(n.value, dtype=event_rec_dtype)
        if n.value:
            m = C.c_uint64(0)
            check(lib().hy_events_drain(self._ctx, ptr(recs), C.c_uint64(n.value), C.byref(m)))
            recs = recs[: m.value]
        return recs

    def get_cooldowns(self, elapsed, total):
        check(lib().hy_get_cooldowns(self._ctx, ptr(elapsed), ptr(total)))

    def set_cooldowns(self, elapsed, total):
        check(lib().hy_set_cooldowns(self._ctx, ptr(elapsed), ptr(total)))

    def reset_cooldowns(self, lane=-1):
        check(lib().hy_reset_cooldowns(self._ctx, C.c_int64(int(lane))))

    def event_stats(self):
        a, b = C.c_uint64(0), C.c_uint64(0)
        check(lib().hy_get_event_stats(self._ctx, C.byref(a), C.byref(b)))
        return a.value, b.value

    def launch_info(self):
        li = launch_info_t()
        check(lib().hy_get_launch_info(self._ctx, C.byref(li)))
        return {k: getattr(li, k) for k, _ in launch_info_t._fields_}


class CoutRecord:
    """Owner of one hy_cout (a recorded continuous output in device memory)."""

    def __init__(self, handle, ctx):
        import weakref

        self._h = handle
        self._ctx_ref = weakref.ref(ctx) if ctx is not None else (lambda: None)

    def info(self, n_steps):
        mx = C.c_uint64(0)
        check(lib().hy_cout_info(self._h, ptr(n_steps), C.byref(mx)))
        return mx.value

    def get(self, tcs, thi, tlo, S):
        check(lib().hy_cout_get(self._h, ptr(tcs), ptr(thi), ptr(tlo), C.c_uint64(int(S))))

    def eval(self, t, k, out):
        check(lib().hy_cout_eval(self._h, ptr(t), C.c_size_t(int(k)), ptr(out)))

    def eval_dev(self, d_t, k, d_out):
        """hy_cout_eval_dev: query times and output are device pointers (ints)."""
        check(lib().hy_cout_eval_dev(self._h, C.c_void_p(int(d_t)), C.c_size_t(int(k)), C.c_void_p(int(d_out))))

    def close(self):
        if self._h is not None and self._h.value:
            ctx = self._ctx_ref()
            # give the pool back to the context it came from, if that is alive and idle
            if (ctx is not None and ctx._ctx is not None and ctx._ctx.value
                    and ctx._lock.acquire(False)):
                try:
                    lib().hy_cout_free(self._h, ctx._ctx)
                finally:
                    ctx._lock.release()
            else:
                lib().hy_cout_free(self._h, None)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def cout_from_host(fp, n, order, B, n_steps, tcs, thi, tlo, S, device=0):
    """hy_cout_from_host (include/hy_cuda_cout.h): a record rebuilt from the arrays hy_cout_get exports -
    how a copied / unpickled continuous_output_batch gets its device record back."""
    h = C.c_void_p()
    ns = np.ascontiguousarray(n_steps, dtype=np.uint64)
    tcs = np.ascontiguousarray(tcs, dtype=fp)
    thi = np.ascontiguousarray(thi, dtype=fp)
    tlo = np.ascontiguousarray(tlo, dtype=fp)
    check(lib().hy_cout_from_host(C.byref(h), C.c_int(int(device)), C.c_int(64 if np.dtype(fp) == np.float64 else 32),
                                  C.c_uint32(int(n)), C.c_uint32(int(order)), C.c_uint32(int(B)), ptr(ns), ptr(tcs),
                                  ptr(thi), ptr(tlo), C.c_uint64(int(S))))
    return CoutRecord(h, None)


def tape_kernel_variant(dc):
    """Kernel hy_create would pick for the decomposition `dc` (no device needed):
    0 = tape interpreter, N = register-resident N-body kernel for N bodies."""
    dims = dims_t(dc.n_state, dc.n_par, dc.order, dc.n_rows, len(dc.ops), len(dc.terms),
                  len(dc.level_start) - 1, dc.n_events, 0)
    v = C.c_uint32(0)
    check(lib().hy_tape_kernel_variant(C.byref(dims), ptr(dc.ops), ptr(dc.terms), C.byref(v)))
    return v.value


def measure_fma_peak(device=0, fp_bits=64):
    tf = C.c_double(0)
    check(lib().hy_measure_fma_peak(C.c_int(device), C.c_int(fp_bits), C.byref(tf)))
    return tf.value
