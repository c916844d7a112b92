"""ctypes front end of the C oracle (oracle/hy_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs, never by the product.
"""

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libhy_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("hy_oracle.c", "hy_oracle_impl.h", "hy_baseline_simd.c")]
    srcs.append(os.path.join(_HERE, "..", "include", "hy_cuda.h"))
    stale = not os.path.exists(so) or any(
        os.path.getmtime(s) > os.path.getmtime(so) for s in srcs if os.path.exists(s)
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "all"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = build()  # no-op unless the sources are newer than the library
        _LIB = C.CDLL(so)
    return _LIB


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


def make_dims(dc, n_tevents=0):
    return dims_t(
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


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class COracle:
    """Batch integrator state on the host driven by the C oracle."""

    def __init__(self, dc, state, time=None, pars=None, tol=0.0, fp_type=np.float64,
                 high_accuracy=False, nthreads=0):
        self.dc = dc
        self.T = np.dtype(fp_type)
        self.suffix = "_f64" if self.T == np.float64 else "_f32"
        self.n = dc.n_state
        self.state = np.ascontiguousarray(np.array(state, dtype=self.T).reshape(self.n, -1))
        self.B = self.state.shape[1]
        self.pars = (
            np.zeros((dc.n_par, self.B), dtype=self.T)
            if pars is None
            else np.ascontiguousarray(np.array(pars, dtype=self.T).reshape(dc.n_par, self.B))
        )
        self.t_hi = (
            np.zeros(self.B, dtype=self.T) if time is None else np.array(time, dtype=self.T).copy()
        )
        self.t_lo = np.zeros(self.B, dtype=self.T)
        self.tol = float(np.finfo(self.T).eps) if tol == 0 else float(tol)
        self.high_accuracy = int(high_accuracy)
        self.nthreads = nthreads
        self.dims = make_dims(dc)
        self.last_h = np.zeros(self.B, dtype=self.T)
        self.tc = np.zeros((self.n, dc.order + 1, self.B), dtype=self.T)
        self.fn = getattr(lib(), "ora_propagate" + self.suffix)
        self.fn.restype = C.c_int

    def _call(self, t, is_delta, max_steps, max_delta_t, single, backward, h_log_cap=0):
        B, T = self.B, self.T
        oc = np.zeros(B, dtype=np.int64)
        mn = np.zeros(B, dtype=T)
        mx = np.zeros(B, dtype=T)
        ns = np.zeros(B, dtype=np.uint64)
        tt = None if t is None else np.ascontiguousarray(np.broadcast_to(np.array(t, dtype=T), (B,)))
        md = (
            None
            if max_delta_t is None
            else np.ascontiguousarray(np.broadcast_to(np.array(max_delta_t, dtype=T), (B,)))
        )
        hl = np.full((B, h_log_cap), np.nan, dtype=T) if h_log_cap else None
        dc = self.dc
        rc = self.fn(
            C.byref(self.dims), _p(dc.ops), _p(dc.terms), _p(dc.ev_ref), C.c_double(self.tol),
            C.c_int(self.high_accuracy), C.c_uint32(B), _p(self.state), _p(self.pars),
            _p(self.t_hi), _p(self.t_lo), _p(tt), C.c_int(is_delta), C.c_uint64(max_steps),
            _p(md), C.c_int(single), C.c_int(backward), _p(oc), _p(mn), _p(mx), _p(ns),
            _p(self.last_h), _p(self.tc), _p(hl), C.c_uint64(h_log_cap), C.c_int(self.nthreads),
        )
        if rc != 0:
            raise RuntimeError("oracle failure {}".format(rc))
        return oc, mn, mx, ns, hl

    def step(self, max_delta_t=None, backward=False):
        oc, _, _, _, hl = self._call(None, 0, 0, max_delta_t, 1, int(backward), 1)
        return oc, hl[:, 0].copy()

    def propagate_until(self, t, max_steps=0, max_delta_t=None, h_log_cap=0):
        return self._call(t, 0, max_steps, max_delta_t, 0, 0, h_log_cap)

    def propagate_for(self, dt, max_steps=0, max_delta_t=None, h_log_cap=0):
        return self._call(dt, 1, max_steps, max_delta_t, 0, 0, h_log_cap)


_SIMD = None
SIMD_FLAGS = "portable (-O3 -mavx2 -mfma)"


def _simd_native():
    """The CPU baseline compiled for the host it runs on (-O3 -march=native: AVX-512 where the
    host has it), into a temporary directory; None if that fails (the portable in-tree build is
    used then).  The in-tree .so must run on any x86-64 box, so it cannot be -march=native."""
    global SIMD_FLAGS
    import tempfile

    if os.environ.get("HY_BASELINE_PORTABLE"):
        return None
    try:
        out = os.path.join(tempfile.mkdtemp(prefix="hy_baseline_"), "libhy_baseline_simd_native.so")
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.check_call([cc, "-O3", "-march=native", "-fPIC", "-shared", "-fopenmp", "-fno-fast-math",
                               "-o", out, os.path.join(_HERE, "hy_baseline_simd.c"), "-lm"],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        SIMD_FLAGS = "native (-O3 -march=native)"
        return out
    except Exception:
        return None


def simd_propagate_until(dc, state, t, pars=None, nthreads=0):
    """SIMD-batched multithreaded CPU baseline (oracle/hy_baseline_simd.c): FP64
    propagate_until from time 0.  Returns (final state, n_steps)."""
    global _SIMD
    if _SIMD is None:
        build()
        _SIMD = C.CDLL(_simd_native() or os.path.join(_HERE, "libhy_baseline_simd.so"))
        _SIMD.ora_simd_propagate_until.restype = C.c_int
    st = np.ascontiguousarray(np.array(state, dtype=np.float64).reshape(dc.n_state, -1))
    B = st.shape[1]
    pr = (np.zeros((max(dc.n_par, 1), B)) if pars is None
          else np.ascontiguousarray(np.array(pars, dtype=np.float64).reshape(dc.n_par, B)))
    t0 = np.zeros(B)
    tf = np.ascontiguousarray(np.broadcast_to(np.array(t, dtype=np.float64), (B,)))
    ns = np.zeros(B, dtype=np.uint64)
    dims = make_dims(dc)
    rc = _SIMD.ora_simd_propagate_until(C.byref(dims), _p(dc.ops), _p(dc.terms), C.c_uint32(B), _p(st), _p(pr),
                                        _p(t0), _p(tf), _p(ns), C.c_int(nthreads))
    if rc != 0:
        raise RuntimeError("simd baseline failure {}".format(rc))
    return st, ns
