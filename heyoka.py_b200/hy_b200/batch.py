"""``taylor_adaptive_batch``: host-side mirror of the reference's pybind11
class (/root/reference/heyoka/expose_batch_integrators.cpp:214-669), driving
libhy_cuda through the C ABI instead of heyoka's JIT-compiled stepper.

Same constructor keywords, properties and methods; the numpy views are host
mirrors in pinned memory: they are pushed to the device at the start of every
``step``/``propagate_*`` call and pulled back at its end, so in-place edits
such as ``ta.state[:] = ic`` (reference: _test_ensemble.py:52) behave as they
do in the reference.
"""

import copy as _copy
import math

import numpy as np

from . import _cabi
from . import _devctx
from . import decompose as _dec
from . import _expression as _E
from .enums import taylor_outcome, code_model, _outcome_from_int
from .events import nt_event_batch_impl, t_event_batch_impl
from .var_ode_sys import var_ode_sys as _var_ode_sys

_LLVM_KW = ("opt_level", "force_avx512", "slp_vectorize", "fast_math", "code_model", "parjit")


class _llvm_state_stub:
    """Inert stand-in for ``ta.llvm_state`` (the reference exposes the JIT
    state, expose_batch_integrators.cpp:670; there is no LLVM here)."""

    def __init__(self, kw, ta=None):
        self._ta = ta  # (the reference's llvm_state property keeps the integrator alive)
        self.opt_level = kw.get("opt_level", 3)
        self.force_avx512 = kw.get("force_avx512", False)
        self.slp_vectorize = kw.get("slp_vectorize", False)
        self.fast_math = kw.get("fast_math", False)
        self.code_model = kw.get("code_model", code_model.small)
        self.parjit = kw.get("parjit", False)

    def get_ir(self):
        return ""

    def __repr__(self):
        return "<llvm_state stub: libhy_cuda does not JIT>"


class _view_owner:
    """Base object of the numpy views of an integrator's buffers: holds the integrator, so that a view
    outlives `del ta` - the buffers are page-locked memory the integrator frees when it goes.  (The
    reference's views hold a reference to the integrator in the same way,
    expose_batch_integrators.cpp:394-518: `sys.getrefcount(ta)` grows by one per live view.)"""

    def __init__(self, ta, a, writeable):
        self._ta = ta
        ai = dict(a.__array_interface__)
        ai["data"] = (ai["data"][0], not writeable)
        self.__array_interface__ = ai


def _check_scalar_type(x, fp_t, what):
    """`.noconvert()` semantics of the reference (custom_casters.hpp:31-43):
    float64 integrators take Python floats, float32 ones need numpy.float32."""
    if fp_t == np.float32:
        ok = isinstance(x, np.float32)
    else:
        ok = isinstance(x, (float, np.float64)) and not isinstance(x, np.float32)
    if not ok:
        raise TypeError(
            "{}: incompatible argument type {} for a {} integrator".format(
                what, type(x).__name__, "single-precision" if fp_t == np.float32 else "double-precision"
            )
        )


class taylor_adaptive_batch_impl:
    """Implementation shared by ``taylor_adaptive_batch_dbl``/``_flt``."""

    _fp = np.float64

    def __init__(self, sys, state, time=None, pars=None, tol=0.0, high_accuracy=False,
                 compact_mode=False, t_events=None, nt_events=None, parallel_mode=False, **kw):
        for k in kw:
            if k not in _LLVM_KW and k != "device":
                raise TypeError("__init__(): unexpected keyword argument '{}'".format(k))
        fp = self._fp
        if isinstance(tol, (int, float)) and not isinstance(tol, bool) and tol == 0:
            tol = fp(0)  # the default: eps of the working precision
        _check_scalar_type(tol, fp, "tol")
        t_events = list(t_events) if t_events is not None else []
        nt_events = list(nt_events) if nt_events is not None else []
        for e in t_events:
            if not isinstance(e, t_event_batch_impl) or e._fp != fp:
                raise TypeError("t_events must be a list of t_event_batch objects of matching type")
        for e in nt_events:
            if not isinstance(e, nt_event_batch_impl) or e._fp != fp:
                raise TypeError("nt_events must be a list of nt_event_batch objects of matching type")

        state_ = np.asarray(state)
        if state_.ndim != 2:
            raise ValueError(
                "Invalid state vector passed to the constructor of a batch integrator: "
                "the expected number of dimensions is 2, but the input array has a dimension of {}".format(
                    state_.ndim
                )
            )
        if state_.dtype != fp:
            state_ = state_.astype(fp, casting="safe")
        B = int(state_.shape[1])
        if B == 0:
            raise ValueError("The batch size in a batch integrator cannot be zero")

        self._vsys = None
        auto_var_ic = False
        n_orig = 0
        if isinstance(sys, _var_ode_sys):
            self._vsys = sys
            sys_list = sys.sys
            n_orig = sys.n_orig_sv
            if state_.shape[0] == n_orig:
                # Auto-fill the variational initial conditions (identity).
                full = np.zeros((len(sys_list), B), dtype=fp)
                full[:n_orig] = state_
                full[n_orig:] = sys._initial_var_state(fp)[:, None]
                state_ = full
                auto_var_ic = True
        else:
            sys_list = list(sys)
        self._sys = [(l, _E._wrap(r)) for l, r in sys_list]
        n = len(self._sys)
        if state_.shape[0] == 0:
            # an empty initial state: zeros (the C++ constructor value-initialises a missing state;
            # /root/reference/heyoka/_test_batch_integrator.py:530-539)
            state_ = np.zeros((n, B), dtype=fp)
        if state_.shape[0] != n:
            raise ValueError(
                "Inconsistent sizes detected in the initialization of an adaptive Taylor "
                "integrator: the state vector has a dimension of {} and a batch size of {}, "
                "while the number of equations is {}".format(state_.shape[0] * B, B, n)
            )

        eps = float(np.finfo(fp).eps)
        tol = float(tol)
        if not math.isfinite(tol) or tol < 0:
            raise ValueError(
                "The tolerance in an adaptive Taylor integrator must be finite and positive, "
                "but it is {} instead".format(tol)
            )
        self._tol = eps if tol == 0.0 else tol
        self._order = _dec.taylor_order(self._tol)
        self._high_accuracy = bool(high_accuracy)
        self._compact_mode = bool(compact_mode)
        self._parallel_mode = bool(parallel_mode)
        self._llvm_kw = {k: v for k, v in kw.items() if k in _LLVM_KW}
        # device: a CUDA device index, or "all" / a list of indices to split the lanes of THIS
        # integrator over several GPUs by contiguous trajectory range
        dev = kw.get("device", 0)
        if isinstance(dev, str):
            if dev != "all":
                raise ValueError("device must be an index, a list of indices or 'all'")
        elif isinstance(dev, (list, tuple)):
            dev = [int(x) for x in dev]
        else:
            dev = int(dev)
        self._device = dev
        self._t_events = t_events
        self._nt_events = nt_events
        ev_exprs = [e.expression for e in t_events] + [e.expression for e in nt_events]
        self._dc = _dec.decompose(self._sys, self._order, events=ev_exprs)
        # event-carrying systems: the ODE alone and the events alone, for the register-resident kernels
        self._dc_ode = self._evt = None
        if ev_exprs:
            self._evt = _dec.decompose_event_tape(ev_exprs, [l.name for l, _ in self._sys], self._order)
            if self._evt is not None:
                self._dc_ode = _dec.decompose(self._sys, self._order)
        m = self._dc.n_par

        if pars is not None:
            pars_ = np.asarray(pars)
            if pars_.ndim != 2 or pars_.shape[1] != B:
                raise ValueError(
                    "Invalid parameter vector passed to the constructor of a batch integrator: "
                    "the expected array shape is (n, {}), but the input array has either the wrong "
                    "number of dimensions or the wrong shape".format(B)
                )
            if pars_.dtype != fp:
                pars_ = pars_.astype(fp, casting="safe")
            if pars_.shape[0] < m:
                full = np.zeros((m, B), dtype=fp)
                full[: pars_.shape[0]] = pars_
                pars_ = full
            elif pars_.shape[0] > m:
                # Extra parameters are legal: keep them as inert rows.
                self._dc.n_par = m = pars_.shape[0]
        else:
            pars_ = np.zeros((m, B), dtype=fp)
        if time is not None:
            time_ = np.asarray(time)
            if time_.ndim != 1 or time_.shape[0] != B:
                raise ValueError(
                    "Invalid time vector passed to the constructor of a batch integrator: "
                    "the expected array shape is ({}), but the input array has either the wrong "
                    "number of dimensions or the wrong shape".format(B)
                )
            time_ = time_.astype(fp, casting="safe")
        else:
            time_ = np.zeros(B, dtype=fp)

        if self._vsys is not None and getattr(self._vsys, "_ic_sym", None) is not None and auto_var_ic:
            # the initial time is a variational argument: its initial conditions depend on the initial state,
            # the parameters and the time (var_ode_sys.py)
            state_[n_orig:] = self._vsys._initial_var_state_at(state_[:n_orig], pars_, time_, fp)
        self._B = B
        self._n = n
        self._alloc(state_, pars_, time_, np.zeros(B, dtype=fp))

    # ------------------------------------------------------------------
    def _alloc(self, state, pars, t_hi, t_lo):
        """Create the host mirrors.  The device context is created on first use (or taken from
        the pool of idle contexts / cloned from the integrator this one was copied from)."""
        fp, B, n, m = self._fp, self._B, self._n, self._dc.n_par
        self._ctx_obj = None
        self._ctx_src = None      # context to clone from (hy_clone: no re-scheduling)
        self._snap = None         # device-only data to restore into a new context
        self._reducer_idx = ()    # device-side angle reduction currently configured
        self._p_state = _devctx.host_array((n, B), fp)
        self._p_pars = _devctx.host_array((m, B), fp)
        self._p_thi = _devctx.host_array((B,), fp)
        self._p_tlo = _devctx.host_array((B,), fp)
        self._p_lasth = _devctx.host_array((B,), fp)
        self._p_tc = None
        self._p_dout = None
        self._p_state.array[...] = state
        self._p_pars.array[...] = pars
        self._p_thi.array[...] = t_hi
        self._p_tlo.array[...] = t_lo
        self._outcome = np.zeros(B, dtype=np.int64)
        self._outcome_s = np.zeros(B, dtype=np.int64)
        self._h = np.zeros(B, dtype=fp)
        self._min_h = np.zeros(B, dtype=fp)
        self._max_h = np.zeros(B, dtype=fp)
        self._nsteps = np.zeros(B, dtype=np.uint64)
        self._step_res = [(taylor_outcome.success, fp(0))] * B
        self._prop_res = [(taylor_outcome.success, fp(0), fp(0), 0)] * B
        self._tc_valid = False
        self._tc_written = False
        if not getattr(self, "_lazy_ctx", False):
            self._ensure_ctx()  # the constructor fails loudly without libhy_cuda / a CUDA device

    # ---- device context ----
    def _ctx_key(self, device):
        return (id(self._dc), self._fp, self._B, self._tol, self._high_accuracy, self._compact_mode, device,
                tuple(int(e.direction) for e in self._t_events + self._nt_events),
                tuple(float(e.cooldown) for e in self._t_events))

    def _devices(self):
        d = self._device
        if d == "all":
            return list(range(max(1, _cabi.device_count())))
        return d if isinstance(d, list) else None

    def _ensure_ctx(self):
        if self._ctx_obj is not None:
            return self._ctx_obj
        fp, B = self._fp, self._B
        fp_bits = 64 if fp == np.float64 else 32
        ev_dir = [int(e.direction) for e in self._t_events + self._nt_events]
        ev_cd = [float(e.cooldown) for e in self._t_events]
        devs = self._devices()
        ctx = None
        if devs is not None and len(devs) > 1:
            ctx = _devctx.MultiContext(self._dc, fp_bits, B, self._tol, self._high_accuracy, devs,
                                       n_tevents=len(self._t_events), ev_dir=ev_dir or None,
                                       ev_cooldown=ev_cd or None, dc_ode=self._dc_ode, evt=self._evt,
                                       compact_mode=self._compact_mode)
        else:
            dev = devs[0] if devs else self._device
            ctx = _devctx.POOL.acquire(self._ctx_key(dev))
            if ctx is None and self._ctx_src is not None and not isinstance(self._ctx_src, _devctx.MultiContext) \
                    and self._ctx_src._ctx is not None and self._ctx_src._ctx.value:
                ctx = self._ctx_src.clone(dev)
            if ctx is None:
                ctx = _cabi.Context(self._dc, fp_bits, B, self._tol, self._high_accuracy, device=dev,
                                    n_tevents=len(self._t_events), ev_dir=ev_dir or None,
                                    ev_cooldown=ev_cd or None, dc_ode=self._dc_ode, evt=self._evt,
                                    compact_mode=self._compact_mode)
            else:
                # a recycled / cloned context carries another integrator's device-only data
                if self._t_events:
                    ctx.reset_cooldowns(-1)
                ctx.set_angle_reducer(())
        self._ctx_src = None
        self._ctx_obj = ctx
        self._reducer_idx = ()
        snap, self._snap = self._snap, None
        if snap:
            if snap.get("cooldowns") is not None and self._t_events:
                ctx.set_cooldowns(np.ascontiguousarray(snap["cooldowns"][0]),
                                  np.ascontiguousarray(snap["cooldowns"][1]))
            if snap.get("tc") is not None:
                ctx.set_tc(np.ascontiguousarray(snap["tc"]))
                self._tc_written = True
        # last_h lives on the device too (dense output reads it)
        if np.any(self._p_lasth.array != 0):
            ctx.set_last_h(self._p_lasth.array)
        return ctx

    @property
    def _ctx(self):
        return self._ensure_ctx()

    def _release_ctx(self):
        """Detach the device context (ensemble driver: the finished iteration keeps only its
        host data; the context goes back to the pool for the next iteration)."""
        ctx = self._ctx_obj
        if ctx is None:
            return
        self._snap = self._device_snapshot()
        self._ctx_obj = None
        if isinstance(ctx, _devctx.MultiContext):
            ctx.close()
        else:
            _devctx.POOL.release(self._ctx_key(ctx.device), ctx)

    def _device_snapshot(self):
        """Host copy of what lives only on the device: cooldowns, and tc if it was ever written."""
        if self._ctx_obj is None:
            return self._snap
        snap = {"cooldowns": self._get_cooldown_arrays(), "tc": None}
        if self._tc_written:
            snap["tc"] = np.array(self.tc)
        return snap

    # ---- read-only / writable views (expose_batch_integrators.cpp:394-518) ----
    @staticmethod
    def _ro(a):
        v = a.view()
        v.flags.writeable = False
        return v

    def _view(self, a, writeable=False):
        if a.size == 0:
            return a if writeable else self._ro(a)
        return np.asarray(_view_owner(self, a, writeable))

    @property
    def state(self):
        return self._view(self._p_state.array, True)

    @property
    def pars(self):
        return self._view(self._p_pars.array, True)

    @property
    def time(self):
        return self._view(self._p_thi.array)

    @property
    def dtime(self):
        return (self._view(self._p_thi.array), self._view(self._p_tlo.array))

    @property
    def last_h(self):
        return self._view(self._p_lasth.array)

    @property
    def tc(self):
        if self._p_tc is None:
            self._p_tc = _devctx.host_array((self._n, self._order + 1, self._B), self._fp)
        if not self._tc_valid:
            if self._ctx_obj is None and self._snap and self._snap.get("tc") is not None:
                self._p_tc.array[...] = self._snap["tc"]
            elif self._ctx_obj is not None or self._tc_written:
                self._ctx.get_tc(self._p_tc.array)
            self._tc_valid = True
        return self._view(self._p_tc.array)

    @property
    def d_output(self):
        if self._p_dout is None:
            self._p_dout = _devctx.host_array((self._n, self._B), self._fp)
        return self._view(self._p_dout.array)

    def set_time(self, tm):
        fp, B = self._fp, self._B
        if isinstance(tm, (list, tuple, np.ndarray)):
            arr = np.asarray(tm)
            if arr.ndim != 1 or arr.shape[0] != B:
                raise ValueError(
                    "Invalid number of new times specified in a Taylor integrator in batch mode: the "
                    "batch size is {}, but the number of specified times is {}".format(B, arr.size)
                )
            self._p_thi.array[...] = arr.astype(fp, casting="same_kind")
        else:
            _check_scalar_type(tm, fp, "set_time()")
            self._p_thi.array[...] = fp(tm)
        self._p_tlo.array[...] = 0

    def set_dtime(self, hi, lo):
        fp, B = self._fp, self._B
        hi_v = isinstance(hi, (list, tuple, np.ndarray))
        lo_v = isinstance(lo, (list, tuple, np.ndarray))
        if hi_v != lo_v:
            raise TypeError(
                "The two arguments to the set_dtime() method must be of the same type"
            )
        if hi_v:
            h = np.asarray(hi, dtype=fp)
            l = np.asarray(lo, dtype=fp)
            if h.shape != (B,) or l.shape != (B,):
                raise ValueError("Invalid number of new times specified in set_dtime()")
        else:
            _check_scalar_type(hi, fp, "set_dtime()")
            _check_scalar_type(lo, fp, "set_dtime()")
            h = np.full(B, hi, dtype=fp)
            l = np.full(B, lo, dtype=fp)
        if not (np.all(np.isfinite(h)) and np.all(np.isfinite(l))):
            raise ValueError("Non-finite time passed to set_dtime()")
        # Normalise (hi, lo): (1.0, 0.5) -> (1.5, 0.0)
        # (reference: _test_batch_integrator.py:472-475).
        s = h + l
        e = l - (s - h)
        self._p_thi.array[...] = s
        self._p_tlo.array[...] = e

    # ---- scalar properties ----
    @property
    def order(self):
        return self._order

    @property
    def tol(self):
        return self._fp(self._tol)

    @property
    def dim(self):
        return self._n

    @property
    def batch_size(self):
        return self._B

    @property
    def compact_mode(self):
        return self._compact_mode

    @property
    def high_accuracy(self):
        return self._high_accuracy

    @property
    def with_events(self):
        return bool(self._t_events or self._nt_events)

    @property
    def t_events(self):
        return list(self._t_events)

    @property
    def nt_events(self):
        return list(self._nt_events)

    @property
    def sys(self):
        return list(self._sys)

    @property
    def decomposition(self):
        """The opcode tape as a list of (opcode name, record) - the analogue
        of the reference's list of u-variable definitions."""
        return [(_dec.OP_NAMES[int(o["opcode"])], o) for o in self._dc.ops]

    @property
    def llvm_state(self):
        return _llvm_state_stub(self._llvm_kw, self)

    @property
    def step_res(self):
        if self._step_res is None:
            fp = self._fp
            self._step_res = [
                (_outcome_from_int(int(o)), fp(h)) for o, h in zip(self._outcome_s, self._h)
            ]
        return list(self._step_res)

    @property
    def propagate_res(self):
        # Built lazily: a 1M-lane list of tuples costs more than the transfer.
        if self._prop_res is None:
            fp = self._fp
            self._prop_res = [
                (_outcome_from_int(int(o)), fp(a), fp(b), int(s))
                for o, a, b, s in zip(self._outcome, self._min_h, self._max_h, self._nsteps)
            ]
        return list(self._prop_res)

    # bulk numpy forms (1M-lane lists of tuples are slow to build)
    @property
    def propagate_res_arrays(self):
        return (self._outcome.copy(), self._min_h.copy(), self._max_h.copy(), self._nsteps.copy())

    # ---- variational helpers (expose_batch_integrators.cpp:550-649) ----
    @property
    def is_variational(self):
        return self._vsys is not None

    @property
    def n_orig_sv(self):
        return self._vsys.n_orig_sv if self._vsys is not None else self._n

    def _need_var(self):
        if self._vsys is None:
            raise ValueError("The function cannot be invoked on a non-variational integrator")

    @property
    def vargs(self):
        self._need_var()
        return self._vsys.vargs_list

    @property
    def vorder(self):
        self._need_var()
        return self._vsys.order

    def get_vslice(self, order, component=None):
        self._need_var()
        return self._vsys.get_vslice(order, component)

    def get_mindex(self, i):
        self._need_var()
        return self._vsys.get_mindex(i)

    @property
    def tstate(self):
        self._need_var()
        if not hasattr(self, "_tstate"):
            self._tstate = np.zeros((self.n_orig_sv, self._B), dtype=self._fp)
        return self._view(self._tstate)

    def eval_taylor_map(self, inputs):
        """Reference: expose_batch_integrators.cpp:573-649 (argument checks and messages)."""
        self._need_var()
        if isinstance(inputs, np.ndarray):
            if inputs.dtype != self._fp:
                raise TypeError(
                    "Invalid dtype detected for the inputs of a Taylor map evaluation: the expected dtype "
                    "is '{}', but the dtype of the inputs array is '{}' instead".format(
                        np.dtype(self._fp), inputs.dtype))
            if not inputs.flags.c_contiguous:
                raise ValueError(
                    "Invalid inputs array detected in a Taylor map evaluation: the array is not C-style "
                    "contiguous, please consider using numpy.ascontiguousarray() to turn it into one")
        else:
            inputs = np.ascontiguousarray(inputs, dtype=self._fp)
        nv = len(self._vsys.vargs_list)
        if inputs.ndim != 2:
            raise ValueError(
                "The array of inputs provided for the evaluation of a Taylor map has {} dimension(s), "
                "but it must have 2 dimensions instead".format(inputs.ndim))
        if inputs.shape[0] != nv:
            raise ValueError(
                "The array of inputs provided for the evaluation of a Taylor map has {} row(s), "
                "but it must have {} row(s) instead".format(inputs.shape[0], nv))
        if inputs.shape[1] != self._B:
            raise ValueError(
                "The array of inputs provided for the evaluation of a Taylor map has {} column(s), "
                "but it must have {} column(s) instead".format(inputs.shape[1], self._B))
        if not hasattr(self, "_tstate"):
            self._tstate = np.zeros((self.n_orig_sv, self._B), dtype=self._fp)
        if np.may_share_memory(inputs, self._p_state.array) or np.may_share_memory(inputs, self._tstate):
            raise ValueError(
                "Invalid inputs array detected in a Taylor map evaluation: the array may overlap with the "
                "internal data of the integrator")
        # (written in place: tstate keeps referring to the same buffer, as in the reference)
        self._tstate[...] = self._vsys.eval_taylor_map(self._p_state.array, inputs)
        return self.tstate

    # ---- host <-> device sync ----
    def _push(self):
        self._ctx.upload(
            self._p_state.array, self._p_pars.array, self._p_thi.array, self._p_tlo.array
        )

    def _pull(self):
        self._ctx.download(
            self._p_state.array, self._p_thi.array, self._p_tlo.array, self._p_lasth.array
        )
        self._tc_valid = False

    def _vec_arg(self, x, what, allow_empty=False):
        """Scalar-or-vector argument -> array [B] (or None for "not given")."""
        fp, B = self._fp, self._B
        if isinstance(x, (list, tuple, np.ndarray)):
            arr = np.asarray(x)
            if allow_empty and arr.size == 0:
                return None
            if arr.dtype != fp:
                if fp == np.float32 or arr.dtype == np.float32 or arr.dtype.kind not in "f":
                    if not (arr.dtype.kind == "f" and fp == np.float64 and arr.dtype == np.float64):
                        if not all(self._scalar_ok(v) for v in np.asarray(x, dtype=object).ravel()):
                            raise TypeError(
                                "{}: incompatible element type for this integrator".format(what)
                            )
                arr = arr.astype(fp)
            if arr.ndim != 1 or arr.shape[0] != B:
                raise ValueError(
                    "Invalid number of {} specified in a Taylor integrator in batch mode: the batch "
                    "size is {}, but the number of specified values is {}".format(what, B, arr.size)
                )
            return np.ascontiguousarray(arr)
        _check_scalar_type(x, fp, what)
        return np.full(B, x, dtype=fp)

    def _scalar_ok(self, v):
        try:
            _check_scalar_type(v, self._fp, "")
            return True
        except TypeError:
            return False

    # ---- stepping (expose_batch_integrators.cpp:233-242) ----
    def step(self, max_delta_t=None, write_tc=False):
        if isinstance(max_delta_t, bool):
            # step(write_tc) positional form
            write_tc, max_delta_t = max_delta_t, None
        mdt = None
        if max_delta_t is not None:
            if not isinstance(max_delta_t, (list, tuple, np.ndarray)):
                raise TypeError("step(): max_delta_t must be a list of floating-point values")
            mdt = self._vec_arg(max_delta_t, "max_delta_t")
        self._do_step(mdt, False, write_tc)

    def step_backward(self, write_tc=False):
        self._do_step(None, True, write_tc)

    def _do_step(self, mdt, backward, write_tc):
        self._set_reducer(())
        self._push()
        self._ctx.step(mdt, backward, write_tc, self._outcome_s, self._h)
        self._pull()
        if write_tc:
            self._tc_written = True
        self._step_res = None
        term = self._dispatch_events()
        if term:
            # A terminal event whose callback returned True is "continuing":
            # outcome idx instead of -idx-1 (Event detection.ipynb cell 28).
            for lane, (ev, keep) in term.items():
                if keep:
                    self._outcome_s[lane] = ev

    # ---- propagate (expose_batch_integrators.cpp:243-314) ----
    def _wrap_callbacks(self, callback):
        from .callback import _normalise_callbacks

        return _normalise_callbacks(callback)

    def _set_reducer(self, idx):
        idx = tuple(int(i) for i in idx)
        if idx != self._reducer_idx:
            self._ctx.set_angle_reducer(idx)
            self._reducer_idx = idx

    def _split_builtin_callbacks(self, cbs):
        """Step callbacks that exist as device-side post-step ops (the reference implements
        angle_reducer in C++, expose_callbacks.cpp:67-72: no Python runs per step) are taken out
        of the host loop.  Returns (host callbacks, state indices to reduce) - only when ALL
        callbacks are builtins may the propagation stay on the device for more than one step."""
        from .callback import angle_reducer

        if cbs and all(isinstance(cb, angle_reducer) for cb in cbs):
            idx = []
            for cb in cbs:
                cb.pre_hook(self)
                idx += [i for i in cb._idx if i not in idx]
            return [], idx
        return cbs, []

    def _propagate(self, t, is_delta, max_steps, max_delta_t, callback, write_tc, c_output):
        fp, B = self._fp, self._B
        tt = self._vec_arg(t, "delta_t" if is_delta else "t")
        if not np.all(np.isfinite(tt)):
            raise ValueError(
                "A non-finite time was passed to the propagate_{}() function of an adaptive "
                "Taylor integrator in batch mode".format("for" if is_delta else "until")
            )
        mdt = self._check_mdt(max_delta_t)
        self._check_max_steps(max_steps)
        cbs, cb_ret = self._wrap_callbacks(callback)
        from .c_output import continuous_output_batch_impl

        host_cbs, red = self._split_builtin_callbacks(cbs)
        self._set_reducer(red)
        host_loop = bool(host_cbs) or self._needs_host_events()
        self._push()
        if not host_loop:
            self._ctx.propagate(tt, is_delta, max_steps, mdt, write_tc or c_output, c_output,
                                self._outcome, self._min_h, self._max_h, self._nsteps)
            self._pull()
            self._dispatch_events()
        else:
            self._propagate_host_loop(host_cbs, max_steps, mdt, write_tc, c_output, t=tt,
                                      is_delta=is_delta)
        if write_tc or c_output:
            self._tc_written = True
        self._prop_res = None
        cout = None
        if c_output:
            cout = continuous_output_batch_impl._from_integrator(self)
        return (cout, cb_ret)

    def _check_mdt(self, max_delta_t):
        mdt = self._vec_arg(max_delta_t, "max_delta_t", allow_empty=True)
        if mdt is not None:
            if np.any(np.isnan(mdt)):
                raise ValueError("A nan max_delta_t was passed to propagate_for/until/grid()")
            if np.any(mdt <= 0):
                raise ValueError("A non-positive max_delta_t was passed to propagate_for/until/grid()")
        return mdt

    @staticmethod
    def _check_max_steps(max_steps):
        if not isinstance(max_steps, (int, np.integer)) or isinstance(max_steps, bool) or max_steps < 0:
            raise TypeError("max_steps must be a non-negative integer")

    def _needs_host_events(self):
        return any(e.callback is not None for e in self._t_events) or bool(self._nt_events)

    def _propagate_host_loop(self, cbs, max_steps, mdt, write_tc, c_output, t=None, is_delta=False,
                             grid=None, grid_out=None):
        """Driver used when Python must run between steps (step callbacks:
        step_cb_utils.cpp:70-98; event callbacks: taylor_expose_events.cpp:109-138).

        The device keeps everything between launches (final times, step counters, min/max h,
        grid position, the continuous output being recorded): a launch runs every ACTIVE lane
        until it finishes or needs the host - after ONE step if there are step callbacks, else
        after a step that logged a non-terminal event, or at a terminal event - and hands the
        lane back with the internal outcome PAUSED.  The next launch resumes the lanes that go
        on; finished lanes are masked out and keep their results."""
        B = self._B
        for cb in cbs:
            if hasattr(cb, "pre_hook"):
                cb.pre_hook(self)
        active = np.ones(B, dtype=np.uint8)
        first = True
        stopped = None
        while np.any(active):
            if not first:
                self._push()  # callbacks may have edited state / pars
            self._ctx.propagate_ex(
                self._outcome, self._min_h, self._max_h, self._nsteps, t=t, is_delta=is_delta,
                max_steps=max_steps, max_delta_t=mdt, write_tc=bool(write_tc or c_output),
                c_output=(1 if first else 2) if c_output else 0, active=active, resume=not first,
                launch_steps=1 if cbs else 0, pause_on_nt=bool(self._nt_events), grid=grid,
                grid_out=grid_out)
            first = False
            self._pull()
            term = self._dispatch_events() or {}
            act = active.astype(bool)
            done = act & (self._outcome != _cabi.OUTCOME_PAUSED)
            for lane, (ev, keep) in term.items():
                if keep and act[lane]:
                    # continuing terminal event: the lane goes on (if it also sits at its final
                    # time the next launch reports time_limit without stepping)
                    done[lane] = False
            active[done] = 0
            stop = False
            for cb in cbs:
                r = cb(self)
                if not isinstance(r, (bool, np.bool_)):
                    raise TypeError(
                        "The call operator of a step callback is expected to return a boolean, "
                        "but a value of type \"{}\" was returned instead".format(type(r).__name__)
                    )
                if not r:
                    stop = True
            if stop:
                stopped = active.astype(bool)
                break
        if stopped is not None:
            self._outcome[stopped] = int(taylor_outcome.cb_stop)

    def _dispatch_events(self):
        """Drain the device event log and run the Python callbacks in
        chronological order per lane (taylor_expose_events.cpp:109-138)."""
        if not self.with_events:
            return None
        from .events import dispatch

        return dispatch(self)

    def propagate_for(self, delta_t, max_steps=0, max_delta_t=(), callback=None, write_tc=False,
                      c_output=False):
        return self._propagate(delta_t, True, max_steps, max_delta_t, callback, write_tc, c_output)

    def propagate_until(self, t, max_steps=0, max_delta_t=(), callback=None, write_tc=False,
                        c_output=False):
        return self._propagate(t, False, max_steps, max_delta_t, callback, write_tc, c_output)

    def propagate_grid(self, grid, max_steps=0, max_delta_t=(), callback=None):
        fp, B, n = self._fp, self._B, self._n
        g = np.asarray(grid)
        if g.ndim != 2:
            raise ValueError(
                "Invalid grid passed to the propagate_grid() method of a batch integrator: "
                "the expected number of dimensions is 2, but the input array has a dimension of {}".format(
                    g.ndim
                )
            )
        if g.shape[1] != B:
            raise ValueError(
                "Invalid grid passed to the propagate_grid() method of a batch integrator: "
                "the shape must be (n, {}) but the number of columns is {} instead".format(B, g.shape[1])
            )
        if g.shape[0] == 0:
            raise ValueError(
                "Cannot invoke propagate_grid() in an adaptive Taylor integrator in batch mode "
                "if the time grid is empty"
            )
        g = np.ascontiguousarray(g.astype(fp, casting="same_kind"))
        if not np.all(np.isfinite(g)):
            raise ValueError("A non-finite time value was passed to propagate_grid()")
        if g.shape[0] > 1:
            d = np.diff(g, axis=0)
            if not (np.all(d > 0) or np.all(d < 0)):
                raise ValueError("A non-monotonic time grid was passed to propagate_grid()")
        mdt = self._check_mdt(max_delta_t)
        self._check_max_steps(max_steps)
        cbs, cb_ret = self._wrap_callbacks(callback)
        host_cbs, red = self._split_builtin_callbacks(cbs)
        self._set_reducer(red)
        out = np.empty((g.shape[0], n, B), dtype=fp)
        self._push()
        if host_cbs or self._needs_host_events():
            self._propagate_host_loop(host_cbs, max_steps, mdt, True, False, grid=g, grid_out=out)
        else:
            self._ctx.propagate_grid(g, g.shape[0], max_steps, mdt, out, self._outcome, self._min_h,
                                     self._max_h, self._nsteps)
            self._pull()
            self._dispatch_events()
        self._tc_written = True
        self._prop_res = None
        return (cb_ret, out)

    # ---- dense output (expose_batch_integrators.cpp:519-541) ----
    def update_d_output(self, t, rel_time=False):
        tt = self._vec_arg(t, "t")
        if self._p_dout is None:
            self._p_dout = _devctx.host_array((self._n, self._B), self._fp)
        # Times may have been edited through set_time(): push them.
        self._ctx.upload(None, None, self._p_thi.array, self._p_tlo.array)
        self._ctx.dense_eval(tt, rel_time, self._p_dout.array)
        return self.d_output

    # ---- events ----
    @property
    def te_cooldowns(self):
        nte = len(self._t_events)
        el = np.zeros((self._B, max(nte, 1)), dtype=self._fp)
        tot = np.zeros((self._B, max(nte, 1)), dtype=self._fp)
        if nte:
            self._ctx.get_cooldowns(el, tot)
        out = []
        for l in range(self._B):
            out.append(
                [None if not (tot[l, e] >= 0) else (self._fp(el[l, e]), self._fp(tot[l, e]))
                 for e in range(nte)]
            )
        return out

    def reset_cooldowns(self, i=None):
        if not self._t_events:
            raise ValueError("No events were defined for this integrator")
        if i is None:
            self._ctx.reset_cooldowns(-1)
        else:
            if i >= self._B or i < 0:
                raise ValueError("Invalid batch index {} passed to reset_cooldowns()".format(i))
            self._ctx.reset_cooldowns(int(i))

    # ---- copy / pickle (expose_batch_integrators.cpp:665-669) ----
    def _state_dict(self):
        """Everything a copy needs, as host data (pickle_wrappers.hpp:35-73)."""
        return dict(
            fp=self._fp,
            sys=self._vsys if self._vsys is not None else self._sys,
            state=self._p_state.array.copy(),
            pars=self._p_pars.array.copy(),
            t_hi=self._p_thi.array.copy(),
            t_lo=self._p_tlo.array.copy(),
            last_h=self._p_lasth.array.copy(),
            tol=self._tol,
            high_accuracy=self._high_accuracy,
            compact_mode=self._compact_mode,
            parallel_mode=self._parallel_mode,
            t_events=self._t_events,
            nt_events=self._nt_events,
            llvm_kw=self._llvm_kw,
            device=self._device,
            step_res=self._step_res,
            prop_res=self._prop_res,
            outcome_s=self._outcome_s.copy(),
            h=self._h.copy(),
            res_arrays=(self._outcome.copy(), self._min_h.copy(), self._max_h.copy(),
                        self._nsteps.copy()),
            snap=self._device_snapshot(),
        )

    def _get_cooldown_arrays(self):
        nte = len(self._t_events)
        if not nte:
            return None
        if self._ctx_obj is None:
            return self._snap.get("cooldowns") if self._snap else None
        el = np.zeros((self._B, nte), dtype=self._fp)
        tot = np.zeros((self._B, nte), dtype=self._fp)
        self._ctx_obj.get_cooldowns(el, tot)
        return el, tot

    def _copy_impl(self, deep, memo=None):
        """Copy WITHOUT re-running the decomposition or touching the device: the copy shares the
        immutable tape, owns host copies of the lane data and gets its device context on first
        use - cloned from this integrator's (hy_clone) or taken from the pool of idle contexts."""
        cls = type(self)
        ta = cls.__new__(cls)
        for k in ("_vsys", "_sys", "_tol", "_order", "_high_accuracy", "_compact_mode",
                  "_parallel_mode", "_llvm_kw", "_device", "_dc", "_dc_ode", "_evt", "_B", "_n"):
            setattr(ta, k, getattr(self, k))
        if deep:
            ta._t_events = _copy.deepcopy(self._t_events, memo)
            ta._nt_events = _copy.deepcopy(self._nt_events, memo)
        else:
            ta._t_events = list(self._t_events)
            ta._nt_events = list(self._nt_events)
        ta._lazy_ctx = True
        ta._alloc(self._p_state.array, self._p_pars.array, self._p_thi.array, self._p_tlo.array)
        ta._p_lasth.array[...] = self._p_lasth.array
        ta._outcome[...] = self._outcome
        ta._outcome_s[...] = self._outcome_s
        ta._h[...] = self._h
        ta._min_h[...] = self._min_h
        ta._max_h[...] = self._max_h
        ta._nsteps[...] = self._nsteps
        ta._step_res = None if self._step_res is None else list(self._step_res)
        ta._prop_res = None if self._prop_res is None else list(self._prop_res)
        ta._snap = self._device_snapshot()
        ta._tc_written = self._tc_written
        ta._ctx_src = self._ctx_obj
        if hasattr(self, "_tstate"):
            ta._tstate = self._tstate.copy()
        for k, v in self.__dict__.items():
            if not k.startswith("_"):
                ta.__dict__[k] = _copy.deepcopy(v, memo) if deep else v
        return ta

    def __copy__(self):
        return self._copy_impl(False)

    def __deepcopy__(self, memo):
        return self._copy_impl(True, memo)

    def __getstate__(self):
        sd = self._state_dict()
        dyn = {k: v for k, v in self.__dict__.items() if not k.startswith("_")}
        return (sd, dyn)

    def __setstate__(self, st):
        sd, dyn = st
        fp = sd["fp"]
        self._lazy_ctx = True
        type(self).__init__(
            self, sd["sys"], sd["state"], time=sd["t_hi"],
            pars=sd["pars"] if sd["pars"].shape[0] else None, tol=fp(sd["tol"]),
            high_accuracy=sd["high_accuracy"], compact_mode=sd["compact_mode"],
            t_events=list(sd["t_events"]), nt_events=list(sd["nt_events"]),
            parallel_mode=sd["parallel_mode"], device=sd["device"], **sd["llvm_kw"])
        self._p_tlo.array[...] = sd["t_lo"]
        self._p_lasth.array[...] = sd["last_h"]
        self._step_res = None if sd["step_res"] is None else list(sd["step_res"])
        self._prop_res = None if sd["prop_res"] is None else list(sd["prop_res"])
        self._outcome_s[...] = sd["outcome_s"]
        self._h[...] = sd["h"]
        for dst, src in zip((self._outcome, self._min_h, self._max_h, self._nsteps), sd["res_arrays"]):
            dst[...] = src
        self._snap = sd.get("snap")
        self._tc_written = bool(self._snap and self._snap.get("tc") is not None)
        self.__dict__.update(dyn)

    def __repr__(self):
        return (
            "C++ datatype            : {}\nTolerance               : {}\nHigh accuracy           : {}\n"
            "Compact mode            : {}\nTaylor order            : {}\nDimension               : {}\n"
            "Batch size              : {}\nTime                    : {}\nState                   : {}\n"
            "Backend                 : libhy_cuda (sm_100a)\n"
        ).format(
            "double" if self._fp == np.float64 else "float", self._tol, self._high_accuracy,
            self._compact_mode, self._order, self._n, self._B, list(self._p_thi.array[:8]),
            list(self._p_state.array.ravel()[:8]),
        )


class taylor_adaptive_batch_dbl(taylor_adaptive_batch_impl):
    _fp = np.float64


class taylor_adaptive_batch_flt(taylor_adaptive_batch_impl):
    _fp = np.float32
