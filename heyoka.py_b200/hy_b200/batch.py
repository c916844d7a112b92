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
from . import decompose as _dec
from . import _expression as _E
from .enums import taylor_outcome, code_model, _outcome_from_int
from .events import nt_event_batch_impl, t_event_batch_impl
from .var_ode_sys import var_ode_sys as _var_ode_sys

_LLVM_KW = ("opt_level", "force_avx512", "slp_vectorize", "fast_math", "code_model", "parjit")


class _llvm_state_stub:
    """Inert stand-in for ``ta.llvm_state`` (the reference exposes the JIT
    state, expose_batch_integrators.cpp:670; there is no LLVM here)."""

    def __init__(self, kw):
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
        else:
            sys_list = list(sys)
        self._sys = [(l, _E._wrap(r)) for l, r in sys_list]
        n = len(self._sys)
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
        self._device = int(kw.get("device", 0))
        self._t_events = t_events
        self._nt_events = nt_events
        ev_exprs = [e.expression for e in t_events] + [e.expression for e in nt_events]
        self._dc = _dec.decompose(self._sys, self._order, events=ev_exprs)
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

        self._B = B
        self._n = n
        self._alloc(state_, pars_, time_, np.zeros(B, dtype=fp))

    # ------------------------------------------------------------------
    def _alloc(self, state, pars, t_hi, t_lo):
        """Create the device context and the pinned host mirrors."""
        fp, B, n, m, p = self._fp, self._B, self._n, self._dc.n_par, self._order
        ev_dir = [int(e.direction) for e in self._t_events + self._nt_events]
        ev_cd = [float(e.cooldown) for e in self._t_events]
        self._ctx = _cabi.Context(
            self._dc,
            64 if fp == np.float64 else 32,
            B,
            self._tol,
            self._high_accuracy,
            device=self._device,
            n_tevents=len(self._t_events),
            ev_dir=ev_dir if ev_dir else None,
            ev_cooldown=ev_cd if ev_cd else None,
        )
        self._p_state = _cabi.PinnedArray((n, B), fp)
        self._p_pars = _cabi.PinnedArray((m, B), fp)
        self._p_thi = _cabi.PinnedArray((B,), fp)
        self._p_tlo = _cabi.PinnedArray((B,), fp)
        self._p_lasth = _cabi.PinnedArray((B,), fp)
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

    # ---- read-only / writable views (expose_batch_integrators.cpp:394-518) ----
    @staticmethod
    def _ro(a):
        v = a.view()
        v.flags.writeable = False
        return v

    @property
    def state(self):
        return self._p_state.array

    @property
    def pars(self):
        return self._p_pars.array

    @property
    def time(self):
        return self._ro(self._p_thi.array)

    @property
    def dtime(self):
        return (self._ro(self._p_thi.array), self._ro(self._p_tlo.array))

    @property
    def last_h(self):
        return self._ro(self._p_lasth.array)

    @property
    def tc(self):
        if self._p_tc is None:
            self._p_tc = _cabi.PinnedArray((self._n, self._order + 1, self._B), self._fp)
        if not self._tc_valid:
            self._ctx.get_tc(self._p_tc.array)
            self._tc_valid = True
        return self._ro(self._p_tc.array)

    @property
    def d_output(self):
        if self._p_dout is None:
            self._p_dout = _cabi.PinnedArray((self._n, self._B), self._fp)
        return self._ro(self._p_dout.array)

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
        return _llvm_state_stub(self._llvm_kw)

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
        return self._ro(self._tstate)

    def eval_taylor_map(self, inputs):
        self._need_var()
        inputs = np.asarray(inputs, dtype=self._fp)
        nv = len(self._vsys.vargs_list)
        if inputs.shape != (nv, self._B):
            raise ValueError(
                "Invalid inputs array passed to eval_taylor_map(): the expected shape is "
                "({}, {}) but the shape of the input is {}".format(nv, self._B, inputs.shape)
            )
        self._tstate = self._vsys.eval_taylor_map(self._p_state.array, inputs).astype(self._fp)
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
        self._push()
        self._ctx.step(mdt, backward, write_tc, self._outcome_s, self._h)
        self._pull()
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

    def _propagate(self, t, is_delta, max_steps, max_delta_t, callback, write_tc, c_output):
        fp, B = self._fp, self._B
        tt = self._vec_arg(t, "delta_t" if is_delta else "t")
        if not np.all(np.isfinite(tt)):
            raise ValueError(
                "A non-finite time was passed to the propagate_{}() function of an adaptive "
                "Taylor integrator in batch mode".format("for" if is_delta else "until")
            )
        mdt = self._vec_arg(max_delta_t, "max_delta_t", allow_empty=True)
        if mdt is not None:
            if np.any(np.isnan(mdt)):
                raise ValueError("A nan max_delta_t was passed to propagate_for/until()")
            if np.any(mdt <= 0):
                raise ValueError("A non-positive max_delta_t was passed to propagate_for/until()")
        if not isinstance(max_steps, (int, np.integer)) or isinstance(max_steps, bool) or max_steps < 0:
            raise TypeError("max_steps must be a non-negative integer")
        cbs, cb_ret = self._wrap_callbacks(callback)
        from .c_output import continuous_output_batch_impl

        host_loop = bool(cbs) or self._needs_host_events()
        self._push()
        if not host_loop:
            self._ctx.propagate(tt, is_delta, max_steps, mdt, write_tc or c_output, c_output,
                                self._outcome, self._min_h, self._max_h, self._nsteps)
            self._pull()
            self._dispatch_events()
        else:
            self._propagate_host_loop(tt, is_delta, max_steps, mdt, cbs, write_tc, c_output)
        self._prop_res = None
        cout = None
        if c_output:
            cout = continuous_output_batch_impl._from_integrator(self)
        return (cout, cb_ret)

    def _needs_host_events(self):
        return any(e.callback is not None for e in self._t_events) or bool(self._nt_events)

    def _propagate_host_loop(self, tt, is_delta, max_steps, mdt, cbs, write_tc, c_output):
        """Step-by-step driver used when Python must run between steps (step
        callbacks: step_cb_utils.cpp:70-98; event callbacks:
        taylor_expose_events.cpp:109-138).  One kernel launch per batch step."""
        fp, B = self._fp, self._B
        if is_delta:
            # Fix the absolute final times once (double-length).
            hi = self._p_thi.array.copy()
            lo = self._p_tlo.array.copy()
            s = hi + tt
            bb = s - hi
            err = (hi - (s - bb)) + (tt - bb) + lo
            tf = (s + err).astype(fp)
        else:
            tf = tt
        for cb in cbs:
            if hasattr(cb, "pre_hook"):
                cb.pre_hook(self)
        tot_n = np.zeros(B, dtype=np.uint64)
        mn = np.full(B, np.inf, dtype=fp)
        mx = np.zeros(B, dtype=fp)
        final = np.full(B, int(taylor_outcome.time_limit), dtype=np.int64)
        active = np.ones(B, dtype=bool)
        oc1 = np.zeros(B, dtype=np.int64)
        a1 = np.zeros(B, dtype=fp)
        b1 = np.zeros(B, dtype=fp)
        n1 = np.zeros(B, dtype=np.uint64)
        first = True
        while np.any(active):
            if not first:
                self._push()
            first = False
            # Inactive lanes are parked by asking them to go nowhere.
            target = np.where(active, tf, self._p_thi.array).astype(fp)
            if np.any(~active):
                self._park = True
            self._ctx.propagate(target, 0, 1, mdt, write_tc or c_output, c_output,
                                oc1, a1, b1, n1)
            self._pull()
            term = self._dispatch_events() or {}
            stepped = active & (n1 > 0)
            tot_n[stepped] += n1[stepped]
            succ = stepped & np.isfinite(a1) & (b1 > 0)
            mn = np.where(succ & (a1 < mn), a1, mn).astype(fp)
            mx = np.where(succ & (b1 > mx), b1, mx).astype(fp)
            done = active & (oc1 != int(taylor_outcome.step_limit))
            for lane, (ev, keep) in term.items():
                if keep and active[lane]:
                    # continuing terminal event: the lane goes on (unless it also
                    # reached its final time, which the next launch reports)
                    done[lane] = False
            final[done] = oc1[done]
            active &= ~done
            if max_steps:
                lim = active & (tot_n >= max_steps)
                final[lim] = int(taylor_outcome.step_limit)
                active &= ~lim
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
                final[active] = int(taylor_outcome.cb_stop)
                break
        self._outcome[...] = final
        self._min_h[...] = mn
        self._max_h[...] = mx
        self._nsteps[...] = tot_n

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
        mdt = self._vec_arg(max_delta_t, "max_delta_t", allow_empty=True)
        cbs, cb_ret = self._wrap_callbacks(callback)
        if cbs or self._needs_host_events():
            raise NotImplementedError(
                "propagate_grid() with step/event callbacks is not available in this build"
            )
        out = np.empty((g.shape[0], n, B), dtype=fp)
        self._push()
        self._ctx.propagate_grid(g, g.shape[0], max_steps, mdt, out, self._outcome, self._min_h,
                                 self._max_h, self._nsteps)
        self._pull()
        self._dispatch_events()
        self._prop_res = None
        return (cb_ret, out)

    # ---- dense output (expose_batch_integrators.cpp:519-541) ----
    def update_d_output(self, t, rel_time=False):
        tt = self._vec_arg(t, "t")
        if self._p_dout is None:
            self._p_dout = _cabi.PinnedArray((self._n, self._B), self._fp)
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
            step_res=self.step_res,
            prop_res=self.propagate_res,
            outcome_s=self._outcome_s.copy(),
            h=self._h.copy(),
            res_arrays=(self._outcome.copy(), self._min_h.copy(), self._max_h.copy(),
                        self._nsteps.copy()),
            tc=np.array(self.tc),
            cooldowns=self._get_cooldown_arrays(),
        )

    def _get_cooldown_arrays(self):
        nte = len(self._t_events)
        if not nte:
            return None
        el = np.zeros((self._B, nte), dtype=self._fp)
        tot = np.zeros((self._B, nte), dtype=self._fp)
        self._ctx.get_cooldowns(el, tot)
        return el, tot

    @classmethod
    def _from_state_dict(cls, sd, dyn=None, deep=True):
        fp = sd["fp"]
        tev = _copy.deepcopy(sd["t_events"]) if deep else list(sd["t_events"])
        ntev = _copy.deepcopy(sd["nt_events"]) if deep else list(sd["nt_events"])
        ta = cls(
            sd["sys"], sd["state"], time=sd["t_hi"], pars=sd["pars"] if sd["pars"].shape[0] else None,
            tol=fp(sd["tol"]), high_accuracy=sd["high_accuracy"], compact_mode=sd["compact_mode"],
            t_events=tev, nt_events=ntev, parallel_mode=sd["parallel_mode"], device=sd["device"],
            **sd["llvm_kw"],
        )
        ta._p_tlo.array[...] = sd["t_lo"]
        ta._p_lasth.array[...] = sd["last_h"]
        ta._step_res = list(sd["step_res"])
        ta._prop_res = list(sd["prop_res"])
        ta._outcome_s[...] = sd["outcome_s"]
        ta._h[...] = sd["h"]
        for dst, src in zip((ta._outcome, ta._min_h, ta._max_h, ta._nsteps), sd["res_arrays"]):
            dst[...] = src
        ta._restore_device_extras(sd)
        if dyn:
            ta.__dict__.update(dyn)
        return ta

    def _restore_device_extras(self, sd):
        """tc / last_h / cooldowns live on the device: push them back."""
        from . import _cabi as cabi

        cds = sd.get("cooldowns")
        if cds is not None and self._t_events:
            self._ctx.set_cooldowns(np.ascontiguousarray(cds[0]), np.ascontiguousarray(cds[1]))
        self._saved_tc = sd.get("tc")
        if self._saved_tc is not None:
            if self._p_tc is None:
                self._p_tc = cabi.PinnedArray((self._n, self._order + 1, self._B), self._fp)
            self._p_tc.array[...] = self._saved_tc
            self._tc_valid = True

    _OWN = None

    def _dyn_attrs(self):
        own = type(self)._OWN
        return {k: v for k, v in self.__dict__.items() if not k.startswith("_") or k not in own}

    def __copy__(self):
        sd = self._state_dict()
        dyn = {k: v for k, v in self.__dict__.items() if not k.startswith("_")}
        return type(self)._from_state_dict(sd, dyn, deep=False)

    def __deepcopy__(self, memo):
        sd = self._state_dict()
        dyn = {k: _copy.deepcopy(v, memo) for k, v in self.__dict__.items() if not k.startswith("_")}
        return type(self)._from_state_dict(sd, dyn, deep=True)

    def __getstate__(self):
        sd = self._state_dict()
        dyn = {k: v for k, v in self.__dict__.items() if not k.startswith("_")}
        return (sd, dyn)

    def __setstate__(self, st):
        sd, dyn = st
        other = type(self)._from_state_dict(sd, dyn, deep=False)
        self.__dict__.update(other.__dict__)
        # `other` must not free the context we just adopted.
        other.__dict__.clear()

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
