"""``continuous_output_batch``: dense output over a whole propagation.

Mirror of /root/reference/heyoka/taylor_expose_c_output.cpp:260-526.  The
per-step Taylor coefficients and end times are recorded ON THE DEVICE by
hy_propagate(c_output=1); evaluation (per-lane bisection over the step times +
Horner) is one kernel launch for any number of query times.
"""

import numpy as np

from . import _cabi


class _cout_llvm_state:
    """Inert stand-in for ``c_out.llvm_state`` (the reference JIT-compiles the evaluation function,
    taylor_expose_c_output.cpp:505-511; here it is `cout_eval_kernel` of libhy_cuda)."""

    def __init__(self, owner):
        self._owner = owner
        self.ir = "; libhy_cuda: continuous output is evaluated by hy::cout_eval_kernel (no LLVM IR)"

    def get_ir(self):
        return self.ir


class _view_owner:
    """Base object of the arrays a continuous_output_batch hands out: holds the object (the reference's
    arrays do: `sys.getrefcount(c_out)` grows by one per live view, test.py:1582-1585)."""

    def __init__(self, owner, a):
        self._owner = owner
        ai = dict(a.__array_interface__)
        ai["data"] = (ai["data"][0], True)
        self.__array_interface__ = ai
        self._keep = a


class continuous_output_batch_impl:
    _fp = np.float64

    def __init__(self):
        self._rec = None

    def _check(self):
        if self._rec is None:
            raise ValueError("Cannot use a default-constructed continuous_output_batch object")

    def _ro(self, a):
        if a.size == 0:
            v = a.view()
            v.flags.writeable = False
            return v
        return np.asarray(_view_owner(self, a))

    @classmethod
    def _from_integrator(cls, ta):
        co = (continuous_output_batch_dbl if ta._fp == np.float64 else continuous_output_batch_flt)()
        B = ta._B
        # the record is an object of its own (like the reference's continuous_output_batch): it
        # stays valid after the integrator moves on or is destroyed
        rec = ta._ctx.cout_detach()
        if rec is None:
            return None
        ns = np.zeros(B, dtype=np.uint64)
        S = rec.info(ns)
        if S == 0:
            rec.close()
            return None
        co._rec = rec
        co._n = ta._n
        co._B = B
        co._order = ta._order
        co._nsteps = ns
        co._S = int(S)
        co._tcs = None
        co._device = ta._device if isinstance(ta._device, int) else 0
        co._out = np.zeros((co._n, B), dtype=ta._fp)
        return co

    def _fetch(self):
        if self._tcs is None:
            fp = self._fp
            S = self._S
            self._tcs = np.zeros((S, self._n, self._order + 1, self._B), dtype=fp)
            self._thi = np.zeros((S + 1, self._B), dtype=fp)
            self._tlo = np.zeros((S + 1, self._B), dtype=fp)
            self._rec.get(self._tcs, self._thi, self._tlo, S)
            # a lane that recorded fewer steps than the longest one: its end time is repeated down the column
            # (the reference's batch integrator keeps stepping finished lanes with h = 0, so `times` is finite
            # everywhere: test.py:1686-1688); its tcs rows past the end stay NaN
            rows = np.minimum(np.arange(S + 1)[:, None], self._nsteps.astype(np.int64)[None, :])
            cols = np.arange(self._B)[None, :]
            self._thi = np.ascontiguousarray(self._thi[rows, cols])
            self._tlo = np.ascontiguousarray(self._tlo[rows, cols])

    # ---- value semantics (the reference's object can be copied, deep-copied and pickled): a copy shares the
    # device record, a deep copy / an unpickled object gets its own, rebuilt from the exported arrays
    # (hy_cout_from_host, include/hy_cuda_cout.h)
    def __copy__(self):
        co = type(self)()
        co.__dict__.update(self.__dict__)
        if self._rec is not None:
            co._out = self._out.copy()
        return co

    def __getstate__(self):
        d = {k: v for k, v in self.__dict__.items() if k not in ("_rec", "_tcs", "_thi", "_tlo")}
        if self._rec is not None:
            self._fetch()
            d["_arrays"] = (self._tcs, self._thi, self._tlo)
        return d

    def __setstate__(self, d):
        d = dict(d)
        arrays = d.pop("_arrays", None)
        self.__dict__.update(d)
        self._rec = None
        if arrays is not None:
            self._tcs, self._thi, self._tlo = arrays
            self._rec = _cabi.cout_from_host(self._fp, self._n, self._order, self._B, self._nsteps, self._tcs,
                                             self._thi, self._tlo, self._S, device=getattr(self, "_device", 0))

    def __deepcopy__(self, memo):
        import copy

        co = type(self)()
        co.__setstate__(copy.deepcopy(self.__getstate__(), memo))
        return co

    def __call__(self, time):
        self._check()
        t = time
        fp, B, n = self._fp, self._B, self._n
        if isinstance(t, (list, tuple, np.ndarray)):
            arr = np.asarray(t)
            if arr.ndim == 1:
                if arr.shape[0] != B:
                    raise ValueError(
                        "Invalid time array passed to a continuous_output_batch object: the "
                        "length must be {} but it is {} instead".format(B, arr.shape[0])
                    )
                tt = np.ascontiguousarray(arr, dtype=fp).reshape(1, B)
                out = np.zeros((1, n, B), dtype=fp)
                self._rec.eval(tt, 1, out)
                self._out = out[0]
                return self._ro(self._out)
            if arr.ndim == 2:
                if arr.shape[1] != B:
                    raise ValueError(
                        "Invalid time array passed to a continuous_output_batch object: the "
                        "number of columns must be {} but it is {} instead".format(B, arr.shape[1])
                    )
                k = arr.shape[0]
                # (large results land in a recycled page-locked buffer: _cabi.PinnedPool)
                out = _cabi.OUT_POOL.array((k, n, B), fp)
                if k:
                    tt = np.ascontiguousarray(arr, dtype=fp)
                    self._rec.eval(tt, k, out)
                return out
            raise ValueError(
                "Invalid time array passed to a continuous_output_batch object: the number of "
                "dimensions must be 1 or 2, but it is {} instead".format(arr.ndim)
            )
        tt = np.full((1, B), t, dtype=fp)
        out = np.zeros((1, n, B), dtype=fp)
        self._rec.eval(tt, 1, out)
        self._out = out[0]
        return self._ro(self._out)

    @property
    def output(self):
        if self._rec is None:
            return None
        return self._ro(self._out)

    @property
    def times(self):
        if self._rec is None:
            return None
        self._fetch()
        return self._ro(self._thi)

    @property
    def tcs(self):
        if self._rec is None:
            return None
        self._fetch()
        return self._ro(self._tcs)

    @property
    def bounds(self):
        self._check()
        self._fetch()
        idx = self._nsteps.astype(np.int64)
        t1 = self._thi[idx, np.arange(self._B)]
        return (self._thi[0].copy(), t1)

    @property
    def n_steps(self):
        self._check()
        return self._S

    @property
    def batch_size(self):
        return 0 if self._rec is None else self._B

    @property
    def llvm_state(self):
        return _cout_llvm_state(self)

    def __repr__(self):
        if self._rec is None:
            return "Default-constructed continuous_output_batch"
        t0, t1 = self.bounds
        dirs = ", ".join("forward" if b >= a else "backward" for a, b in zip(t0, t1))
        rng = ", ".join("[{}, {})".format(a, b) for a, b in zip(t0, t1))
        return "Directions : [{}]\nTime ranges: [{}]\nN of steps : {}\n".format(dirs, rng, self._S)


class continuous_output_batch_dbl(continuous_output_batch_impl):
    _fp = np.float64


class continuous_output_batch_flt(continuous_output_batch_impl):
    _fp = np.float32
