"""``continuous_output_batch``: dense output over a whole propagation.

Mirror of /root/reference/heyoka/taylor_expose_c_output.cpp:260-526.  The
per-step Taylor coefficients and end times are recorded ON THE DEVICE by
hy_propagate(c_output=1); evaluation (per-lane bisection over the step times +
Horner) is one kernel launch for any number of query times.
"""

import numpy as np

from . import _cabi


class continuous_output_batch_impl:
    _fp = np.float64

    def __init__(self):
        self._ta = None

    def _check(self):
        if self._ta is None:
            raise ValueError("Cannot use a default-constructed continuous_output_batch object")

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
        co._ta = ta
        co._rec = rec
        co._n = ta._n
        co._B = B
        co._order = ta._order
        co._nsteps = ns
        co._S = int(S)
        co._tcs = None
        co._times = None
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

    def __call__(self, t):
        self._check()
        fp, B, n = self._fp, self._B, self._n
        if isinstance(t, (list, tuple, np.ndarray)):
            arr = np.asarray(t)
            if arr.ndim == 1:
                if arr.shape[0] != B:
                    raise ValueError(
                        "Invalid time array passed to a continuous_output_batch object: the "
                        "length must be {} but it is {} instead".format(B, arr.shape[0])
                    )
                tt = np.ascontiguousarray(arr.astype(fp)).reshape(1, B)
                out = np.zeros((1, n, B), dtype=fp)
                self._rec.eval(tt, 1, out)
                self._out = out[0]
                v = self._out.view()
                v.flags.writeable = False
                return v
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
        v = self._out.view()
        v.flags.writeable = False
        return v

    @property
    def output(self):
        self._check()
        v = self._out.view()
        v.flags.writeable = False
        return v

    @property
    def times(self):
        self._check()
        self._fetch()
        v = self._thi.view()
        v.flags.writeable = False
        return v

    @property
    def tcs(self):
        self._check()
        self._fetch()
        v = self._tcs.view()
        v.flags.writeable = False
        return v

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
        self._check()
        return self._B

    def __repr__(self):
        if self._ta is None:
            return "Default-constructed continuous_output_batch"
        return "Directions : ...\nN of steps : {}\n".format(self._S)


class continuous_output_batch_dbl(continuous_output_batch_impl):
    _fp = np.float64


class continuous_output_batch_flt(continuous_output_batch_impl):
    _fp = np.float32
