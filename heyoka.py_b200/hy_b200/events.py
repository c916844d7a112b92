"""Event classes of the batch integrator.

Mirror of /root/reference/heyoka/taylor_expose_events.cpp:185-317 (classes)
and :78-182 (callback wrapper): non-terminal callbacks have signature
``(ta, t, d_sgn, batch_idx) -> None``, terminal ones
``(ta, d_sgn, batch_idx) -> bool``; callbacks are deep-copied on event
construction (:83).  Detection itself happens on the device
(hy_kernels.cuh); the host replays the drained event log through the
callbacks in chronological order per lane.
"""

import copy as _copy

import numpy as np

from . import _expression as _E
from .enums import event_direction


def _check_dir(direction):
    if not isinstance(direction, event_direction):
        try:
            direction = event_direction(direction)
        except Exception:
            raise ValueError(
                "Invalid value selected for the direction of an event: the value must be one of "
                "'event_direction.any', 'event_direction.positive' or 'event_direction.negative'"
            )
    return direction


_OWN = ("expression", "callback", "direction", "cooldown")


class _event_base:
    """Copy / pickle semantics of the reference's event classes (copy_wrapper / deepcopy_wrapper,
    common_utils.hpp; pickle_wrappers.hpp:35-73): the callback is always deep-copied (value semantics of the
    C++ wrapper, taylor_expose_events.cpp:83), dynamic attributes are shared by `copy` and deep-copied by
    `deepcopy`; the callback is serialised with the active backend (cloudpickle by default, so lambdas and
    local classes travel: `set_serialization_backend`)."""

    def _extras(self):
        return {k: v for k, v in self.__dict__.items() if k not in _OWN}

    def __copy__(self):
        new = self._clone()
        new.__dict__.update(self._extras())
        return new

    def __deepcopy__(self, memo):
        new = self._clone()
        new.__dict__.update(_copy.deepcopy(self._extras(), memo))
        return new

    def __getstate__(self):
        from .ensemble import get_serialization_backend

        d = dict(self.__dict__)
        d["callback"] = None if self.callback is None else get_serialization_backend().dumps(self.callback)
        return d

    def __setstate__(self, d):
        from .ensemble import get_serialization_backend

        d = dict(d)
        if d.get("callback") is not None:
            d["callback"] = get_serialization_backend().loads(d["callback"])
        self.__dict__.update(d)


def _not_callable(callback):
    return TypeError("An object of type '{}' cannot be used as an event callback because it is not "
                     "callable".format(str(type(callback))))


class nt_event_batch_impl(_event_base):
    _fp = np.float64

    def __init__(self, ex, callback, direction=event_direction.any):
        if not isinstance(ex, _E.expression):
            raise TypeError("An event needs an expression as first argument")
        if callback is None or not callable(callback):
            raise _not_callable(callback)
        self.expression = ex
        self.callback = _copy.deepcopy(callback)
        self.direction = _check_dir(direction)

    def __repr__(self):
        return (
            "C++ datatype   : {}\nEvent type     : non-terminal\nEvent expression: {}\nEvent direction: "
            "event_direction::{}\nBatch mode     : true\n".format(
                "double" if self._fp == np.float64 else "float", self.expression, self.direction.name
            )
        )

    def _clone(self):
        return type(self)(self.expression, self.callback, direction=self.direction)


class t_event_batch_impl(_event_base):
    _fp = np.float64

    def __init__(self, ex, callback=None, direction=event_direction.any, cooldown=-1):
        if not isinstance(ex, _E.expression):
            raise TypeError("An event needs an expression as first argument")
        if callback is not None and not callable(callback):
            raise _not_callable(callback)
        self.expression = ex
        self.callback = _copy.deepcopy(callback) if callback is not None else None
        self.direction = _check_dir(direction)
        cd = float(cooldown)
        if cd != cd or cd in (float("inf"), float("-inf")):
            raise ValueError("Cannot set a non-finite cooldown value for a terminal event")
        self.cooldown = self._fp(cd)

    def __repr__(self):
        return (
            "C++ datatype   : {}\nEvent type     : terminal\nEvent expression: {}\nEvent direction: "
            "event_direction::{}\nWith callback  : {}\nCooldown       : {}\nBatch mode     : true\n".format(
                "double" if self._fp == np.float64 else "float", self.expression, self.direction.name,
                "yes" if self.callback is not None else "no",
                "auto" if self.cooldown < 0 else self.cooldown,
            )
        )

    def _clone(self):
        return type(self)(self.expression, self.callback, direction=self.direction, cooldown=self.cooldown)


class nt_event_batch_dbl(nt_event_batch_impl):
    _fp = np.float64


class nt_event_batch_flt(nt_event_batch_impl):
    _fp = np.float32


class t_event_batch_dbl(t_event_batch_impl):
    _fp = np.float64


class t_event_batch_flt(t_event_batch_impl):
    _fp = np.float32


def dispatch(ta):
    """Drain the device event log of ``ta`` and run the callbacks in
    chronological order per lane.

    Returns ``None`` when nothing was logged, else a dict
    ``lane -> (event index, keep_going)`` for the terminal events that fired
    (``keep_going`` is the callback's return value, False without callback)."""
    recs = ta._ctx.events_drain()
    if len(recs) == 0:
        return None
    nte = len(ta._t_events)
    # chronological per lane: increasing t forward in time, decreasing t backward
    sgn = np.where(np.signbit(ta._p_lasth.array[recs["lane"]]), -1.0, 1.0)
    order = np.lexsort((recs["ev_idx"], recs["t"] * sgn, recs["step"], recs["lane"]))
    recs = recs[order]
    term = {}
    for r in recs:
        ev = int(r["ev_idx"])
        lane = int(r["lane"])
        if ev >= nte:
            cb = ta._nt_events[ev - nte].callback
            # (errors raised while the reference's C++ loop calls back into Python surface as RuntimeError:
            #  /root/reference/heyoka/test.py:836-846, :968-999)
            try:
                cb(ta, ta._fp(r["t"]), int(r["d_sgn"]), lane)
            except TypeError as e:
                raise RuntimeError(
                    "The call operator of a non-terminal event callback has an incompatible "
                    "signature: {}".format(e)
                )
        else:
            cb = ta._t_events[ev].callback
            keep = False
            if cb is not None:
                try:
                    keep = cb(ta, int(r["d_sgn"]), lane)
                except TypeError as e:
                    raise RuntimeError(
                        "The call operator of a terminal event callback has an incompatible "
                        "signature: {}".format(e)
                    )
                if not isinstance(keep, (bool, np.bool_)):
                    # (taylor_expose_events.cpp:128-134)
                    raise RuntimeError(
                        "Unable to convert a Python object of type '{}' to the C++ type 'bool' in the "
                        "construction of the return value of an event callback".format(type(keep))
                    )
            term[lane] = (ev, bool(keep))
    return term
