"""Step-callback plumbing (host side only).

Mirrors /root/reference/heyoka/step_cb_utils.cpp:46-203: a callback argument
is either one callable or a list of callables; callables are used BY
REFERENCE (never copied by propagate_*; _test_batch_integrator.py:184-194),
must return ``bool`` and may define ``pre_hook(ta)``.  ``angle_reducer``
mirrors the C++ callback exposed by expose_callbacks.cpp:67-72.
"""

import math

import numpy as np

from . import _expression as _E

__all__ = ["angle_reducer"]


def _normalise_callbacks(callback):
    """-> (list of callables, value to hand back to the caller)."""
    if callback is None:
        return [], None
    if isinstance(callback, (list, tuple)):
        cbs = list(callback)
        ret = list(callback)
    else:
        cbs = [callback]
        ret = callback
    for cb in cbs:
        if not callable(cb):
            raise TypeError(
                "An object of type \"{}\" cannot be used as a step callback because it is not "
                "callable".format(type(cb).__name__)
            )
        if hasattr(cb, "pre_hook") and not callable(cb.pre_hook):
            raise TypeError(
                "An object of type \"{}\" cannot be used as a step callback because its "
                "\"pre_hook\" attribute is not callable".format(type(cb).__name__)
            )
    return cbs, ret


class angle_reducer:
    """Reduce the selected state variables to [0, 2pi) after every step."""

    def __init__(self, var_list=()):
        self._vars = [v for v in var_list]
        for v in self._vars:
            if not isinstance(v, _E.expression) or v.kind != "var":
                raise ValueError("angle_reducer needs a list of variables")
        self._idx = None

    def pre_hook(self, ta):
        names = [l.name for l, _ in ta.sys]
        self._idx = [names.index(v.name) for v in self._vars]

    def __call__(self, ta):
        if self._idx is None:
            self.pre_hook(ta)
        two_pi = 2.0 * math.pi
        for i in self._idx:
            x = ta.state[i]
            x -= two_pi * np.floor(x / two_pi)
        return True

    def __repr__(self):
        return "Angle reducer: {}".format(self._vars)
