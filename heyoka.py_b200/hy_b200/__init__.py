"""hy_b200 - a B200-native batch Taylor integrator behind the heyoka.py API.

Drop-in for ONE hot path of heyoka.py (the ``taylor_adaptive_batch`` step /
propagate loop and the ensemble driver): ``import hy_b200 as hy``.  Names,
argument meaning and error behaviour follow /root/reference/heyoka/__init__.py
and the pybind11 layer it wraps.
"""

import numpy as _np

from . import model, callback
from ._expression import (
    expression, make_vars, par, time, sin, cos, exp, log, sqrt, pow, sum, prod, diff, square, tan,
    sinh, cosh, tanh, sigmoid, asinh, acosh, atanh, asin, acos, atan, erf,
)
from .enums import taylor_outcome, event_direction, code_model
from .var_ode_sys import var_ode_sys, var_args
from .batch import taylor_adaptive_batch_dbl, taylor_adaptive_batch_flt
from .events import (
    nt_event_batch_dbl, nt_event_batch_flt, t_event_batch_dbl, t_event_batch_flt,
)
from .c_output import continuous_output_batch_dbl, continuous_output_batch_flt
from .ensemble import (
    ensemble_propagate_until, ensemble_propagate_for, ensemble_propagate_grid,
    ensemble_propagate_until_batch, ensemble_propagate_for_batch, ensemble_propagate_grid_batch,
    set_serialization_backend, get_serialization_backend,
)

__version__ = "0.1.0"

_fp_to_suffix_dict = {_np.float32: "_flt", _np.float64: "_dbl", float: "_dbl"}


def _fp_to_suffix(fp_t):
    # Reference: /root/reference/heyoka/__init__.py:54-66.
    if not isinstance(fp_t, type):
        raise TypeError(
            'A Python type was expected in input, but an object of type "{}" was'
            " provided instead".format(type(fp_t))
        )
    if fp_t in _fp_to_suffix_dict:
        return _fp_to_suffix_dict[fp_t]
    raise TypeError('The floating-point type "{}" is not recognized/supported'.format(fp_t))


def taylor_adaptive_batch(sys, state, **kwargs):
    """Reference: /root/reference/heyoka/__init__.py:78-86."""
    fp_suffix = _fp_to_suffix(kwargs.pop("fp_type", float))
    return globals()["taylor_adaptive_batch{}".format(fp_suffix)](sys, state, **kwargs)


def nt_event_batch(ex, callback, **kwargs):
    fp_suffix = _fp_to_suffix(kwargs.pop("fp_type", float))
    return globals()["nt_event_batch{}".format(fp_suffix)](ex, callback, **kwargs)


def t_event_batch(ex, **kwargs):
    fp_suffix = _fp_to_suffix(kwargs.pop("fp_type", float))
    return globals()["t_event_batch{}".format(fp_suffix)](ex, **kwargs)


def recommended_simd_size(fp_type=float):
    """The reference returns the host SIMD width (4 for FP64 on AVX2).  On the
    GPU a "batch" is a whole shard; one warp's worth of lanes is the natural
    minimum granule."""
    _fp_to_suffix(fp_type)
    return 32  # one warp of lanes, whatever the precision (a GPU has no host-SIMD-width notion)
