"""Ensemble propagation driver.

Same call surface and validation as the reference
(/root/reference/heyoka/__init__.py:251-336, _ensemble_impl.py:23-138):
``n_iter`` independent copies of ``ta`` are made, ``gen(copy, i)`` edits each
one, then every copy is propagated.  Here the iterations are sharded over the
visible GPUs (iteration i -> device i mod G) and driven
by one host thread per device; there is no inter-GPU communication, results
are gathered on the host in iteration order.
"""

import copy as _copy
from collections.abc import Iterable

import numpy as np

from . import _cabi


def _splat_grid(arg, ta):
    if hasattr(ta, "batch_size"):
        return np.repeat(arg, ta.batch_size).reshape((-1, ta.batch_size))
    return arg


def _ensemble_propagate_generic(tp, ta, arg, n_iter, gen, **kwargs):
    if not isinstance(n_iter, int):
        raise TypeError(
            "The n_iter parameter must be an integer, but an object of type {} was"
            " provided instead".format(type(n_iter))
        )
    if n_iter < 0:
        raise ValueError(
            "The n_iter parameter must be non-negative, but it is {} instead".format(n_iter)
        )
    if tp in ("until", "for"):
        if isinstance(arg, Iterable):
            raise TypeError(
                "Cannot perform an ensemble propagate_until/for(): the final epoch/time"
                " interval must be a scalar, not an iterable object"
            )
    else:
        arg = np.array(arg)
        if arg.ndim != 1:
            raise ValueError(
                "Cannot perform an ensemble propagate_grid(): the input time grid must"
                " be one-dimensional, but instead it has {} dimensions".format(arg.ndim)
            )
    if "max_delta_t" in kwargs and isinstance(kwargs["max_delta_t"], Iterable):
        raise TypeError(
            'Cannot perform an ensemble propagate_until/for/grid(): the "max_delta_t"'
            " argument must be a scalar, not an iterable object"
        )
    algo = kwargs.pop("algorithm", "thread")
    allowed_algos = ["thread", "process"]
    if algo not in allowed_algos:
        raise ValueError(
            "The parallelisation algorithm must be one of {}, but '{}' was provided instead".format(
                allowed_algos, algo
            )
        )
    max_workers = kwargs.pop("max_workers", None)
    if algo == "thread" and "chunksize" in kwargs:
        raise TypeError("propagate() got an unexpected keyword argument 'chunksize'")
    kwargs.pop("chunksize", None)
    return _run(tp, ta, arg, n_iter, gen, max_workers, kwargs)


def _run(tp, ta, arg, n_iter, gen, max_workers, kwargs):
    from concurrent.futures import ThreadPoolExecutor

    ndev = max(1, _cabi.device_count())
    if "callback" in kwargs:
        kwargs_list = []
        for _ in range(n_iter):
            kw = _copy.copy(kwargs)
            kw.update(callback=_copy.deepcopy(kwargs["callback"]))
            kwargs_list.append(kw)
    else:
        kwargs_list = [kwargs] * n_iter

    def func(i):
        local_ta = _copy.deepcopy(ta)
        if ndev > 1 and hasattr(local_ta, "_move_to_device"):
            local_ta._move_to_device(i % ndev)
        local_ta = gen(local_ta, i)
        if tp == "until":
            ret = local_ta.propagate_until(arg, **kwargs_list[i])
        elif tp == "for":
            ret = local_ta.propagate_for(arg, **kwargs_list[i])
        else:
            ret = local_ta.propagate_grid(_splat_grid(arg, ta), **kwargs_list[i])
        return (local_ta,) + tuple(ret)

    workers = max_workers if max_workers is not None else max(ndev, 1)
    with ThreadPoolExecutor(max_workers=workers) as ex:
        return list(ex.map(func, range(n_iter)))


def ensemble_propagate_until(ta, t, n_iter, gen, **kwargs):
    return _ensemble_propagate_generic("until", ta, t, n_iter, gen, **kwargs)


def ensemble_propagate_for(ta, delta_t, n_iter, gen, **kwargs):
    return _ensemble_propagate_generic("for", ta, delta_t, n_iter, gen, **kwargs)


def ensemble_propagate_grid(ta, grid, n_iter, gen, **kwargs):
    return _ensemble_propagate_generic("grid", ta, grid, n_iter, gen, **kwargs)


ensemble_propagate_until_batch = ensemble_propagate_until
ensemble_propagate_for_batch = ensemble_propagate_for
ensemble_propagate_grid_batch = ensemble_propagate_grid
