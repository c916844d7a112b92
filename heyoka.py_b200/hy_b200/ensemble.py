"""Ensemble propagation: the product's multi-GPU path.

Call surface and validation follow the reference (/root/reference/heyoka/__init__.py:251-336;
worker semantics _ensemble_impl.py:23-138): ``n_iter`` independent copies of ``ta`` are made,
``gen(copy, i)`` edits each, every copy is propagated, the list ``[(ta_i, *propagate_ret)]``
comes back in iteration order.

What is different underneath (SURVEY.md section 8e - shards only, no inter-GPU traffic):
 * iteration ``i`` runs on device ``i mod G``; the pool has several host threads per device so
   that one iteration's Python (deepcopy, ``gen``, result lists) overlaps another's kernel;
 * a copy is host-only until it propagates; it then borrows an idle device context of the
   right shape from a pool (or clones the template's with ``hy_clone``: no tape matching, no
   scheduling, no per-iteration cudaMalloc) and gives it back when it is done, keeping its
   results - so 250 000 iterations do not make 250 000 contexts;
 * ``algorithm="process"`` spawns worker processes and moves the integrator, ``gen`` and the
   keyword arguments through the active serialization backend, like the reference.
"""

import copy as _copy
import threading as _threading
from collections.abc import Iterable

import numpy as np

from . import _cabi

# ---- serialization backend (reference: __init__.py:182-232) ----
_s11n_backend_mutex = _threading.Lock()


def _make_s11n_backend_maps():
    import pickle

    ret = {"pickle": pickle}
    try:
        import cloudpickle

        ret["cloudpickle"] = cloudpickle
    except ImportError:
        pass
    try:
        import dill

        ret["dill"] = dill
    except ImportError:
        pass
    return ret, {v: k for k, v in ret.items()}


_s11n_backend_map, _s11n_backend_inv_map = _make_s11n_backend_maps()
_s11n_backend = _s11n_backend_map.get("cloudpickle", _s11n_backend_map["pickle"])


def set_serialization_backend(name):
    global _s11n_backend
    if not isinstance(name, str):
        raise TypeError(
            "The serialization backend must be specified as a string, but an object of"
            " type {} was provided instead".format(type(name))
        )
    if name not in _s11n_backend_map:
        raise ValueError(
            "The serialization backend '{}' is not valid. The valid backends are: {}".format(
                name, list(_s11n_backend_map.keys())
            )
        )
    with _s11n_backend_mutex:
        _s11n_backend = _s11n_backend_map[name]


def get_serialization_backend():
    with _s11n_backend_mutex:
        return _s11n_backend


def _splat_grid(arg, ta):
    if hasattr(ta, "batch_size"):
        return np.repeat(arg, ta.batch_size).reshape((-1, ta.batch_size))
    return arg


def _propagate(tp, local_ta, arg, grid_ta, kw):
    if tp == "until":
        return local_ta.propagate_until(arg, **kw)
    if tp == "for":
        return local_ta.propagate_for(arg, **kw)
    return local_ta.propagate_grid(_splat_grid(arg, grid_ta), **kw)


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
    if algo == "thread":
        if "chunksize" in kwargs:
            raise TypeError("propagate() got an unexpected keyword argument 'chunksize'")
        return _run_threads(tp, ta, arg, n_iter, gen, kwargs)
    return _run_processes(tp, ta, arg, n_iter, gen, kwargs)


def _place(local_ta, i, ndev):
    """Iteration i -> device i mod G (a template that is itself split over devices keeps its own
    placement)."""
    if isinstance(getattr(local_ta, "_device", None), int) and ndev > 1:
        local_ta._device = i % ndev


def _run_threads(tp, ta, arg, n_iter, gen, kwargs):
    from concurrent.futures import ThreadPoolExecutor

    max_workers = kwargs.pop("max_workers", None)
    ndev = max(1, _cabi.device_count())
    if "callback" in kwargs:
        # every iteration works on its own deep copy of the callback(s)
        kwargs_list = []
        for _ in range(n_iter):
            kw = _copy.copy(kwargs)
            kw.update(callback=_copy.deepcopy(kwargs["callback"]))
            kwargs_list.append(kw)
    else:
        kwargs_list = [kwargs] * n_iter

    def func(i):
        local_ta = _copy.deepcopy(ta)   # host-only: the device context is borrowed when it propagates
        _place(local_ta, i, ndev)
        local_ta = gen(local_ta, i)
        ret = _propagate(tp, local_ta, arg, ta, kwargs_list[i])
        if hasattr(local_ta, "_release_ctx"):
            local_ta._release_ctx()     # the context serves the next iteration; results stay
        return (local_ta,) + tuple(ret)

    # Several host threads per device: the Python part of one iteration overlaps the kernel of
    # another (ctypes releases the GIL during every libhy_cuda call).
    workers = max_workers if max_workers is not None else max(1, min(32, 4 * ndev))
    with ThreadPoolExecutor(max_workers=workers) as ex:
        return list(ex.map(func, range(n_iter)))


def _mp_propagate(tup):
    tp, ta_s, gen_s, arg, kwargs_s, i, s11n_str = tup
    be = _s11n_backend_map[s11n_str]
    ta, gen, kwargs = be.loads(ta_s), be.loads(gen_s), be.loads(kwargs_s)
    _place(ta, i, max(1, _cabi.device_count()))
    local_ta = gen(ta, i)
    ret = _propagate(tp, local_ta, arg, ta, kwargs)
    return be.dumps((local_ta,) + tuple(ret))


def _run_processes(tp, ta, arg, n_iter, gen, kwargs):
    from concurrent.futures import ProcessPoolExecutor
    import multiprocessing as mp

    be = get_serialization_backend()
    s11n_str = _s11n_backend_inv_map[be]
    ctx = mp.get_context("spawn")  # CUDA cannot be used in a forked child
    max_workers = kwargs.pop("max_workers", None)
    chunksize = kwargs.pop("chunksize", 1)
    ta_s, gen_s, kw_s = be.dumps(ta), be.dumps(gen), be.dumps(kwargs)
    with ProcessPoolExecutor(max_workers=max_workers, mp_context=ctx) as ex:
        ret = list(
            ex.map(
                _mp_propagate,
                ((tp, ta_s, gen_s, arg, kw_s, i, s11n_str) for i in range(n_iter)),
                chunksize=chunksize,
            )
        )
    return [be.loads(r) for r in ret]


def ensemble_propagate_until(ta, t, n_iter, gen, **kwargs):
    return _ensemble_propagate_generic("until", ta, t, n_iter, gen, **kwargs)


def ensemble_propagate_for(ta, delta_t, n_iter, gen, **kwargs):
    return _ensemble_propagate_generic("for", ta, delta_t, n_iter, gen, **kwargs)


def ensemble_propagate_grid(ta, grid, n_iter, gen, **kwargs):
    return _ensemble_propagate_generic("grid", ta, grid, n_iter, gen, **kwargs)


ensemble_propagate_until_batch = ensemble_propagate_until
ensemble_propagate_for_batch = ensemble_propagate_for
ensemble_propagate_grid_batch = ensemble_propagate_grid
