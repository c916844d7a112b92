"""CPU: the batch event classes, scenario by scenario as in the reference's own test
(/root/reference/heyoka/test.py:417-691, event_classes_test_case.test_basic, batch part)."""

import gc
import pickle
from copy import copy, deepcopy

import pytest

import hy_b200 as hy
from hy_b200 import event_direction, nt_event_batch, t_event_batch


def test_nt_event_batch_class():
    x, v = hy.make_vars("x", "v")
    for kw, d in (({}, "any"), ({"fp_type": float}, "any"), ({"direction": event_direction.positive}, "positive"),
                  ({"direction": event_direction.negative}, "negative")):
        ev = nt_event_batch(ex=x + v, callback=lambda _: _, **kw)
        assert " non-terminal" in repr(ev) and "(x + v)" in repr(ev) and "event_direction::" + d in repr(ev)
        assert ev.expression == x + v and ev.direction == getattr(event_direction, d) and ev.callback is not None
    ev = nt_event_batch(x + v, lambda _: _)
    assert " non-terminal" in repr(ev)

    class local_cb:
        def __init__(self):
            self.n = 0

        def __call__(self, ta, t, d_sgn):
            self.n = self.n + 1

    lcb = local_cb()
    ev = nt_event_batch(ex=x + v, callback=lcb, direction=event_direction.negative)
    assert ev.callback.n == 0
    cb = ev.callback
    for _ in range(3):
        cb(1, 2, 3)
    assert ev.callback.n == 3
    ev.callback.n = 0
    assert ev.callback.n == 0 and id(lcb) != id(ev.callback)

    with pytest.raises(ValueError) as cm:
        nt_event_batch(ex=x + v, callback=lambda _: _, direction=event_direction(10))
    assert "10 is not a valid event_direction" in str(cm.value)
    for bad in (3, None):
        with pytest.raises(TypeError) as cm:
            nt_event_batch(ex=x + v, callback=bad)
        assert ("An object of type '{}' cannot be used as an event callback because it is not callable".format(
            str(type(bad)))) in str(cm.value)

    ev = nt_event_batch(ex=x + v, callback=lambda _: _, direction=event_direction.negative)
    ev = pickle.loads(pickle.dumps(ev))
    assert " non-terminal" in repr(ev) and "(x + v)" in repr(ev) and "event_direction::negative" in repr(ev)
    ev.foo = "hello world"
    ev = pickle.loads(pickle.dumps(ev))
    assert ev.foo == "hello world"

    class foo:
        pass

    ev.bar = foo()
    assert id(ev.bar) == id(copy(ev).bar) and id(ev.bar) != id(deepcopy(ev).bar)

    ev = nt_event_batch(ex=x + v, callback=local_cb(), direction=event_direction.negative)
    out_cb = ev.callback
    del ev
    gc.collect()
    for _ in range(3):
        out_cb(1, 2, 3)
    assert out_cb.n == 3
    with pytest.raises(TypeError) as cm:
        nt_event_batch(x + v, lambda _: _, fp_type=str)
    assert 'The floating-point type "{}" is not recognized/supported'.format(str) in str(cm.value)


def test_t_event_batch_class():
    x, v = hy.make_vars("x", "v")
    fp_t = float
    ev = t_event_batch(x + v)
    r = repr(ev)
    assert " terminal" in r and "(x + v)" in r and "event_direction::any" in r and ": no" in r and "auto" in r
    assert ev.expression == x + v and ev.direction == event_direction.any and ev.cooldown == fp_t(-1)
    assert ev.callback is None
    ev = t_event_batch(x + v, direction=event_direction.negative, cooldown=fp_t(3))
    r = repr(ev)
    assert " terminal" in r and "event_direction::negative" in r and ": no" in r and "3" in r
    assert ev.direction == event_direction.negative and ev.cooldown == fp_t(3) and ev.callback is None
    ev = t_event_batch(x + v, direction=event_direction.positive, cooldown=fp_t(3), callback=lambda _: _)
    r = repr(ev)
    assert " terminal" in r and "event_direction::positive" in r and ": yes" in r and "3" in r
    assert ev.cooldown == fp_t(3) and ev.callback is not None

    class local_cb:
        def __init__(self):
            self.n = 0

        def __call__(self, ta, d_sgn):
            self.n = self.n + 1

    lcb = local_cb()
    ev = t_event_batch(x + v, direction=event_direction.positive, cooldown=fp_t(3), callback=lcb)
    assert ": yes" in repr(ev) and ev.callback.n == 0
    cb = ev.callback
    for _ in range(3):
        cb(1, 2)
    assert ev.callback.n == 3
    ev.callback.n = 0
    assert ev.callback.n == 0 and id(lcb) != id(ev.callback)
    ev = t_event_batch(x + v, direction=event_direction.positive, cooldown=fp_t(3), callback=None)
    assert ev.callback is None
    with pytest.raises(ValueError) as cm:
        t_event_batch(x + v, direction=event_direction(45), cooldown=fp_t(3), callback=lambda _: _)
    assert "45 is not a valid event_direction" in str(cm.value)
    with pytest.raises(TypeError) as cm:
        t_event_batch(x + v, callback=3)
    assert ("An object of type '{}' cannot be used as an event callback because it is not callable".format(
        str(type(3)))) in str(cm.value)

    ev = t_event_batch(x + v, direction=event_direction.positive, cooldown=fp_t(3), callback=lambda _: _)
    ev = pickle.loads(pickle.dumps(ev))
    r = repr(ev)
    assert " terminal" in r and "(x + v)" in r and "event_direction::positive" in r and ": yes" in r and "3" in r
    ev.foo = "hello world"
    ev = pickle.loads(pickle.dumps(ev))
    assert ev.foo == "hello world"

    class foo:
        pass

    ev.bar = foo()
    assert id(ev.bar) == id(copy(ev).bar) and id(ev.bar) != id(deepcopy(ev).bar)
    ev = t_event_batch(x + v, direction=event_direction.positive, cooldown=fp_t(3))
    ev = pickle.loads(pickle.dumps(ev))
    r = repr(ev)
    assert " terminal" in r and "event_direction::positive" in r and ": no" in r and "3" in r
    ev = t_event_batch(ex=x + v, callback=local_cb(), direction=event_direction.negative)
    out_cb = ev.callback
    del ev
    gc.collect()
    for _ in range(3):
        out_cb(1, 2)
    assert out_cb.n == 3
    with pytest.raises(TypeError) as cm:
        t_event_batch(x + v, fp_type=list)
    assert 'The floating-point type "{}" is not recognized/supported'.format(list) in str(cm.value)
