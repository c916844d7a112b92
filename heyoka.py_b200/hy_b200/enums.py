"""Enums of the hot path, value-compatible with the reference
(/root/reference/heyoka/core.cpp:324-336)."""

import enum


class taylor_outcome(enum.IntEnum):
    # Values pinned by /root/reference/doc/notebooks/Batch mode overview.ipynb:242,378
    # (success = -4294967297, time_limit = -4294967299).
    success = -4294967297
    step_limit = -4294967298
    time_limit = -4294967299
    err_nf_state = -4294967300
    cb_stop = -4294967301


class event_direction(enum.IntEnum):
    negative = -1
    any = 0
    positive = 1


class code_model(enum.IntEnum):
    tiny = 0
    small = 1
    kernel = 2
    medium = 3
    large = 4


class _term_outcome(int):
    """Outcome produced by a terminal event: ``idx`` (continuing) or
    ``-idx-1`` (stopping); prints like the reference's
    ``taylor_outcome.terminal_event_N``."""

    def __repr__(self):
        v = int(self)
        if v >= 0:
            return "<taylor_outcome.terminal_event_{} (continuing): {}>".format(v, v)
        return "<taylor_outcome.terminal_event_{} (stopping): {}>".format(-v - 1, v)


def _outcome_from_int(v):
    try:
        return taylor_outcome(v)
    except ValueError:
        return _term_outcome(v)
