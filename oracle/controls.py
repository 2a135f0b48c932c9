"""Oracle restatement of the discretisation part of ``src/controls.jl``.

Test infrastructure only.  A *control* is a Python callable ``ϵ(t)`` or a 1-D array of
values (on ``tlist`` or on its interval midpoints), matching the reference's
``Function`` / ``Vector`` controls.
"""

from __future__ import annotations

import numpy as np


class IdDict:
    """Identity-keyed mapping (Julia's ``IdDict``): arrays/functions as keys."""

    def __init__(self, pairs=()):
        self._d = {}
        for k, v in pairs:
            self[k] = v

    def __setitem__(self, k, v):
        self._d[id(k)] = (k, v)

    def __getitem__(self, k):
        return self._d[id(k)][1]

    def __contains__(self, k):
        return id(k) in self._d

    def get(self, k, default=None):
        return self._d[id(k)][1] if id(k) in self._d else default

    def keys(self):
        return [k for k, _ in self._d.values()]

    def items(self):
        return list(self._d.values())

    def __len__(self):
        return len(self._d)


def _is_vector(x) -> bool:
    return isinstance(x, (np.ndarray, list))


def get_controls_of_amplitude(ampl):
    """``get_controls(ampl)``: a function or vector is its own control; numbers have none
    (``src/controls.jl:219-258``)."""
    if callable(ampl) or _is_vector(ampl):
        return (ampl,)
    return tuple()


def get_tlist_midpoints(tlist, preserve_start=True, preserve_end=True) -> np.ndarray:
    """``get_tlist_midpoints`` -- ``src/controls.jl:92-124``: first/last "midpoints" snap
    to ``tlist[1]`` / ``tlist[end]`` by default."""
    tlist = np.asarray(tlist, dtype=np.float64)
    N = len(tlist)
    if N < 3:
        raise ValueError(
            "In `get_tlist_midpoints`, argument `tlist` must have a length of at least 3"
        )
    mid = np.zeros(N - 1)
    if preserve_start:
        mid[0] = tlist[0]
    else:
        dt = float(tlist[1] - tlist[0])
        assert dt > 0.0
        mid[0] = tlist[0] + 0.5 * dt
    if preserve_end:
        mid[-1] = tlist[-1]
    else:
        dt = float(tlist[-1] - tlist[-2])
        assert dt > 0.0
        mid[-1] = tlist[-2] + 0.5 * dt
    for i in range(1, N - 2):
        dt = float(tlist[i + 1] - tlist[i])
        assert dt > 0.0
        mid[i] = tlist[i] + 0.5 * dt
    return mid


def discretize(control, tlist, via_midpoints=True) -> np.ndarray:
    """``discretize(control, tlist)`` -- ``src/controls.jl:43-68``."""
    if callable(control):
        if via_midpoints:
            return discretize(discretize_on_midpoints(control, tlist), tlist)
        return np.array([control(t) for t in tlist], dtype=np.float64)
    control = np.asarray(control)
    if len(control) == len(tlist):
        return np.array(control, dtype=np.float64)
    if len(control) == len(tlist) - 1:
        vals = np.zeros(len(control) + 1)
        vals[0] = control[0]
        vals[-1] = control[-1]
        for i in range(1, len(vals) - 1):
            vals[i] = 0.5 * (control[i - 1] + control[i])
        return vals
    raise ValueError("control array must be defined on intervals of tlist")


def discretize_on_midpoints(control, tlist) -> np.ndarray:
    """``discretize_on_midpoints`` -- ``src/controls.jl:189-208``."""
    if callable(control):
        return discretize(control, get_tlist_midpoints(tlist), via_midpoints=False)
    control = np.asarray(control)
    if len(control) == len(tlist) - 1:
        return np.array(control, dtype=np.float64)
    if len(control) == len(tlist):
        vals = np.empty(len(tlist) - 1)
        vals[0] = control[0]
        vals[-1] = control[-1]
        for i in range(1, len(vals) - 1):
            vals[i] = 2 * control[i] - vals[i - 1]
        return vals
    raise ValueError("control array must be defined on the points of tlist")


def t_mid(tlist, n: int) -> float:
    """``t_mid(tlist, n)`` -- ``src/controls.jl:332-343``; ``n`` is a 1-based interval."""
    assert 1 <= n <= len(tlist) - 1
    if n == 1:
        return float(tlist[0])
    if n == len(tlist) - 1:
        return float(tlist[-1])
    return float(tlist[n - 1] + (tlist[n] - tlist[n - 1]) / 2)


def evaluate_control(obj, *args, vals_dict=None):
    """``evaluate(control, tlist, n; vals_dict)`` / ``evaluate(control, t; vals_dict)`` --
    ``src/controls.jl:302-306, 346-397``.  ``n`` is 1-based as in the reference."""
    if vals_dict is not None and obj in vals_dict:
        return vals_dict[obj]
    if callable(obj):
        if len(args) == 2:
            tlist, n = args
            return obj(t_mid(tlist, n))
        (t,) = args
        return obj(t)
    if _is_vector(obj):
        if len(args) != 2:
            raise ValueError(
                "`evaluate(control::Vector, t::Float64)` is invalid. Use e.g. `evaluate(…, tlist, n)`."
            )
        tlist, n = args
        if len(obj) == len(tlist) - 1:
            return obj[n - 1]
        if len(obj) == len(tlist):
            if n == 1:
                return obj[0]
            if n == len(tlist):
                return obj[n - 1]
            return 2 * obj[n - 1] - obj[n - 2]
        raise ValueError(
            f"control (length {len(obj)}) must be discretized either on `tlist` "
            f"(length {len(tlist)}) or on the midpoints of `tlist`"
        )
    return obj  # fallback: objects without components evaluate to themselves
