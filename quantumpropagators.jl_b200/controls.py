"""Host-side control discretisation (mirror of the reference's ``Controls`` module,
``src/controls.jl``).  PWC propagators consume one number per control per interval; these
helpers turn control functions / vectors into those numbers.  Interval and grid indices
``n`` are 1-based, as in the reference.
"""

from __future__ import annotations

import numpy as np

__all__ = [
    "IdDict",
    "discretize",
    "discretize_on_midpoints",
    "get_tlist_midpoints",
    "t_mid",
    "evaluate",
    "get_controls",
    "substitute",
]


class IdDict:
    """Mapping keyed by object identity (Julia ``IdDict``): controls are functions or arrays,
    neither of which is hashable by value."""

    def __init__(self, pairs=()):
        self._items = {}
        if isinstance(pairs, IdDict):
            pairs = pairs.items()
        elif isinstance(pairs, dict):
            pairs = pairs.items()
        for key, value in pairs:
            self[key] = value

    def __setitem__(self, key, value):
        self._items[id(key)] = (key, value)

    def __getitem__(self, key):
        try:
            return self._items[id(key)][1]
        except KeyError:
            raise KeyError(key) from None

    def __contains__(self, key):
        return id(key) in self._items

    def __len__(self):
        return len(self._items)

    def __iter__(self):
        return iter(self.keys())

    def get(self, key, default=None):
        item = self._items.get(id(key))
        return default if item is None else item[1]

    def keys(self):
        return [k for k, _ in self._items.values()]

    def values(self):
        return [v for _, v in self._items.values()]

    def items(self):
        return list(self._items.values())


def _is_vector(obj) -> bool:
    return isinstance(obj, (np.ndarray, list))


def get_controls(ampl):
    """Controls an amplitude depends on: a function or vector is its own control, a number
    has none (reference ``src/controls.jl:219-258``)."""
    if getattr(ampl, "is_amplitude", False):  # amplitude objects (amplitudes.py) know their controls
        return tuple(ampl.get_controls())
    if callable(ampl) or _is_vector(ampl):
        return (ampl,)
    return ()


def get_tlist_midpoints(tlist, preserve_start=True, preserve_end=True) -> np.ndarray:
    """Interval midpoints of ``tlist`` with the first / last value snapped to the grid ends
    (reference ``src/controls.jl:92-124``; known answers in ``test/test_discretization.jl``)."""
    t = np.asarray(tlist, dtype=np.float64)
    if t.size < 3:
        raise ValueError("In `get_tlist_midpoints`, argument `tlist` must have a length of at least 3")
    steps = np.diff(t)
    if not np.all(steps > 0.0):
        raise AssertionError("tlist must be strictly increasing")
    mid = t[:-1] + 0.5 * steps
    if preserve_start:
        mid[0] = t[0]
    if preserve_end:
        mid[-1] = t[-1]
    return mid


def discretize_on_midpoints(control, tlist) -> np.ndarray:
    """Values of ``control`` on the nt-1 intervals (reference ``src/controls.jl:189-208``).
    A vector on the nt grid points is un-averaged: p_1 = c_1, p_i = 2 c_i - p_{i-1}."""
    nt = len(tlist)
    if callable(control):
        return np.array([control(t) for t in get_tlist_midpoints(tlist)], dtype=np.float64)
    c = np.asarray(control)
    if c.shape[0] == nt - 1:
        return np.array(c, dtype=np.float64)
    if c.shape[0] == nt:
        p = np.empty(nt - 1, dtype=np.float64)
        p[0] = c[0]
        p[-1] = c[-1]
        for i in range(1, nt - 2):
            p[i] = 2.0 * c[i] - p[i - 1]
        return p
    raise ValueError("control array must be defined on the points of tlist")


def discretize(control, tlist, via_midpoints=True) -> np.ndarray:
    """Values of ``control`` on the nt grid points (reference ``src/controls.jl:43-68``)."""
    nt = len(tlist)
    if callable(control):
        if not via_midpoints:
            return np.array([control(t) for t in tlist], dtype=np.float64)
        control = discretize_on_midpoints(control, tlist)
    c = np.asarray(control)
    if c.shape[0] == nt:
        return np.array(c, dtype=np.float64)
    if c.shape[0] == nt - 1:
        v = np.empty(nt, dtype=np.float64)
        v[0] = c[0]
        v[-1] = c[-1]
        v[1:-1] = 0.5 * (c[:-1] + c[1:])
        return v
    raise ValueError("control array must be defined on intervals of tlist")


def t_mid(tlist, n: int) -> float:
    """Midpoint of the n-th interval, snapping at both ends (reference ``src/controls.jl:332-343``)."""
    nt = len(tlist)
    if not 1 <= n <= nt - 1:
        raise AssertionError("n must be an interval of tlist")
    if n == 1:
        return float(tlist[0])
    if n == nt - 1:
        return float(tlist[-1])
    return float(tlist[n - 1] + 0.5 * (tlist[n] - tlist[n - 1]))


def evaluate(obj, *args, vals_dict=None):
    """``evaluate(control, tlist, n; vals_dict)`` / ``evaluate(control, t; vals_dict)``
    (reference ``src/controls.jl:302-306, 346-397``)."""
    if vals_dict is not None and obj in vals_dict:
        return vals_dict[obj]
    if getattr(obj, "is_amplitude", False):
        return obj.evaluate(*args, vals_dict=vals_dict)
    if callable(obj):
        if len(args) == 2:
            return obj(t_mid(args[0], args[1]))
        if len(args) == 1:
            return obj(args[0])
        raise TypeError("evaluate(control, tlist, n) or evaluate(control, t)")
    if _is_vector(obj):
        if len(args) != 2:
            raise ValueError("`evaluate(control::Vector, t::Float64)` is invalid. Use e.g. `evaluate(…, tlist, n)`.")
        tlist, n = args
        nt = len(tlist)
        if len(obj) == nt - 1:
            return obj[n - 1]
        if len(obj) == nt:
            if n == 1 or n == nt:
                return obj[n - 1]
            return 2 * obj[n - 1] - obj[n - 2]
        raise ValueError(
            f"control (length {len(obj)}) must be discretized either on `tlist` (length {nt}) "
            "or on the midpoints of `tlist`"
        )
    return obj


def substitute(obj, replacements):
    """``substitute(object, replacements)`` for controls and amplitudes (reference
    ``src/controls.jl:476-500``): an object that is a key of ``replacements`` (by identity) is
    replaced, anything else is returned unchanged."""
    if isinstance(replacements, dict):
        replacements = IdDict(replacements)
    if getattr(obj, "is_amplitude", False):
        return obj.substitute(replacements)
    return replacements.get(obj, obj)
