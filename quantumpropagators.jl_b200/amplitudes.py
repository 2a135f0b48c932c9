"""Control amplitudes (mirror of the reference's ``Amplitudes`` module, ``src/amplitudes.jl``):
``LockedAmplitude`` (a fixed shape, no control), ``ShapedAmplitude`` (a(t) = S(t) ε(t)) and
``GuidedAmplitude`` (a(t) = G(t) + S(t) ε(t)).  Host objects: on a time interval they evaluate to
one number, which is what reaches the device as an operator coefficient.  ``shape`` / ``guide`` /
``control`` are callables of t or vectors of values on the midpoints of the time grid."""

from __future__ import annotations

import numpy as np

from . import controls as _controls
from .controls import IdDict, discretize_on_midpoints, t_mid

__all__ = ["LockedAmplitude", "ShapedAmplitude", "GuidedAmplitude"]


def _as_part(x, what, check=True):
    if isinstance(x, (list, tuple, np.ndarray)):
        try:
            return np.array(x, dtype=np.float64)
        except (TypeError, ValueError):
            raise ValueError(f"A {what} that is a vector must be convertible to Vector{{Float64}}") from None
    if check:
        try:
            x(0.0)
        except Exception:  # noqa: BLE001 -- the reference reports any failure the same way
            raise ValueError(f"A {what} must either be a Vector{{Float64}} or a callable") from None
    return x


def _is_vec(x) -> bool:
    return isinstance(x, np.ndarray)


def _value(part, tlist, n):
    return part[n - 1] if _is_vec(part) else part(t_mid(tlist, n))


class LockedAmplitude:
    """``LockedAmplitude(shape)`` / ``LockedAmplitude(shape, tlist)`` (reference
    ``src/amplitudes.jl:20-75``): an amplitude without a control."""

    is_amplitude = True

    def __init__(self, shape, tlist=None, check=True):
        if tlist is not None:
            shape, check = discretize_on_midpoints(shape, tlist), False
        self.shape = _as_part(shape, "LockedAmplitude shape", check)

    def get_controls(self):
        return ()

    def substitute(self, replacements):
        return replacements.get(self, self)

    def evaluate(self, *args, vals_dict=None):
        if len(args) == 2:
            return _value(self.shape, *args)
        if _is_vec(self.shape):
            raise ValueError("A LockedAmplitude initialized from a vector can only be evaluated with (tlist, n).")
        return self.shape(args[0])

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.shape, dtype=dtype)

    def __repr__(self):
        return f"LockedAmplitude(::{type(self.shape).__name__})"


class ShapedAmplitude:
    """``ShapedAmplitude(control; shape)`` / ``ShapedAmplitude(control, tlist; shape)`` (reference
    ``src/amplitudes.jl:100-215``): a(t) = S(t) ε(t); callable when both parts are."""

    is_amplitude = True

    def __init__(self, control, tlist=None, *, shape, check=True):
        if tlist is not None:
            control, shape, check = discretize_on_midpoints(control, tlist), discretize_on_midpoints(shape, tlist), False
        self.control = _as_part(control, "ShapedAmplitude control", check)
        self.shape = _as_part(shape, "ShapedAmplitude shape", check)
        if check and _is_vec(self.control) and _is_vec(self.shape) and len(self.control) != len(self.shape):
            raise ValueError("ShapedAmplitude control and shape vectors must have the same length")

    def get_controls(self):
        return (self.control,)

    def substitute(self, replacements):
        if self in replacements:
            return replacements[self]
        return ShapedAmplitude(_controls.substitute(self.control, replacements), shape=self.shape)

    def _extra(self, *args):
        return 0.0

    def evaluate(self, *args, vals_dict=None):
        eps = _controls.evaluate(self.control, *args, vals_dict=vals_dict)
        if len(args) == 2:
            return self._extra(*args) + _value(self.shape, *args) * eps
        if _is_vec(self.shape):
            raise ValueError(f"A {type(self).__name__} with a vector shape can only be evaluated with (tlist, n).")
        return self._extra(*args) + self.shape(args[0]) * eps

    def __call__(self, t):
        return self.evaluate(float(t))

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.control * self.shape, dtype=dtype)

    def __repr__(self):
        return f"{type(self).__name__}(::{type(self.control).__name__}; shape::{type(self.shape).__name__})"


class GuidedAmplitude(ShapedAmplitude):
    """``GuidedAmplitude(control; guide, shape)`` (reference ``src/amplitudes.jl``):
    a(t) = G(t) + S(t) ε(t)."""

    def __init__(self, control, tlist=None, *, guide, shape=None, check=True):
        if shape is None:
            shape = lambda t: 1.0  # noqa: E731
        super().__init__(control, tlist, shape=shape, check=check)
        if tlist is not None:
            guide, check = discretize_on_midpoints(guide, tlist), False
        self.guide = _as_part(guide, "GuidedAmplitude guide", check)

    def substitute(self, replacements):
        if self in replacements:
            return replacements[self]
        return GuidedAmplitude(_controls.substitute(self.control, replacements), guide=self.guide, shape=self.shape)

    def _extra(self, *args):
        if len(args) == 2:
            return _value(self.guide, *args)
        if _is_vec(self.guide):
            raise ValueError("A GuidedAmplitude with a vector guide can only be evaluated with (tlist, n).")
        return self.guide(args[0])

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.guide + self.control * self.shape, dtype=dtype)
