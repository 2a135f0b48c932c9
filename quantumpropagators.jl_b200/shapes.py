"""Pulse-envelope shapes (mirror of the reference's ``Shapes`` module, ``src/shapes.jl``):
``flattop``, ``box``, ``blackman`` -- the functions control amplitudes are usually built from
(e.g. ``test/test_propagate_sequence.jl``).  Host scalars only: a control reaches the device as
one number per interval."""

from __future__ import annotations

import math

__all__ = ["flattop", "box", "blackman"]


def box(t, t0, T) -> float:
    """Θ-function: 1 for ``t0 <= t <= T``, else 0 (reference ``src/shapes.jl:73``)."""
    return 1.0 if t0 <= t <= T else 0.0


def blackman(t, t0, T, a=0.16) -> float:
    """Blackman window ½(1 − a − cos(2π τ) + a cos(4π τ)), τ = (t − t0)/(T − t0), zero outside
    ``[t0, T]`` (reference ``src/shapes.jl:100-107``)."""
    dT = T - t0
    return 0.5 * box(t, t0, T) * (1.0 - a - math.cos(2 * math.pi * (t - t0) / dT) + a * math.cos(4 * math.pi * (t - t0) / dT))


def flattop(t, *, T, t_rise, t0=0.0, t_fall=None, func="blackman") -> float:
    """Flat shape (amplitude 1) with a switch-on over ``t_rise`` after ``t0`` and a switch-off over
    ``t_fall`` before ``T``; zero outside ``[t0, T]``.  ``func="blackman"`` (half a Blackman
    window) or ``"sinsq"`` (reference ``src/shapes.jl:22-58``)."""
    if t_fall is None:
        t_fall = t_rise
    if func not in ("blackman", "sinsq"):
        raise ValueError(f"Unknown func={func}. Accepted values are blackman and sinsq.")
    f = 0.0
    if t0 <= t <= T:
        f = 1.0
        if t < t0 + t_rise:
            f = blackman(t, t0, t0 + 2 * t_rise) if func == "blackman" else math.sin(math.pi * (t - t0) / (2.0 * t_rise)) ** 2
        elif t > T - t_fall:
            f = blackman(t, T - 2 * t_fall, T) if func == "blackman" else math.sin(math.pi * (t - T) / (2.0 * t_fall)) ** 2
    return f
