"""Interface checks (mirror of the reference's ``QuantumPropagators.Interfaces`` module,
``src/interfaces/*.jl``): ``check_state``, ``check_operator``, ``check_generator``,
``check_propagator``, ``check_tlist``, ``check_control``, ``check_amplitude`` and the traits
``supports_inplace`` / ``supports_vector_interface`` / ``supports_matrix_interface``.

Same contract as the reference: every ``check_*`` returns ``True`` / ``False`` and, unless
``quiet=True``, logs one error per violated requirement (logger
``qprop_b200.interfaces``).  They accept both host objects (NumPy vectors, NumPy / SciPy
matrices) and the device-resident types of this package (``DeviceState``, ``Operator``,
``ScaledOperator``, ``Generator``, the propagators); a ``DeviceState`` is *not* a vector
(``supports_vector_interface`` is ``False``, like any non-``AbstractVector`` state in the
reference, ``src/interfaces/supports_vector_interface.jl:24-28``), so only the Hilbert-space
verbs are required of it.
"""

from __future__ import annotations

import logging
import math

import numpy as np
import scipy.sparse as sp

from . import controls as _controls
from . import generators as _generators
from .controls import IdDict
from .device import DeviceState
from .generators import Generator, Operator, ScaledOperator

__all__ = [
    "supports_inplace",
    "supports_vector_interface",
    "supports_matrix_interface",
    "check_tlist",
    "check_control",
    "check_amplitude",
    "check_state",
    "check_operator",
    "check_generator",
    "check_propagator",
]

log = logging.getLogger("qprop_b200.interfaces")


# ---------------------------------------------------------------------------------------
# traits (src/interfaces/supports_inplace.jl, supports_vector_interface.jl,
# supports_matrix_interface.jl)
# ---------------------------------------------------------------------------------------


def supports_inplace(obj) -> bool:
    """Whether ``obj`` can be mutated in place (reference
    ``src/interfaces/supports_inplace.jl:1-61``).  Undefined for unknown types: raises
    ``TypeError`` like the reference's fallback method throws."""
    if isinstance(obj, DeviceState):
        return True
    if isinstance(obj, np.ndarray):
        return bool(obj.flags.writeable)
    if sp.issparse(obj):
        return True
    if isinstance(obj, Operator):
        return True
    if isinstance(obj, ScaledOperator):
        return supports_inplace(obj.operator)
    if isinstance(obj, (tuple, int, float, complex, np.number)):
        return False
    raise TypeError(f"`supports_inplace` is not defined for type {type(obj).__name__}")


def supports_vector_interface(state) -> bool:
    """``True`` only for one-dimensional host arrays (reference
    ``src/interfaces/supports_vector_interface.jl:1-28``)."""
    return isinstance(state, np.ndarray) and state.ndim == 1


def supports_matrix_interface(op) -> bool:
    """``True`` only for two-dimensional host arrays / sparse matrices (reference
    ``src/interfaces/supports_matrix_interface.jl:1-36``)."""
    return (isinstance(op, np.ndarray) and op.ndim == 2) or sp.issparse(op)


# ---------------------------------------------------------------------------------------
# verbs that work on host vectors and on DeviceState alike
# ---------------------------------------------------------------------------------------


def _dot(a, b):
    return a.dot(b) if isinstance(a, DeviceState) else np.vdot(a, b)


def _norm(a):
    return a.norm() if isinstance(a, DeviceState) else float(np.linalg.norm(a))


def _copy(a):
    return a.copy()


def _zero(a):
    return a.zero() if isinstance(a, DeviceState) else np.zeros_like(a)


def _similar(a):
    return a.similar() if isinstance(a, DeviceState) else np.empty_like(a)


def _copyto(dst, src):
    if isinstance(dst, DeviceState):
        return dst.copyto(src)
    np.copyto(dst, src)
    return dst


def _fill(a, c):
    if isinstance(a, DeviceState):
        return a.fill(c)
    a[...] = c
    return a


def _lmul(c, a):
    if isinstance(a, DeviceState):
        return a.lmul(c)
    a *= c
    return a


def _axpy(c, x, y):
    if isinstance(y, DeviceState):
        return y.axpy(c, x)
    y += c * x
    return y


def _is_complex_scalar(v) -> bool:
    return isinstance(v, (complex, np.complexfloating))


def _is_number(v) -> bool:
    return isinstance(v, (int, float, complex, np.number)) and not isinstance(v, bool)


def _apply(op, state):
    """``op * state``."""
    if isinstance(op, (Operator, ScaledOperator)):
        return op @ state
    if isinstance(state, DeviceState):
        return _generators._as_operator(op) @ state
    return op @ state


def _mul(y, op, x, alpha=1.0, beta=0.0):
    """``mul!(y, op, x, α, β)``; returns the object it wrote into."""
    if isinstance(op, (Operator, ScaledOperator)):
        return op.mul(y, x, alpha, beta)
    if isinstance(x, DeviceState):
        return _generators._as_operator(op).mul(y, x, alpha, beta)
    y[...] = beta * y + alpha * (op @ x) if beta != 0 else alpha * (op @ x)
    return y


def _dot3(x, op, y):
    """``dot(x, op, y)``."""
    if isinstance(op, (Operator, ScaledOperator)):
        return op.dot(x, y)
    if isinstance(x, DeviceState):
        return _generators._as_operator(op).dot(x, y)
    return np.vdot(x, op @ y)


class _Report:
    def __init__(self, quiet, prefix):
        self.quiet, self.prefix, self.success = quiet, prefix, True

    def fail(self, msg, exc=None):
        self.success = False
        if not self.quiet:
            log.error("%s%s%s", self.prefix, msg, f" ({type(exc).__name__}: {exc})" if exc is not None else "")


# ---------------------------------------------------------------------------------------
# check_tlist / check_control / check_amplitude
# ---------------------------------------------------------------------------------------


def check_tlist(tlist, quiet=False, _message_prefix="") -> bool:
    """A valid time grid is a float64 vector of at least two monotonically increasing points
    (reference ``src/interfaces/tlist.jl:1-51``)."""
    r = _Report(quiet, _message_prefix)
    if not (isinstance(tlist, np.ndarray) and tlist.ndim == 1 and tlist.dtype == np.float64):
        r.fail(f"`tlist` must be a Vector{{Float64}}, not {type(tlist).__name__}")
        return False
    if tlist.size < 2:
        r.fail("`tlist` must contain at least two points")
        return False
    if not np.all(np.diff(tlist) > 0.0):
        r.fail("`tlist` must be monotonically increasing")
    return r.success


def check_control(control, tlist, for_parameterization=False, for_time_continuous=None, quiet=False,
                  _message_prefix="") -> bool:
    """``check_control`` (reference ``src/interfaces/control.jl:1-193``): ``evaluate`` on an
    interval returns a float and honours ``vals_dict``; ``discretize`` /
    ``discretize_on_midpoints`` return finite vectors of length nt / nt-1 (when nt > 2);
    a function control can also be evaluated at a time ``t``.  ``for_parameterization`` is not
    supported by this package (parameterised controls are out of scope)."""
    r = _Report(quiet, _message_prefix)
    if for_time_continuous is None:
        for_time_continuous = callable(control)
    if for_parameterization:
        r.fail("`get_parameters(control)` is not supported by this package")
    tlist = np.asarray(tlist, dtype=np.float64)
    nt = len(tlist)
    try:
        v = _controls.evaluate(control, tlist, 1)
        if not isinstance(v, (float, np.floating)):
            r.fail(f"`evaluate(control, tlist, 1)` must return a Float64, not {type(v).__name__}")
    except Exception as exc:  # noqa: BLE001 -- every failure is a reported violation
        r.fail("`evaluate(control, tlist, n)` must be defined.", exc)
    try:
        marker = 0.123456789
        v = _controls.evaluate(control, tlist, 1, vals_dict=IdDict([(control, marker)]))
        if v != marker:
            r.fail(f"`evaluate(control, tlist, 1; vals_dict=IdDict(control => v))` must return v, not {v!r}")
    except Exception as exc:  # noqa: BLE001
        r.fail("`evaluate(control, tlist, n; vals_dict)` must be defined.", exc)
    if nt > 2:
        for fn, want in ((_controls.discretize, nt), (_controls.discretize_on_midpoints, nt - 1)):
            try:
                vals = fn(control, tlist)
                if not (isinstance(vals, np.ndarray) and vals.dtype == np.float64 and vals.shape == (want,)):
                    r.fail(f"`{fn.__name__}(control, tlist)` must return a vector of {want} floats")
                elif not np.all(np.isfinite(vals)):
                    r.fail(f"all values in `{fn.__name__}(control, tlist)` must be finite")
            except Exception as exc:  # noqa: BLE001
                r.fail(f"`{fn.__name__}(control, tlist)` must be defined.", exc)
    if for_time_continuous:
        t = float(tlist[0])
        try:
            v = _controls.evaluate(control, t)
            if not isinstance(v, (float, np.floating)):
                r.fail(f"`evaluate(control, t)` must return a Float64, not {type(v).__name__}")
        except Exception as exc:  # noqa: BLE001
            r.fail("`evaluate(control, t)` must be defined.", exc)
        try:
            if _controls.evaluate(control, t, vals_dict=IdDict([(control, 0.5)])) != 0.5:
                r.fail("`evaluate(control, t; vals_dict=IdDict(control => v))` must return v")
        except Exception as exc:  # noqa: BLE001
            r.fail("`evaluate(control, t; vals_dict)` must be defined.", exc)
    return r.success


def check_amplitude(ampl, tlist, for_parameterization=False, quiet=False, _message_prefix="") -> bool:
    """``check_amplitude`` (reference ``src/interfaces/amplitude.jl:1-138``): ``get_controls``
    returns a tuple of valid controls, ``substitute`` is defined, ``evaluate`` on an interval
    (with and without ``vals_dict``) returns a number."""
    r = _Report(quiet, _message_prefix)
    tlist = np.asarray(tlist, dtype=np.float64)
    controls = ()
    try:
        controls = _controls.get_controls(ampl)
        if not isinstance(controls, tuple):
            r.fail(f"`get_controls(ampl)` must return a tuple, not {type(controls).__name__}")
            controls = tuple(controls)
        for i, control in enumerate(controls):
            if not check_control(control, tlist, for_parameterization=for_parameterization, quiet=quiet,
                                 _message_prefix=f"{_message_prefix}On control {i + 1} in `ampl`: "):
                r.fail(f"control {i + 1} in `ampl` must pass `check_control`")
    except Exception as exc:  # noqa: BLE001
        r.fail("`get_controls(ampl)` must be defined.", exc)
    try:
        _controls.substitute(ampl, IdDict([(c, c) for c in controls]))
    except Exception as exc:  # noqa: BLE001
        r.fail("`substitute(ampl, replacements)` must be defined.", exc)
    try:
        v = _controls.evaluate(ampl, tlist, 1)
        if not _is_number(v):
            r.fail(f"`evaluate(ampl, tlist, 1)` must return a Number, not {type(v).__name__}")
    except Exception as exc:  # noqa: BLE001
        r.fail("`evaluate(ampl, tlist, n)` must be defined.", exc)
    try:
        vals = IdDict([(c, 1.0) for c in controls])
        v = _controls.evaluate(ampl, tlist, 1, vals_dict=vals)
        if not _is_number(v):
            r.fail(f"`evaluate(ampl, tlist, 1; vals_dict)` must return a Number, not {type(v).__name__}")
    except Exception as exc:  # noqa: BLE001
        r.fail("`evaluate(ampl, tlist, n; vals_dict)` must be defined.", exc)
    return r.success


# ---------------------------------------------------------------------------------------
# check_state
# ---------------------------------------------------------------------------------------


def check_state(state, normalized=False, atol=1e-15, quiet=False, _message_prefix="") -> bool:
    """``check_state`` (reference ``src/interfaces/state.jl:7-602``): the Hilbert-space verbs
    (inner product → Complex, norm induced by it, ``+ - c* copy zero`` and their norm
    properties), the in-place verbs when ``supports_inplace`` (``similar copyto! fill! lmul!
    axpy!``), unit norm if ``normalized``, and the vector interface for host vectors.
    Batched device states (B > 1) are a bundle of B Hilbert-space elements: the scalar
    requirements are checked per trajectory."""
    r = _Report(quiet, _message_prefix)
    tol = max(atol, 1e-14)  # reductions on the device sum in a different order than the host

    def close(a, b):
        return bool(np.all(np.abs(np.asarray(a) - np.asarray(b)) <= tol * max(1.0, float(np.max(np.abs(b))))))

    try:
        inplace = supports_inplace(state)
    except Exception as exc:  # noqa: BLE001
        r.fail(f"The `supports_inplace` method must be defined for type `{type(state).__name__}`.", exc)
        inplace = False
    try:
        c = _dot(state, state)
        cs = np.atleast_1d(c)
        if not np.issubdtype(cs.dtype, np.complexfloating):
            r.fail(f"`dot(state, state)` must return a Complex number type, not {type(c).__name__}")
        n = _norm(state)
        if not close(np.asarray(n) ** 2, cs.real if cs.size > 1 else cs.real[0]):
            r.fail("`norm(state)^2` must match `dot(state, state)`")
    except Exception as exc:  # noqa: BLE001
        r.fail("The inner product (`dot`) and `norm` must be defined.", exc)
    try:
        n = np.asarray(_norm(state))
        two = state + state
        if not np.all(np.asarray(_norm(two)) <= 2 * n + tol):
            r.fail("`norm(state + state)` must fulfill the triangle inequality")
        diff = state - state
        if not close(_norm(diff), 0.0 * n):
            r.fail("`state - state` must have norm 0")
        cp = _copy(state)
        if type(cp) is not type(state):
            r.fail(f"`copy(state)::{type(cp).__name__}` must have the same type as `state::{type(state).__name__}`")
        if cp is state:
            r.fail("`copy(state)` must return a new object")
        if not close(_norm(cp - state), 0.0 * n):
            r.fail("`copy(state) - state` must have norm 0")
        scaled = 0.5 * state
        if not close(_norm(scaled), 0.5 * n):
            r.fail("`norm(state)` must have absolute homogeneity: `norm(s * state) = s * norm(state)`")
        if not close(_norm(0.0 * state), 0.0 * n):
            r.fail("`0.0 * state` must produce a state with norm 0")
        if not close(_norm(_zero(state)), 0.0 * n):
            r.fail("`zero(state)` must produce a state with norm 0")
    except Exception as exc:  # noqa: BLE001
        r.fail("`state + state`, `state - state`, `c * state`, `copy(state)` and `zero(state)` must be defined.", exc)
    if inplace:
        try:
            n = np.asarray(_norm(state))
            other = _similar(state)
            if type(other) is not type(state):
                r.fail("`similar(state)` must return a state of the same type")
            if _copyto(other, state) is not other:
                r.fail("`copyto!(other, state)` must return `other`")
            if not close(_norm(other - state), 0.0 * n):
                r.fail("`copyto!(other, state)` must copy the state")
            _lmul(2.0, other)
            if not close(_norm(other), 2.0 * n):
                r.fail("`lmul!(c, state)` must scale the state")
            _axpy(-2.0, state, other)
            if not close(_norm(other), 0.0 * n):
                r.fail("`axpy!(c, state, other)` must add `c * state` to `other`")
            _fill(other, 0.0)
            if not close(_norm(other), 0.0 * n):
                r.fail("`fill!(state, 0)` must produce a state with norm 0")
        except Exception as exc:  # noqa: BLE001
            r.fail("`similar`, `copyto!`, `fill!`, `lmul!` and `axpy!` must be defined for an in-place state.", exc)
    if normalized:
        try:
            if not close(_norm(state), np.ones_like(np.asarray(_norm(state)))):
                r.fail("`norm(state)` must be 1")
        except Exception as exc:  # noqa: BLE001
            r.fail("`norm(state)` must be defined.", exc)
    try:
        vec = supports_vector_interface(state)
    except Exception as exc:  # noqa: BLE001
        r.fail("`supports_vector_interface(state)` must be defined.", exc)
        vec = False
    if vec:
        try:
            if not np.issubdtype(state.dtype, np.number):
                r.fail("`eltype(state)` must be a numeric type")
            if len(state) != int(np.prod(state.shape)):
                r.fail("`length(state)` must equal `prod(size(state))`")
            if type(state[0]) is not state.dtype.type:
                r.fail("`getindex(state, i)` must return elements matching `eltype`")
            iter(state)
            sim = np.empty_like(state)
            if sim.shape != state.shape or sim.dtype != state.dtype or not sim.flags.writeable:
                r.fail("`similar(state)` must return a mutable vector with the same length and element type")
        except Exception as exc:  # noqa: BLE001
            r.fail("the vector interface must be defined.", exc)
    return r.success


# ---------------------------------------------------------------------------------------
# check_operator
# ---------------------------------------------------------------------------------------


def check_operator(op, state, tlist=None, for_expval=True, atol=1e-14, quiet=False, _message_prefix="") -> bool:
    """``check_operator`` (reference ``src/interfaces/operator.jl:7-442``): ``size``, static
    (evaluates to itself, no controls), ``op * state``, 3- and 5-argument ``mul!`` matching
    ``*`` and returning the target, 3-argument ``dot`` matching ``dot(state, op * state)``,
    and the matrix interface for host matrices."""
    r = _Report(quiet, _message_prefix)
    tlist = np.array([0.0, 1.0]) if tlist is None else np.asarray(tlist, dtype=np.float64)
    assert check_state(state, atol=atol, quiet=True), "`state` must pass `check_state`"
    tol = max(atol, 1e-13)

    def close(a, b):
        a, b = np.asarray(a), np.asarray(b)
        return bool(np.all(np.abs(a - b) <= tol * max(1.0, float(np.max(np.abs(b))))))

    try:
        supports_inplace(op)
    except Exception as exc:  # noqa: BLE001
        r.fail(f"The `supports_inplace` method must be defined for type `{type(op).__name__}`.", exc)
    try:
        s = op.shape
        if not isinstance(s, tuple):
            r.fail(f"`size(op)` must return a tuple, not {type(s).__name__}")
        elif not all(isinstance(d, (int, np.integer)) for d in s):
            r.fail(f"`size(op)` must return a tuple of integers, not {s}")
    except Exception as exc:  # noqa: BLE001
        r.fail("`size(op)` must be defined.", exc)
    try:
        if _generators.evaluate(op, tlist, 1) is not op:
            r.fail("`evaluate(op, tlist, 1) must return operator ≡ op")
    except Exception as exc:  # noqa: BLE001
        r.fail("`evaluate(op, tlist, 1)` must be defined.", exc)
    try:
        controls = _generators.get_controls(op)
        if len(controls) != 0:
            r.fail(f"get_controls(op) must return an empty tuple, not {controls}")
    except Exception as exc:  # noqa: BLE001
        r.fail("`get_controls(op)` must be defined.", exc)
    phi = None
    try:
        phi = _apply(op, state)
        if type(phi) is not type(state):
            r.fail(f"`op * state` must return an object of the same type as `state`, not {type(phi).__name__}")
    except Exception as exc:  # noqa: BLE001
        r.fail("`op * state` must be defined.", exc)
    inplace = False
    try:
        inplace = supports_inplace(state)
    except Exception:  # noqa: BLE001
        pass
    if inplace and phi is not None:
        try:
            out = _similar(state)
            _fill(out, 0.0)
            ret = _mul(out, op, state)
            if ret is not out:
                r.fail("`mul!(ϕ, op, state)` must return the resulting ϕ")
            if not close(_norm(out - phi), 0.0):
                r.fail("`mul!(ϕ, op, state)` must match `op * state`")
        except Exception as exc:  # noqa: BLE001
            r.fail("The 3-argument `mul!` must apply `op` to the given `state`.", exc)
        try:
            out = _copy(state)
            ret = _mul(out, op, state, 0.5, 0.5)
            if ret is not out:
                r.fail("`mul!(ϕ, op, state, α, β)` must return the resulting ϕ")
            expected = 0.5 * state + 0.5 * phi
            if not close(_norm(out - expected), 0.0):
                r.fail("`mul!(ϕ, op, state, α, β)` must match β*ϕ + α*op*state")
        except Exception as exc:  # noqa: BLE001
            r.fail("The 5-argument `mul!` must apply `op` to the given `state`.", exc)
    if for_expval:
        try:
            val = _dot3(state, op, state)
            if not all(_is_number(v) for v in np.atleast_1d(val)):
                r.fail(f"`dot(state, op, state)` must return a number, not {type(val).__name__}")
            elif phi is not None and not close(val, _dot(state, phi)):
                r.fail("`dot(state, op, state)` must match `dot(state, op * state)`")
        except Exception as exc:  # noqa: BLE001
            r.fail("`dot(state, op, state)` must return a number.", exc)
    try:
        mat = supports_matrix_interface(op)
    except Exception as exc:  # noqa: BLE001
        r.fail("`supports_matrix_interface(op)` must be defined.", exc)
        mat = False
    if mat:
        try:
            if not np.issubdtype(op.dtype, np.number):
                r.fail("`eltype(op)` must be a numeric type")
            if len(op.shape) != 2:
                r.fail("`size(op)` must have two dimensions")
            v = op[0, 0]
            if not _is_number(v):
                r.fail("`getindex(op, i, j)` must return a number")
        except Exception as exc:  # noqa: BLE001
            r.fail("the matrix interface must be defined.", exc)
    return r.success


# ---------------------------------------------------------------------------------------
# check_generator
# ---------------------------------------------------------------------------------------


def check_generator(generator, state, tlist, for_expval=True, for_pwc=True, for_time_continuous=False,
                    for_parameterization=False, atol=1e-14, quiet=False, _message_prefix="",
                    _check_amplitudes=True) -> bool:
    """``check_generator`` (reference ``src/interfaces/generator.jl:4-336``): ``get_controls``
    returns a tuple of valid controls, ``substitute`` is defined, the amplitudes of a
    ``Generator`` are valid, and (``for_pwc``) ``evaluate(generator, tlist, n)`` gives a valid
    operator into which ``evaluate!`` can re-evaluate in place."""
    r = _Report(quiet, _message_prefix)
    assert check_state(state, atol=atol, quiet=True), "`state` must pass `check_state`"
    tlist = np.asarray(tlist, dtype=np.float64)
    assert len(tlist) >= 2
    px = _message_prefix
    generator_c = _generators.canonical(generator)
    controls = ()
    try:
        controls = _generators.get_controls(generator)
        if not isinstance(controls, tuple):
            r.fail(f"`get_controls(generator)` must return a tuple, not {type(controls).__name__}")
            controls = tuple(controls)
        for i, control in enumerate(controls):
            if not check_control(control, tlist, for_parameterization=for_parameterization,
                                 for_time_continuous=for_time_continuous and callable(control), quiet=quiet,
                                 _message_prefix=f"{px}On control {i + 1}: "):
                r.fail(f"control {i + 1} must pass `check_control`")
    except Exception as exc:  # noqa: BLE001
        r.fail("`get_controls(generator)` must be defined.", exc)
    try:
        _generators.substitute(generator, IdDict([(c, c) for c in controls]))
    except Exception as exc:  # noqa: BLE001
        r.fail("`substitute(generator, replacements)` must be defined.", exc)
    if isinstance(generator_c, Generator) and _check_amplitudes:
        for i, ampl in enumerate(generator_c.amplitudes):
            if not check_amplitude(ampl, tlist, for_parameterization=for_parameterization, quiet=quiet,
                                   _message_prefix=f"{px}On ampl {i + 1}: "):
                r.fail(f"amplitude {i + 1} must pass `check_amplitude`")
    if for_parameterization:
        r.fail("`get_parameters(generator)` is not supported by this package")

    def _check_evaluated(args, label):
        try:
            op = _generators.evaluate(generator, *args)
            if not check_operator(op, state, tlist=tlist, for_expval=for_expval, atol=atol, quiet=quiet,
                                  _message_prefix=f"{px}On `op = evaluate(generator, {label})`: "):
                r.fail(f"`evaluate(generator, {label})` must return an operator that passes `check_operator`")
            if supports_inplace(op) and isinstance(generator_c, Generator):
                if _generators.evaluate_(op, generator_c, *args) is not op:
                    r.fail(f"`evaluate!(op, generator, {label})` must return `op`")
        except Exception as exc:  # noqa: BLE001
            r.fail(f"`evaluate(generator, {label})` / `evaluate!` must be defined.", exc)

    if for_pwc:
        _check_evaluated((tlist, 1), "tlist, n")
    if for_time_continuous:
        _check_evaluated((float(tlist[0]),), "t")
    return r.success


# ---------------------------------------------------------------------------------------
# check_propagator
# ---------------------------------------------------------------------------------------


def check_propagator(propagator, atol=1e-14, quiet=False, _message_prefix="") -> bool:
    """``check_propagator`` (reference ``src/interfaces/propagator.jl:10-338``) for a freshly
    initialised propagator: required properties, a valid state, ``t`` at the right end of
    the grid, ``prop_step!`` returning the propagator's own state object (in-place) or a new
    one and advancing ``t`` one grid step until it returns ``None`` beyond the grid,
    ``set_t!`` / ``set_state!`` (in place for an in-place propagator, returning the set
    state), per-interval ``parameters``, and an idempotent ``reinit_prop!``."""
    from . import propagator as P  # local import: propagator.py imports nothing from here

    r = _Report(quiet, _message_prefix)
    p = propagator
    for name in ("state", "tlist", "t", "parameters", "backward", "inplace"):
        if not hasattr(p, name):
            r.fail(f"`propagator` does not have the required property `{name}`")
    if not r.success:
        return False
    try:
        if hasattr(p, "generator") and not getattr(type(p), "_exposes_generator", False):
            r.fail("`propagator.generator` must not be accessible")
    except Exception:  # noqa: BLE001 -- raising on access is the expected behaviour
        pass
    try:
        psi0 = p.state.copy()
        if not check_state(p.state, atol=atol, quiet=quiet, _message_prefix=f"{_message_prefix}On `propagator.state`: "):
            r.fail("`propagator.state` must pass `check_state`")
        if p.inplace and not supports_inplace(p.state):
            r.fail("If `propagator.inplace` is true, `supports_inplace(propagator.state)` must be true")
        tlist = np.asarray(p.tlist, dtype=np.float64)
        if not check_tlist(tlist, quiet=quiet, _message_prefix=f"{_message_prefix}On `propagator.tlist`: "):
            r.fail("`propagator.tlist` must be monotonically increasing")
        t_start = tlist[-1] if p.backward else tlist[0]
        if p.t != t_start:
            r.fail(f"`propagator.t` must be the {'last' if p.backward else 'first'} element of `propagator.tlist`")
        if isinstance(p, P.PWCPropagator):
            if not isinstance(p.parameters, (dict, IdDict)):
                r.fail("`propagator.parameters` must be a dict mapping controls to vectors")
            else:
                for control, vals in p.parameters.items():
                    if len(vals) != len(tlist) - 1:
                        r.fail("`propagator.parameters` must hold one value per interval of `propagator.tlist`")
        # stepping
        order = range(len(tlist) - 2, -1, -1) if p.backward else range(1, len(tlist))
        state_obj = p.state
        for k, i in enumerate(order):
            out = P.prop_step(p)
            if out is None:
                r.fail("`prop_step!(propagator)` must return a valid state until the time grid is exhausted")
                break
            if k == 0:
                if p.inplace and out is not state_obj:
                    r.fail("For an in-place propagator, the state returned by `prop_step!` must be the `propagator.state` object")
                if not p.inplace and out is state_obj:
                    r.fail("For a not-in-place propagator, the state returned by `prop_step!` must be a new object")
                if not check_state(out, atol=atol, quiet=True):
                    r.fail("`prop_step!(propagator)` must return a valid state")
            if not math.isclose(p.t, tlist[i], rel_tol=0.0, abs_tol=1e-12):
                r.fail("`prop_step!` must advance `propagator.t` forward or backward one step on the time grid")
                break
        if P.prop_step(p) is not None:
            r.fail("`prop_step!` must return `nothing` when going beyond the time grid")
        # set_t! / set_state!
        t_mid = float(tlist[len(tlist) // 2])
        P.set_t(p, t_mid)
        if p.t != t_mid:
            r.fail("`set_t!(propagator, t)` must set `propagator.t`")
        before = p.state
        ret = P.set_state(p, psi0)
        if ret is not p.state:
            r.fail("`set_state!` must return the set `propagator.state`")
        if p.inplace and p.state is not before:
            r.fail("`set_state!(propagator, state)` for an in-place propagator must overwrite `propagator.state` in-place")
        if float(np.max(np.atleast_1d(_norm(p.state - psi0)))) > max(atol, 1e-14):
            r.fail("`set_state!(propagator, state)` must set `propagator.state`")
        # reinit_prop! (idempotent)
        P.reinit_prop(p, psi0)
        t1 = p.t
        s1 = p.state.copy()
        P.reinit_prop(p, psi0)
        if p.t != t1 or p.t != t_start:
            r.fail("`reinit_prop!` must reset `propagator.t`")
        if float(np.max(np.atleast_1d(_norm(p.state - s1)))) > max(atol, 1e-14):
            r.fail("`reinit_prop!(propagator, state)` must be idempotent")
    except Exception as exc:  # noqa: BLE001
        r.fail("the propagator interface must be fully defined.", exc)
    return r.success
