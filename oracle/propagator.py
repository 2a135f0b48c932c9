"""Oracle restatement of the piecewise-constant propagator protocol for the Cheby and Newton
methods: ``src/propagator.jl``, ``src/pwc_utils.jl``, ``src/cheby_propagator.jl``,
``src/newton_propagator.jl`` and the step loop of ``src/propagate.jl``.

Test infrastructure only.  Interval / grid indices ``n`` are 1-based like the reference.
"""

from __future__ import annotations

import warnings

import numpy as np

from .cheby import ChebyWrk, cheby_inplace, cheby
from .newton import NewtonWrk, newton_inplace
from .specrad import specrange
from .controls import IdDict, discretize, discretize_on_midpoints
from .generators import (
    Generator,
    Operator,
    evaluate,
    evaluate_inplace,
    get_controls,
    hamiltonian,
)


def _get_uniform_dt(tlist, tol=1e-12, warn=False):
    """``_get_uniform_dt`` -- ``src/propagator.jl:267-280``."""
    dt = float(tlist[1] - tlist[0])
    for i in range(1, len(tlist) - 1):
        dt_i = tlist[i + 1] - tlist[i]
        if abs(dt_i - dt) > tol:
            if warn:
                warnings.warn(f"Non-uniform time grid: dt = {dt_i:.2e} in interval {i+1}")
            return None
    return dt


def _pwc_process_parameters(parameters, controls, tlist):
    """``_pwc_process_parameters`` -- ``src/pwc_utils.jl:29-45``."""
    if parameters is None:
        return IdDict((c, discretize_on_midpoints(c, tlist)) for c in controls)
    for c in controls:
        amplitude = parameters[c]
        assert len(amplitude) == len(tlist) - 1
    return parameters


def _pwc_get_max_genop(generator, controls, tlist):
    """``_pwc_get_max_genop`` -- ``src/pwc_utils.jl:74-83``."""
    controlvals = [discretize(c, tlist) for c in controls]
    n = len(tlist) // 2
    max_vals = IdDict((c, np.max(controlvals[i])) for i, c in enumerate(controls))
    return evaluate(generator, tlist, n, vals_dict=max_vals)


def cheby_get_spectral_envelope(generator, tlist, control_ranges, method, **kwargs):
    """``cheby_get_spectral_envelope`` -- ``src/cheby_propagator.jl:331-345``."""
    min_vals = IdDict((c, r[0]) for c, r in control_ranges.items())
    n = len(tlist) // 2
    G_min = evaluate(generator, tlist, n, vals_dict=min_vals)
    max_vals = IdDict((c, r[1]) for c, r in control_ranges.items())
    G_max = evaluate(generator, tlist, n, vals_dict=max_vals)
    E_min, E_max = specrange(G_max, method, **kwargs)
    _E_min, _E_max = specrange(G_min, method, **kwargs)
    E_min = _E_min if _E_min < E_min else E_min
    E_max = _E_max if _E_max > E_max else E_max
    return E_min, E_max


class _PWCPropagator:
    def _set_t(self, t):
        """``_pwc_set_t!`` -- ``src/pwc_utils.jl:48-71``."""
        tlist = self.tlist
        if t <= tlist[0]:
            n = 1
        else:
            N = len(tlist)
            if t >= tlist[-1]:
                n = N
            else:
                n = min(int(np.searchsorted(tlist, t, side="left")) + 1, N)
        if not np.isclose(t, tlist[n - 1], rtol=1.5e-8, atol=0.0):
            warnings.warn(f"Snapping t={t} to time grid value {tlist[n-1]}")
        self.n = n - 1 if self.backward else n
        self.t = float(tlist[n - 1])

    def _vals_dict(self, n):
        return IdDict((c, self.parameters[c][n - 1]) for c in self.controls)

    def _set_genop(self, n):
        """``_pwc_set_genop!`` -- ``src/pwc_utils.jl:86-92``."""
        if isinstance(self.genop, Operator) and isinstance(self.generator, Generator):
            evaluate_inplace(self.genop, self.generator, self.tlist, n, vals_dict=self._vals_dict(n))
        else:
            self.genop = evaluate(self.generator, self.tlist, n, vals_dict=self._vals_dict(n))
        return self.genop

    def _advance_time(self):
        """``_pwc_advance_time!`` -- ``src/pwc_utils.jl:102-112``."""
        n = self.n
        if self.backward:
            self.t = float(self.tlist[n - 1])
            self.n = n - 1
        else:
            self.t = float(self.tlist[n])
            self.n = n + 1


class ChebyPropagator(_PWCPropagator):
    """``ChebyPropagator`` -- ``src/cheby_propagator.jl:9-27``."""


class NewtonPropagator(_PWCPropagator):
    """``NewtonPropagator`` -- ``src/newton_propagator.jl:9-26``."""


def _canonical_generator(generator):
    # tuple generators (H0, (H1, eps)) -> Generator / Operator / matrix (src/generators.jl:729-733)
    if isinstance(generator, (tuple, list)):
        return hamiltonian(*generator, check=False)
    return generator


def init_prop(
    state,
    generator,
    tlist,
    method,
    inplace=True,
    backward=False,
    parameters=None,
    **kwargs,
):
    """``init_prop(state, generator, tlist; method, ...)`` -- ``src/propagator.jl:208-264``
    dispatching to ``src/cheby_propagator.jl:87-175`` / ``src/newton_propagator.jl:62-113``."""
    name = str(method).lower().lstrip(":")
    tlist = np.asarray(tlist, dtype=np.float64)
    generator = _canonical_generator(generator)
    controls = get_controls(generator)
    if name == "cheby":
        control_ranges = kwargs.pop("control_ranges", None)
        specrange_method = kwargs.pop("specrange_method", "auto")
        specrange_buffer = kwargs.pop("specrange_buffer", 0.01)
        cheby_coeffs_limit = kwargs.pop("cheby_coeffs_limit", 1e-12)
        check_normalization = kwargs.pop("check_normalization", False)
        uniform_dt_tolerance = kwargs.pop("uniform_dt_tolerance", 1e-12)
        specrange_kwargs = kwargs
        controlvals = [discretize(c, tlist) for c in controls]
        G = _pwc_get_max_genop(generator, controls, tlist)
        parameters = _pwc_process_parameters(parameters, controls, tlist)
        if control_ranges is None:
            control_ranges = IdDict(
                (c, (np.min(controlvals[i]), np.max(controlvals[i])))
                for i, c in enumerate(controls)
            )
        else:
            for c in controls:
                assert c in control_ranges
                assert control_ranges[c][0] <= control_ranges[c][1]
        E_min, E_max = cheby_get_spectral_envelope(
            generator, tlist, control_ranges, specrange_method, **specrange_kwargs
        )
        Delta = E_max - E_min
        assert Delta > 0.0
        delta = specrange_buffer * Delta
        E_min = E_min - delta / 2
        Delta = Delta + delta
        dt = _get_uniform_dt(tlist, tol=uniform_dt_tolerance, warn=True)
        if dt is None:
            raise RuntimeError("Chebychev propagation only works on a uniform time grid")
        p = ChebyPropagator()
        p.wrk = ChebyWrk(state, Delta, E_min, dt, limit=cheby_coeffs_limit)
        p.control_ranges = control_ranges
        p.specrange_method = specrange_method
        p.specrange_buffer = specrange_buffer
        p.check_normalization = check_normalization
        p.specrange_options = specrange_kwargs
    elif name == "newton":
        if not inplace:
            raise RuntimeError("The Newton propagator is only implemented in-place")
        G = _pwc_get_max_genop(generator, controls, tlist)
        parameters = _pwc_process_parameters(parameters, controls, tlist)
        p = NewtonPropagator()
        p.wrk = NewtonWrk(state, m_max=kwargs.get("m_max", 10))
        p.func = kwargs.get("func", None)
        p.norm_min = kwargs.get("norm_min", 1e-14)
        p.relerr = kwargs.get("relerr", 1e-12)
        p.max_restarts = kwargs.get("max_restarts", 50)
    else:
        raise ValueError(f"Unknown propagation `method`: {method}")
    p.generator = generator
    p.state = state.copy() if inplace else state
    p.tlist = tlist
    p.parameters = parameters
    p.controls = controls
    p.genop = G
    p.backward = backward
    p.inplace = inplace
    p.n = len(tlist) - 1 if backward else 1
    p.t = float(tlist[-1]) if backward else float(tlist[0])
    return p


def prop_step(p):
    """``prop_step!`` -- ``src/cheby_propagator.jl:348-386`` / ``src/newton_propagator.jl:120-153``.
    Returns the state, or ``None`` beyond the time grid."""
    n = p.n
    tlist = p.tlist
    if not (0 < n < len(tlist)):
        return None
    if isinstance(p, ChebyPropagator):
        dt = -p.wrk.dt if p.backward else p.wrk.dt
        H = p._set_genop(n)
        if p.inplace:
            cheby_inplace(p.state, H, dt, p.wrk, check_normalization=p.check_normalization)
        else:
            p.state = cheby(p.state, H, dt, p.wrk, check_normalization=p.check_normalization)
    else:
        dt = float(tlist[n] - tlist[n - 1])
        if p.backward:
            dt = -dt
        H = p._set_genop(n)
        newton_inplace(
            p.state,
            H,
            dt,
            p.wrk,
            func=p.func,
            norm_min=p.norm_min,
            relerr=p.relerr,
            max_restarts=p.max_restarts,
        )
    p._advance_time()
    return p.state


def set_state(p, state):
    """``set_state!`` -- ``src/propagator.jl:367-377``."""
    if state is not p.state:
        if p.inplace:
            p.state[...] = state
        else:
            p.state = np.asarray(state, dtype=p.state.dtype)
    return p.state


def set_t(p, t):
    """``set_t!`` for PWC propagators -- ``src/cheby_propagator.jl:30`` → ``_pwc_set_t!``."""
    p._set_t(t)


def reinit_prop(p, state, transform_control_ranges=None, **_):
    """``reinit_prop!`` -- ``src/cheby_propagator.jl:243-299`` (Cheby: may recompute the
    coefficients) / ``src/propagator.jl:298-312`` (default)."""
    state = set_state(p, state)
    tlist = p.tlist
    if isinstance(p, ChebyPropagator):
        if transform_control_ranges is None:
            transform_control_ranges = lambda c, lo, hi, check: (lo, hi)  # noqa: E731
        wrk = p.wrk
        need = False
        control_ranges = IdDict(
            (c, (np.min(p.parameters[c]), np.max(p.parameters[c]))) for c in p.controls
        )
        for c in p.controls:
            lo, hi = control_ranges[c]
            lo_c, hi_c = transform_control_ranges(c, lo, hi, True)
            if (lo_c < p.control_ranges[c][0]) or (hi_c > p.control_ranges[c][1]):
                need = True
                break
        if need:
            for c in p.controls:
                lo, hi = control_ranges[c]
                control_ranges[c] = transform_control_ranges(c, lo, hi, False)
            E_min, E_max = cheby_get_spectral_envelope(
                p.generator, tlist, control_ranges, p.specrange_method, **p.specrange_options
            )
            Delta = E_max - E_min
            assert Delta > 0.0
            delta = p.specrange_buffer * Delta
            E_min = E_min - delta / 2
            Delta = Delta + delta
            dt = float(tlist[1] - tlist[0])
            p.control_ranges = control_ranges
            wrk = ChebyWrk(state, Delta, E_min, dt, limit=wrk.limit)
        p.wrk = wrk
    p._set_t(float(tlist[-1] if p.backward else tlist[0]))


def propagate(
    state,
    generator,
    tlist,
    method,
    storage=None,
    observables=None,
    callback=None,
    **kwargs,
):
    """``propagate(state, generator, tlist; method, ...)`` -- ``src/propagate.jl:167-235,
    283-344``.  ``storage=True`` returns an array with one column per time point holding
    the state (default) or the tuple of ``observables(state)`` values."""
    p = init_prop(state, generator, tlist, method, **kwargs)
    return propagate_propagator(p, storage=storage, observables=observables, callback=callback)


def _observe(observables, state):
    if observables is None:
        return state.copy()
    return np.array([obs(state) for obs in observables])


def propagate_propagator(p, storage=None, observables=None, callback=None):
    tlist = p.tlist
    nt = len(tlist)
    return_storage = False
    if storage is True:
        first = _observe(observables, p.state)
        storage = np.zeros(first.shape + (nt,), dtype=first.dtype)
        return_storage = True
    if storage is not None:
        storage[..., nt - 1 if p.backward else 0] = _observe(observables, p.state)
    order = range(nt - 1, 0, -1) if p.backward else range(1, nt)
    for i in order:  # i = 1-based interval index
        prop_step(p)
        if callback is not None:
            callback(p, observables)
        if storage is not None:
            slot = (i - 1) if p.backward else i  # 0-based column: i+(backward?0:1) - 1
            storage[..., slot] = _observe(observables, p.state)
    return storage if return_storage else p.state
