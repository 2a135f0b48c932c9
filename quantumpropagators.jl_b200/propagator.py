"""The propagator protocol: ``init_prop`` / ``prop_step`` / ``set_state`` / ``set_t`` /
``reinit_prop`` / ``propagate`` for ``method="cheby"`` and ``method="newton"``, backed by
device-resident states.

Host mirror of the reference's ``src/propagator.jl``, ``src/pwc_utils.jl``,
``src/cheby_propagator.jl``, ``src/newton_propagator.jl`` and the step loop of
``src/propagate.jl``; Julia's ``!`` functions drop the bang (``prop_step!`` -> ``prop_step``).
Interval indices ``n`` are 1-based like the reference's.  A host (NumPy) initial state is
uploaded once; ``propagator.state`` is always a :class:`DeviceState`, and ``prop_step``
returns that same object (the identity the reference's ``check_propagator`` demands of
in-place propagators, ``src/interfaces/propagator.jl:162-167``).
"""

from __future__ import annotations

import warnings

import numpy as np

from . import controls as _c
from .cheby import ChebyWrk, cheby_, cheby_propagate_
from .controls import IdDict, discretize, discretize_on_midpoints
from .device import Context, DeviceState, default_context
from .generators import Generator, Operator, canonical, evaluate, evaluate_, get_controls, _as_operator
from .newton import NewtonWrk, newton_
from .specrad import specrange

__all__ = [
    "init_prop",
    "prop_step",
    "set_state",
    "set_t",
    "reinit_prop",
    "propagate",
    "ChebyPropagator",
    "NewtonPropagator",
    "PWCPropagator",
]

_PROTECTED = ("generator",)  # hidden like in the reference (src/propagator.jl:80-84)


def _get_uniform_dt(tlist, tol=1e-12, warn=False):
    """reference ``src/propagator.jl:267-280``"""
    dt = float(tlist[1] - tlist[0])
    for i in range(1, len(tlist) - 1):
        dt_i = float(tlist[i + 1] - tlist[i])
        if abs(dt_i - dt) > tol:
            if warn:
                warnings.warn(
                    f"Non-uniform time grid: dt = {dt_i:.2e} in interval {i + 1} differs from the first "
                    f"dt={dt:.2e} by Δ = {abs(dt_i - dt):.2e} > tol = {tol:.2e}"
                )
            return None
    return dt


def _process_parameters(parameters, controls, tlist):
    """reference ``src/pwc_utils.jl:29-45``: control -> nt-1 midpoint values."""
    if parameters is None:
        return IdDict((c, discretize_on_midpoints(c, tlist)) for c in controls)
    for c in controls:
        amplitude = parameters[c]
        if len(amplitude) != len(tlist) - 1:
            raise AssertionError("parameters must hold one value per interval of tlist")
    return parameters


def _max_genop(generator, controls, tlist):
    """reference ``src/pwc_utils.jl:74-83``"""
    vals = IdDict((c, float(np.max(discretize(c, tlist)))) for c in controls)
    return evaluate(generator, tlist, len(tlist) // 2, vals_dict=vals)


def cheby_get_spectral_envelope(generator, tlist, control_ranges, method, ctx=None, **kwargs):
    """Spectral range of the generator over all control values within ``control_ranges``:
    union of the ranges at the all-minimum and all-maximum amplitudes (reference
    ``src/cheby_propagator.jl:331-345``)."""
    n = len(tlist) // 2
    G_min = evaluate(generator, tlist, n, vals_dict=IdDict((c, r[0]) for c, r in control_ranges.items()))
    G_max = evaluate(generator, tlist, n, vals_dict=IdDict((c, r[1]) for c, r in control_ranges.items()))
    E_min, E_max = specrange(G_max, method, ctx=ctx, **kwargs)
    lo, hi = specrange(G_min, method, ctx=ctx, **kwargs)
    return min(lo, E_min), max(hi, E_max)


class PWCPropagator:
    """Common state of the piecewise-constant propagators (reference
    ``src/propagator.jl:48-126``): public properties ``state, tlist, t, parameters, backward,
    inplace``; ``generator`` is hidden."""

    def __getattribute__(self, name):
        if name in _PROTECTED:
            raise AttributeError(f"`{name}` is a private field of the propagator")
        return object.__getattribute__(self, name)

    def propertynames(self):
        return ("state", "tlist", "t", "parameters", "backward", "inplace")

    # -- src/pwc_utils.jl -----------------------------------------------------------------
    def _generator(self):
        return object.__getattribute__(self, "generator")

    def _coeffs_for(self, n):
        """``_pwc_set_genop!`` (reference ``src/pwc_utils.jl:86-92``): evaluate the generator on
        interval n with the *current* ``parameters`` (the caller may mutate them between steps)."""
        vals = IdDict((c, self.parameters[c][n - 1]) for c in self.controls)
        gen = self._generator()
        if isinstance(gen, Generator):
            evaluate_(self.genop, gen, self.tlist, n, vals_dict=vals)
        return self.genop

    def _advance_time(self):
        """reference ``src/pwc_utils.jl:102-112``"""
        n = self.n
        if self.backward:
            object.__setattr__(self, "t", float(self.tlist[n - 1]))
            object.__setattr__(self, "n", n - 1)
        else:
            object.__setattr__(self, "t", float(self.tlist[n]))
            object.__setattr__(self, "n", n + 1)

    def _set_t(self, t):
        """reference ``src/pwc_utils.jl:48-71``"""
        tlist = self.tlist
        N = len(tlist)
        if t <= tlist[0]:
            n = 1
        elif t >= tlist[-1]:
            n = N
        else:
            n = min(int(np.searchsorted(tlist, t, side="left")) + 1, N)
        snapped = float(tlist[n - 1])
        if abs(t - snapped) > 1.4901161193847656e-08 * max(abs(t), abs(snapped)):
            warnings.warn(f"Snapping t={t} to time grid value {snapped}")
        object.__setattr__(self, "n", n - 1 if self.backward else n)
        object.__setattr__(self, "t", snapped)


class ChebyPropagator(PWCPropagator):
    """reference ``src/cheby_propagator.jl:9-27``"""


class NewtonPropagator(PWCPropagator):
    """reference ``src/newton_propagator.jl:9-26``"""


def _method_name(method) -> str:
    if isinstance(method, str):
        return method.lstrip(":").lower()
    name = getattr(method, "__name__", None)  # a module / class named Cheby or Newton
    if name:
        return name.rsplit(".", 1)[-1].lower()
    raise ValueError(f"Unknown propagation `method`: {method!r}")


def _as_device_state(state, ctx):
    if isinstance(state, DeviceState):
        return state, False
    ctx = default_context() if ctx is None else ctx
    return DeviceState.from_host(ctx, np.asarray(state)), True


def init_prop(
    state,
    generator,
    tlist,
    method,
    backward=False,
    inplace=True,
    verbose=False,
    piecewise=None,
    pwc=None,
    parameters=None,
    ctx: Context = None,
    matrix_format="auto",
    **kwargs,
):
    """``init_prop(state, generator, tlist; method, ...)`` (reference ``src/propagator.jl:208-264``,
    ``src/cheby_propagator.jl:87-175``, ``src/newton_propagator.jl:62-113``).

    ``state`` is a DeviceState or a NumPy vector / (N, B) batch (uploaded once); with
    ``inplace=True`` the propagator works on its own copy, as the reference does.
    Cheby keywords: control_ranges, specrange_method, specrange_buffer, cheby_coeffs_limit,
    check_normalization, uniform_dt_tolerance + specrange kwargs (E_min, E_max, rng, ...).
    Newton keywords: m_max, func, norm_min, relerr, max_restarts.
    Engine keyword: matrix_format in {"auto", "csr", "sell", "selld", "bitflip", "dense", "leftright"} -- the device
    storage of the operators (include/qprop.h: QP_FORMAT_*); "auto" picks the fastest one the operators allow.
    """
    name = _method_name(method)
    if name not in ("cheby", "newton"):
        raise ValueError(f"Unknown propagation `method`: {method}")
    tlist = np.array(tlist, dtype=np.float64)
    generator = canonical(generator)
    controls = get_controls(generator)
    dev_state, uploaded = _as_device_state(state, ctx)
    ctx = dev_state.ctx
    G = _max_genop(generator, controls, tlist)
    if not isinstance(G, Operator):
        G = _as_operator(G)  # static generator: one drift term, no coefficients

    if name == "cheby":
        p = ChebyPropagator()
        control_ranges = kwargs.pop("control_ranges", None)
        p.specrange_method = kwargs.pop("specrange_method", "auto")
        p.specrange_buffer = kwargs.pop("specrange_buffer", 0.01)
        limit = kwargs.pop("cheby_coeffs_limit", 1e-12)
        p.check_normalization = kwargs.pop("check_normalization", False)
        uniform_dt_tolerance = kwargs.pop("uniform_dt_tolerance", 1e-12)
        p.specrange_options = dict(kwargs)
        parameters = _process_parameters(parameters, controls, tlist)
        if control_ranges is None:
            control_ranges = IdDict()
            for c in controls:
                vals = discretize(c, tlist)
                control_ranges[c] = (float(np.min(vals)), float(np.max(vals)))
        else:
            for c in controls:
                if c not in control_ranges or not control_ranges[c][0] <= control_ranges[c][1]:
                    raise AssertionError("control_ranges must map every control to (min, max)")
        E_min, E_max = cheby_get_spectral_envelope(
            generator, tlist, control_ranges, p.specrange_method, ctx=ctx, **p.specrange_options
        )
        Delta = E_max - E_min
        if not Delta > 0.0:
            raise AssertionError("spectral range must be positive")
        delta = p.specrange_buffer * Delta
        E_min -= delta / 2
        Delta += delta
        dt = _get_uniform_dt(tlist, tol=uniform_dt_tolerance, warn=True)
        if dt is None:
            raise RuntimeError("Chebychev propagation only works on a uniform time grid")
        p.control_ranges = control_ranges
        p.wrk = ChebyWrk(dev_state, G.to_device(ctx, matrix_format), Delta, E_min, dt, limit=limit)
    else:
        if not inplace:
            raise RuntimeError("The Newton propagator is only implemented in-place")
        p = NewtonPropagator()
        parameters = _process_parameters(parameters, controls, tlist)
        p.wrk = NewtonWrk(dev_state, G.to_device(ctx, matrix_format), m_max=kwargs.get("m_max", 10))
        p.func = kwargs.get("func", None)
        p.norm_min = kwargs.get("norm_min", 1e-14)
        p.relerr = kwargs.get("relerr", 1e-12)
        p.max_restarts = kwargs.get("max_restarts", 50)

    object.__setattr__(p, "generator", generator)
    # the reference copies the caller's state when in-place (src/cheby_propagator.jl:158);
    # a state we just uploaded from the host is already private
    p.state = dev_state.copy() if (inplace and not uploaded) else dev_state
    p.tlist = tlist
    p.parameters = parameters
    p.controls = controls
    p.genop = G
    p.backward = bool(backward)
    p.inplace = bool(inplace)
    p.ctx = ctx
    p.n = len(tlist) - 1 if backward else 1
    p.t = float(tlist[-1]) if backward else float(tlist[0])
    p._host_io = uploaded
    if piecewise is True or pwc is True:
        pass  # both methods are piecewise-constant propagators
    return p


def prop_step(p):
    """``prop_step!(propagator)`` (reference ``src/cheby_propagator.jl:348-386``,
    ``src/newton_propagator.jl:120-153``): one ``qp_cheby_step`` / one Newton restart loop.
    Returns ``propagator.state``, or ``None`` once the time grid is exhausted."""
    n = p.n
    tlist = p.tlist
    if not (0 < n < len(tlist)):
        return None
    H = p._coeffs_for(n)
    if isinstance(p, ChebyPropagator):
        dt = -p.wrk.dt if p.backward else p.wrk.dt
        if p.inplace:
            cheby_(p.state, H, dt, p.wrk, check_normalization=p.check_normalization)
        else:
            p.state = cheby_(p.state.copy(), H, dt, p.wrk, check_normalization=p.check_normalization)
    else:
        dt = float(tlist[n] - tlist[n - 1])
        if p.backward:
            dt = -dt
        newton_(p.state, H, dt, p.wrk, func=p.func, norm_min=p.norm_min, relerr=p.relerr, max_restarts=p.max_restarts)
    p._advance_time()
    return p.state


def set_state(p, state):
    """``set_state!(propagator, state)`` (reference ``src/propagator.jl:367-377``)."""
    if state is not p.state:
        if p.inplace:
            p.state.copyto(state)
        else:
            p.state = state if isinstance(state, DeviceState) else DeviceState.from_host(p.ctx, state)
    return p.state


def set_t(p, t):
    """``set_t!(propagator, t)`` (reference ``src/cheby_propagator.jl:30`` -> ``_pwc_set_t!``)."""
    p._set_t(float(t))


def reinit_prop(p, state, transform_control_ranges=None, **_):
    """``reinit_prop!(propagator, state; transform_control_ranges)`` (reference
    ``src/cheby_propagator.jl:243-299``; default ``src/propagator.jl:298-312``).  For Cheby the
    coefficients are recomputed -- and re-uploaded without touching the operators -- only when
    the current amplitudes left the control ranges they were derived for."""
    state = set_state(p, state)
    tlist = p.tlist
    if isinstance(p, ChebyPropagator):
        transform = transform_control_ranges or (lambda c, lo, hi, check: (lo, hi))
        current = IdDict(
            (c, (float(np.min(p.parameters[c])), float(np.max(p.parameters[c])))) for c in p.controls
        )
        recalc = False
        for c in p.controls:
            lo, hi = transform(c, current[c][0], current[c][1], True)
            if lo < p.control_ranges[c][0] or hi > p.control_ranges[c][1]:
                recalc = True
                break
        if recalc:
            for c in p.controls:
                current[c] = transform(c, current[c][0], current[c][1], False)
            E_min, E_max = cheby_get_spectral_envelope(
                p._generator(), tlist, current, p.specrange_method, ctx=p.ctx, **p.specrange_options
            )
            Delta = E_max - E_min
            if not Delta > 0.0:
                raise AssertionError("spectral range must be positive")
            delta = p.specrange_buffer * Delta
            p.control_ranges = current
            p.wrk.set_spectral_range(Delta + delta, E_min - delta / 2, float(tlist[1] - tlist[0]))
    p.ctx.reset_timings()
    p._set_t(float(tlist[-1] if p.backward else tlist[0]))


def _is_matrix_observable(obs):
    import scipy.sparse as sp

    return sp.issparse(obs) or (isinstance(obs, np.ndarray) and obs.ndim == 2)


def _observable_data(p, observables, i):
    """Data stored for time slot ``i`` (1-based): the reference's ``_write_to_storage!``
    (``src/propagate.jl:346-351``) = ``map_observables(observables, tlist, i, state)``; the default
    observable ``_StoreState`` stores a copy of the state (``src/propagate.jl:14``) -- a host vector
    when the propagation was started from a host array, else a device-resident copy."""
    from .storage import map_observables

    if observables is None:
        return p.state.to_host() if p._host_io else p.state.copy()
    return map_observables(observables, p.tlist, i, p.state)


def _device_observables(ctx, observables):
    """Device generators of matrix observables (cached per matrix by ``storage._device_observable``)."""
    from .storage import _device_observable

    return [_device_observable(o, ctx) for o in observables]


def _propagate_on_device(p, observables, storage):
    """Fast path of ``propagate``: the whole remaining grid in ONE library call
    (``qp_cheby_propagate``), matrix observables evaluated on the device; no per-step host work
    and no state download.  Legal because without a callback nothing can change
    ``propagator.parameters`` between the steps of one ``propagate`` call.  The recorded data go
    through ``write_to_storage`` slot by slot, so any storage the loop path accepts works here."""
    from .storage import write_to_storage

    tlist = p.tlist
    nt = len(tlist)
    steps = list(range(p.n, 0, -1)) if p.backward else list(range(p.n, nt))
    table = np.zeros((len(steps), p.wrk.gen.n_coeffs), dtype=np.complex128)
    for s, n in enumerate(steps):
        H = p._coeffs_for(n)
        if isinstance(H, Operator):
            table[s, :] = H.coeffs
    dt = -p.wrk.dt if p.backward else p.wrk.dt
    obs_gens = _device_observables(p.ctx, observables) if observables else []
    ev, _ = cheby_propagate_(p.state, p.wrk, table, dt, observables=obs_gens)
    for _ in steps:
        p._advance_time()
    if storage is not None:
        single = len(obs_gens) == 1  # map_observables: a single observable gives its bare value
        for s in range(nt):  # storage is written back to front when propagating backward
            slot = (nt - s) if p.backward else s + 1
            write_to_storage(storage, slot, complex(ev[s][0]) if single else np.array(ev[s]))
    return p.state


def propagate(
    state,
    generator=None,
    tlist=None,
    method=None,
    storage=None,
    observables=None,
    callback=None,
    show_progress=False,
    **kwargs,
):
    """``propagate(state, generator, tlist; method, storage, observables, callback, ...)``
    (reference ``src/propagate.jl:167-235, 283-344``), also ``propagate(propagator; ...)`` when
    the first argument is an initialised propagator.

    Returns the final state (a NumPy array if the initial state was one, else the
    DeviceState), or the storage when ``storage=True``.  Storage follows the reference's
    ``Storage`` module exactly (``storage.py``): ``init_storage(state, tlist, observables)``
    allocates it, every slot receives ``map_observables(observables, tlist, i, state)`` through
    ``write_to_storage`` -- so observables may take ``(state)`` or ``(state, tlist, i)``, a single
    observable stores its bare value (scalar data: a length-nt vector; vector data: an n x nt
    matrix), several same-typed ones an n_obs x nt matrix, anything else a list of slots
    (filled back to front when propagating backward)."""
    from .storage import init_storage, write_to_storage

    if isinstance(state, PWCPropagator):
        p = state
    else:
        p = init_prop(state, generator, tlist, method, **kwargs)
    tlist = p.tlist
    nt = len(tlist)
    return_storage = storage is True
    if observables is not None:
        observables = tuple(observables)
    # whole grid in one library call when nothing has to run on the host between the steps:
    # Chebyshev, in place, no callback, and storage (if any) of matrix observables only
    on_device = (
        isinstance(p, ChebyPropagator) and p.inplace and callback is None and not p.check_normalization
        and p.state.batch == 1 and p.n == (nt - 1 if p.backward else 1)
        and (storage is None or (observables is not None and len(observables) > 0
                                 and all(_is_matrix_observable(o) for o in observables)))
    )
    first_slot = nt if p.backward else 1
    if storage is True:
        # init_storage(state, tlist, observables), src/propagate.jl:297-300 / src/storage.jl:38-44
        storage = init_storage(_observable_data(p, observables, 1), nt)
    if on_device:
        _propagate_on_device(p, observables, storage)
        if return_storage:
            return storage
        return p.state.to_host() if p._host_io else p.state
    if storage is not None:
        write_to_storage(storage, first_slot, _observable_data(p, observables, first_slot))
    intervals = range(nt - 1, 0, -1) if p.backward else range(1, nt)
    for i in intervals:
        prop_step(p)
        if callback is not None:
            callback(p, observables)
        if storage is not None:
            slot = i if p.backward else i + 1  # i + (backward ? 0 : 1), src/propagate.jl:331
            write_to_storage(storage, slot, _observable_data(p, observables, slot))
    if return_storage:
        return storage
    return p.state.to_host() if p._host_io else p.state


class Propagation:
    """Wrapper around the arguments of one :func:`propagate` call inside
    :func:`propagate_sequence` (reference ``src/propagate_sequence.jl:1-33``):
    ``Propagation(generator, tlist, **kwargs)`` or ``Propagation(propagator, **kwargs)``; may
    carry its own ``pre_propagation`` / ``post_propagation`` functions."""

    def __init__(self, *args, **kwargs):
        self.args = list(args)
        self.kwargs = dict(kwargs)


def propagate_sequence(state, propagations, pre_propagation=None, post_propagation=None, **kwargs):
    """``propagate_sequence(state, propagations; storage, pre_propagation, post_propagation,
    kwargs...)`` (reference ``src/propagate_sequence.jl:36-137``): a sequence of
    :func:`propagate` calls, each starting from the state the previous one ended with, optionally
    transformed instantaneously before / after each step.  Returns the list of states after each
    step (copies), or the list of storage arrays when ``storage=True``; every other keyword is
    forwarded to ``propagate`` (per-step keywords in the ``Propagation`` are overridden by common
    ones, as in the reference)."""
    psi = state
    results = []
    for prop in propagations:
        if not isinstance(prop, Propagation):
            raise TypeError("propagations must be a list of Propagation instances")
        kw = dict(prop.kwargs)
        pre = kw.pop("pre_propagation", pre_propagation)
        post = kw.pop("post_propagation", post_propagation)
        kw.update(kwargs)
        if pre is not None:
            psi = pre(psi, *prop.args, **kw)
        run_kw = {k: kw.pop(k) for k in ("storage", "observables", "callback", "show_progress") if k in kw}
        if len(prop.args) == 1 and isinstance(prop.args[0], PWCPropagator):
            p = prop.args[0]  # pre-initialised propagator: restarted from the current state
            reinit_prop(p, psi, **kw)
        else:
            p = init_prop(psi, *prop.args, **kw)
        out = propagate(p, **run_kw)
        psi = p.state.to_host() if p._host_io else p.state
        if post is not None:
            psi = post(psi, *prop.args, **kw, **run_kw)
        results.append(out if run_kw.get("storage", None) is True else psi.copy())
    return results
