"""Storage of propagated states / observable data (mirror of the reference's ``Storage`` module,
``src/storage.jl``): ``init_storage``, ``map_observables``, ``map_observable``,
``write_to_storage`` (``write_to_storage!``), ``get_from_storage_`` (``get_from_storage!``) and
``get_from_storage``.  Time-slot indices ``i`` are 1-based like everywhere in the host mirror.

Device awareness: a ``DeviceState`` is not a vector, so the default
storage for states is a list of per-slot copies (device-resident, as the reference stores
``copy(state)`` per slot for non-vector states, ``src/storage.jl:46``); vector-valued data goes
into an ``n × nt`` matrix; the expectation value of a matrix observable with a device state is the
fused one-pass ``qp_gen_expval`` instead of a download (``src/storage.jl:115-117``)."""

from __future__ import annotations

import inspect

import numpy as np
import scipy.sparse as sp

from .device import DeviceState

__all__ = [
    "init_storage",
    "map_observables",
    "map_observable",
    "write_to_storage",
    "get_from_storage_",
    "get_from_storage",
]

_OBS_CACHE_MAX = 32
_obs_cache = {}  # (id(observable), id(ctx)) -> (observable, data fingerprint, device generator); bounded


def _is_matrix(obs) -> bool:
    return sp.issparse(obs) or (isinstance(obs, np.ndarray) and obs.ndim == 2)


def _fingerprint(obs):
    """Cheap content check of a matrix observable, so that a matrix mutated in place after its
    first use is uploaded again instead of silently using the stale device copy."""
    data = obs.data if sp.issparse(obs) else obs
    data = np.asarray(data)
    step = max(1, data.size // 64)
    return (data.shape, getattr(obs, "nnz", None), complex(np.sum(data.reshape(-1)[::step])), complex(data.reshape(-1)[-1]) if data.size else 0)


def _device_observable(observable, ctx):
    """Device generator (no free coefficients) of a matrix observable, uploaded once per
    (matrix, context); the cache is bounded and keyed on identity + a content fingerprint."""
    from .generators import _as_operator

    key = (id(observable), id(ctx))
    fp = _fingerprint(observable)
    hit = _obs_cache.get(key)
    if hit is None or hit[0] is not observable or hit[1] != fp:
        if len(_obs_cache) >= _OBS_CACHE_MAX:
            _obs_cache.pop(next(iter(_obs_cache)))
        hit = (observable, fp, _as_operator(observable).to_device(ctx))
        _obs_cache[key] = hit
    return hit[2]


def map_observable(observable, tlist, i, state):
    """``map_observable(observable, tlist, i, state)`` (reference ``src/storage.jl:98-117``): a
    function of ``(state, tlist, i)`` or of ``(state,)``, or a matrix whose expectation value
    ⟨state|O|state⟩ is returned."""
    if _is_matrix(observable):
        if isinstance(state, DeviceState):
            return _device_observable(observable, state.ctx).expval(state, [])
        return np.vdot(state, observable @ state)
    if callable(observable):
        try:
            n_pos = sum(
                1 for p in inspect.signature(observable).parameters.values()
                if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD) and p.default is p.empty
            )
        except (TypeError, ValueError):
            n_pos = 1
        if n_pos >= 3:
            return observable(state, tlist, i)
        if n_pos == 1:
            return observable(state)
        raise TypeError(
            f"The `observable` function {observable} must take either the single argument `state`, "
            "or the three arguments `state`, `tlist`, and `i`."
        )
    raise TypeError(f"unsupported observable of type {type(observable).__name__}")


def map_observables(observables, tlist, i, state):
    """``map_observables(observables, tlist, i, state)`` (reference ``src/storage.jl:65-80``):
    a single observable gives its own data; several give a vector when all results have the same
    type (compact matrix storage), otherwise a tuple."""
    observables = tuple(observables)
    if len(observables) == 1:
        return map_observable(observables[0], tlist, i, state)
    vals = tuple(map_observable(O, tlist, i, state) for O in observables)
    first = type(vals[0])
    if all(type(v) is first for v in vals[1:]) and not isinstance(vals[0], (DeviceState, tuple, list)):
        return np.array(vals)
    return vals


def _is_scalar(data) -> bool:
    return isinstance(data, (bool, int, float, complex, np.number))


def init_storage(state_or_data, tlist_or_nt, observables=None):
    """``init_storage(state, tlist)``, ``init_storage(state, tlist, observables)`` or
    ``init_storage(data, nt)`` (reference ``src/storage.jl:33-46``): an ``n × nt`` matrix for
    vector data of length ``n``, a length-``nt`` vector for numbers (``Vector{T}(undef, nt)``),
    otherwise a list of ``nt`` slots."""
    if observables is not None:
        data = map_observables(observables, tlist_or_nt, 1, state_or_data)
        nt = len(tlist_or_nt)
    else:
        data = state_or_data
        nt = int(tlist_or_nt) if np.isscalar(tlist_or_nt) else len(tlist_or_nt)
    if isinstance(data, np.ndarray) and data.ndim == 1:
        return np.empty((data.shape[0], nt), dtype=data.dtype)
    if _is_scalar(data):
        return np.empty(nt, dtype=np.result_type(data))
    return [None] * nt


def write_to_storage(storage, i, data):
    """``write_to_storage!(storage, i, data)`` (reference ``src/storage.jl:138-144``): column ``i``
    of a matrix storage, element ``i`` of a vector, slot ``i`` of a list (device states are stored
    as copies)."""
    if isinstance(storage, np.ndarray):
        if storage.ndim == 1:
            storage[i - 1] = data
        else:
            storage[:, i - 1] = data
    else:
        storage[i - 1] = data.copy() if isinstance(data, (DeviceState, np.ndarray)) else data
    return storage


def get_from_storage_(data, storage, i):
    """``get_from_storage!(data, storage, i)`` (reference ``src/storage.jl:168-169``): copies time
    slot ``i`` into ``data`` (a NumPy array or a ``DeviceState``) and returns it."""
    src = get_from_storage(storage, i)
    if isinstance(data, DeviceState):
        return data.copyto(src)
    np.copyto(data, src.to_host() if isinstance(src, DeviceState) else src)
    return data


def get_from_storage(storage, i):
    """``get_from_storage(storage, i)`` (reference ``src/storage.jl:180-181``)."""
    if isinstance(storage, np.ndarray) and storage.ndim == 2:
        return storage[:, i - 1]
    return storage[i - 1]
