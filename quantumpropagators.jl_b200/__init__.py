"""qprop_b200: B200-native Chebyshev / Newton-Arnoldi propagation engine behind the
QuantumPropagators.jl propagator API (``init_prop`` / ``prop_step!`` / ``propagate``).

The arithmetic runs in ``csrc/libqprop_b200.so`` (hand-written sm_100a CUDA behind the
C ABI of ``include/qprop.h``); this package is the host-side mirror of the reference's
Julia interface for that path (import it as ``qprop_b200`` through the shim at the
repository root).  There is no CPU fallback: any device operation raises
``QPropLibraryError`` / ``QPropError`` if the library or a GPU is missing.
"""

from . import workloads  # noqa: F401  (pure host-side input builders)
from ._lib import QPropError, QPropLibraryError, LIB_PATH  # noqa: F401
from .device import Context, DeviceState, DeviceOperator, DeviceGenerator, default_context  # noqa: F401
from .controls import (  # noqa: F401
    IdDict,
    discretize,
    discretize_on_midpoints,
    get_tlist_midpoints,
    t_mid,
)
from .generators import (  # noqa: F401
    Generator,
    Operator,
    ScaledOperator,
    hamiltonian,
    evaluate,
    evaluate_,
    get_controls,
    substitute,
    liouvillian,
    LeftRightOperator,
)
from . import amplitudes, interfaces, shapes, storage  # noqa: F401
from .amplitudes import LockedAmplitude, ShapedAmplitude, GuidedAmplitude  # noqa: F401
from .storage import (  # noqa: F401
    init_storage,
    map_observables,
    map_observable,
    write_to_storage,
    get_from_storage_,
    get_from_storage,
)
from .interfaces import (  # noqa: F401
    check_amplitude,
    check_control,
    check_generator,
    check_operator,
    check_propagator,
    check_state,
    check_tlist,
    supports_inplace,
    supports_matrix_interface,
    supports_vector_interface,
)
from .cheby import cheby_coeffs, cheby_coeffs_, ChebyWrk, cheby_, cheby, cheby_propagate_  # noqa: F401
from .newton import (  # noqa: F401
    KrylovWrk,
    NewtonWrk,
    newton_,
    arnoldi_,
    extend_arnoldi_,
    diagonalize_hessenberg_matrix,
    extend_leja_,
    extend_newton_coeffs_,
    leja_radius,
)
from .specrad import specrange, ritzvals, random_state  # noqa: F401
from .propagator import (  # noqa: F401
    init_prop,
    prop_step,
    set_state,
    set_t,
    reinit_prop,
    propagate,
    propagate_sequence,
    Propagation,
    ChebyPropagator,
    NewtonPropagator,
    PWCPropagator,
)

__version__ = "0.1.0"
