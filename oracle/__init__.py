"""CPU oracle: a NumPy/SciPy restatement of QuantumPropagators.jl's Chebyshev and
Newton/Arnoldi propagation path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product package (``quantumpropagators.jl_b200``) never does and
fails loudly if its CUDA library is missing.

Every function cites the reference file:line it restates (paths relative to the
reference checkout).  The reference is pure Julia and cannot run in this image (no
``julia``), and it ships no golden vectors; the oracle is therefore pinned against
every analytic / known-answer pin the reference's own tests hold for this path
(``tests/test_oracle_pins.py``):

* ``test/test_propagate.jl:74-150``  TLS Rabi flip, analytic answer, 1e-12
* ``test/test_propagate.jl:153-163`` + ``test/optomech.jl``  Newton == Cheby to 1e-10
* ``test/test_cheby.jl:6-49``        random Hermitian N=1000 vs dense exp, 1e-10,
                                      267/268 coefficients
* ``test/test_newton.jl:7-177``      Hermitian / non-Hermitian / Liouvillian vs dense exp
* ``test/test_specrad.jl:80-223``    spectral range brackets and the exact
                                      ``E_min``/``Δ`` arithmetic of ``init_prop``
* ``test/test_discretization.jl``    midpoint known answers
* ``docs/src/benchmarks/profiling.md:112``  matrix-vector-product counts

Third-party arithmetic restated with SciPy: ``SpecialFunctions.besselj`` ->
``scipy.special.jv``; LAPACK ``eigvals`` -> ``numpy.linalg.eigvals`` (sorted by
(real, imag) like Julia).
"""

from .cheby import cheby_coeffs, cheby_coeffs_inplace, ChebyWrk, cheby_inplace, cheby
from .arnoldi import arnoldi, extend_arnoldi, diagonalize_hessenberg_matrix
from .newton import NewtonWrk, newton_inplace, extend_leja, extend_newton_coeffs, leja_radius
from .specrad import specrange, ritzvals, random_state
from .generators import Generator, Operator, ScaledOperator, hamiltonian, op_mul, op_dot
from .controls import (
    discretize,
    discretize_on_midpoints,
    get_tlist_midpoints,
    t_mid,
    evaluate_control,
)
from .propagator import (
    init_prop,
    prop_step,
    propagate,
    set_state,
    set_t,
    reinit_prop,
    ChebyPropagator,
    NewtonPropagator,
)
