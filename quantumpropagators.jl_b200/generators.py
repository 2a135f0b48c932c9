"""Host-side generator / operator model (mirror of the operator-algebra part of the
reference's ``src/generators.jl``).

A ``Generator`` is H(t) = Σ_drift H_l + Σ_l a_l(t) H_l; evaluating it on an interval gives a
static ``Operator`` (lazy sum Σ c_l H_l) that shares the component operators and carries only
the numbers c_l.  The component operators are uploaded to the GPU ONCE per generator
(``to_device``); every later ``evaluate_`` only rewrites the coefficient list, which is what
``qp_cheby_step`` / ``qp_arnoldi`` take per call.
"""

from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import controls as _controls
from . import _lib as L
from .device import Context, DeviceGenerator, DeviceState

__all__ = [
    "Generator",
    "Operator",
    "ScaledOperator",
    "hamiltonian",
    "evaluate",
    "evaluate_",
    "get_controls",
    "substitute",
    "liouvillian",
    "ham_to_superop",
    "lindblad_to_superop",
    "LeftRightOperator",
]


def _is_number(x) -> bool:
    return isinstance(x, (int, float, complex, np.number)) and not isinstance(x, bool)


class _DeviceCache:
    """ops -> DeviceGenerator, shared between a Generator and the Operators evaluated from it
    (they share ``ops`` by reference, like the reference asserts in ``src/generators.jl:759``)."""

    def __init__(self):
        self.by_ctx = {}


class Generator:
    """``Generator(ops, amplitudes)`` (reference ``src/generators.jl:44-61``)."""

    def __init__(self, ops, amplitudes):
        ops, amplitudes = list(ops), list(amplitudes)
        if len(amplitudes) > len(ops):
            raise ValueError("The number of amplitudes cannot exceed the number of operators in a Generator")
        if len(amplitudes) < 1:
            raise ValueError("A Generator requires at least one amplitude")
        self.ops = ops
        self.amplitudes = amplitudes
        self._cache = _DeviceCache()

    @property
    def shape(self):
        return self.ops[0].shape

    def __repr__(self):
        return f"Generator with {len(self.ops)} ops and {len(self.amplitudes)} amplitudes"

    def to_device(self, ctx: Context, fmt="auto") -> DeviceGenerator:
        return _to_device(self, ctx, len(self.amplitudes), fmt)


class Operator:
    """``Operator(ops, coeffs)``: lazy sum Σ c_l H_l; with fewer coeffs than ops the leading
    ops have c = 1 (reference ``src/generators.jl:111-125``)."""

    def __init__(self, ops, coeffs, _cache=None):
        coeffs = list(coeffs)
        if len(coeffs) > len(ops):
            raise ValueError("The number of coefficients cannot exceed the number of operators in an Operator")
        self.ops = ops if isinstance(ops, list) else list(ops)
        self.coeffs = coeffs
        self._cache = _cache if _cache is not None else _DeviceCache()

    @property
    def shape(self):
        return self.ops[0].shape

    def __repr__(self):
        return f"Operator with {len(self.ops)} ops and {len(self.coeffs)} coeffs"

    def toarray(self) -> np.ndarray:
        """``Array(O)``: dense sum on the host (small operators only; used by specrange :diag)."""
        drift = len(self.ops) - len(self.coeffs)
        out = np.zeros(self.shape, dtype=np.complex128)
        for i, op in enumerate(self.ops):
            c = self.coeffs[i - drift] if i >= drift else 1.0
            out += c * (op.toarray() if (sp.issparse(op) or hasattr(op, "toarray")) else np.asarray(op))
        return out

    def to_device(self, ctx: Context, fmt="auto") -> DeviceGenerator:
        return _to_device(self, ctx, len(self.coeffs), fmt)

    # -- operator verbs on device states (src/interfaces/operator.jl:21-44) ---------------
    def mul(self, y: DeviceState, x: DeviceState, alpha=1.0, beta=0.0) -> DeviceState:
        """``mul!(y, self, x, α, β)`` (reference ``src/generators.jl:634-645``)."""
        return self.to_device(x.ctx).mul(y, x, self.coeffs, alpha, beta)

    def dot(self, x: DeviceState, y: DeviceState):
        """``dot(x, self, y)`` (reference ``src/generators.jl:648-660``)."""
        return self.to_device(x.ctx).dot(x, y, self.coeffs)

    def __matmul__(self, x: DeviceState) -> DeviceState:
        """``self * x`` (reference ``src/generators.jl:671-684``)."""
        return self.mul(x.similar(), x)

    def __rmul__(self, alpha):
        if not _is_number(alpha):
            return NotImplemented
        return ScaledOperator(alpha, self)

    __mul__ = __rmul__


class ScaledOperator:
    """``ScaledOperator(α, Ĥ)`` (reference ``src/generators.jl:238-249``); α == 1 returns Ĥ."""

    def __new__(cls, coeff, operator):
        if coeff == 1.0:
            return operator
        self = super().__new__(cls)
        self.coeff = coeff
        self.operator = operator
        return self

    @property
    def shape(self):
        return self.operator.shape

    def toarray(self):
        return self.coeff * _toarray(self.operator)

    def mul(self, y, x, alpha=1.0, beta=0.0):
        """reference ``src/generators.jl:701-703``"""
        return _as_operator(self.operator).mul(y, x, self.coeff * alpha, beta)

    def dot(self, x, y):
        """reference ``src/generators.jl:706-708``"""
        return self.coeff * _as_operator(self.operator).dot(x, y)

    def __matmul__(self, x):
        return self.mul(x.similar(), x)

    def __rmul__(self, alpha):
        if not _is_number(alpha):
            return NotImplemented
        return ScaledOperator(alpha * self.coeff, self.operator)

    __mul__ = __rmul__


def _toarray(H):
    if hasattr(H, "toarray"):
        return np.asarray(H.toarray())
    return np.asarray(H)


def _as_operator(H) -> Operator:
    """Wrap a bare matrix as a one-term lazy sum (all drift)."""
    if isinstance(H, Operator):
        return H
    if isinstance(H, ScaledOperator):
        raise TypeError("nested ScaledOperator")
    cache = getattr(H, "_qp_cache", None)
    op = Operator([H], [])
    if cache is not None:
        op._cache = cache
    else:
        try:
            H._qp_cache = op._cache
        except AttributeError:  # ndarray: no attribute slot; cache lives with the wrapper
            pass
    return op


def _to_device(obj, ctx: Context, n_coeffs: int, fmt) -> DeviceGenerator:
    key = (id(ctx), n_coeffs, fmt)
    dev = obj._cache.by_ctx.get(key)
    if dev is None:
        dev = DeviceGenerator(ctx, obj.ops, n_coeffs, fmt)
        obj._cache.by_ctx[key] = dev
    return dev


def hamiltonian(*terms, check=True):
    """``hamiltonian(terms...)`` (reference ``src/generators.jl:388-469``): each term is an
    operator (drift) or a pair ``(op, ampl)``.  Drift terms are summed into one operator,
    terms with the same amplitude are merged; purely numeric amplitudes give an ``Operator``,
    no amplitudes at all give the drift operator itself."""
    drift, ops, amplitudes = [], [], []
    for term in terms:
        if isinstance(term, (tuple, list)):
            if len(term) != 2:
                raise ValueError("time-dependent term must be 2-tuple")
            op, ampl = term
            slot = next(
                (j for j, a in enumerate(amplitudes) if a is ampl or (_is_number(a) and _is_number(ampl) and a == ampl)
                 or (isinstance(a, np.ndarray) and isinstance(ampl, np.ndarray) and np.array_equal(a, ampl))),  # `==` of the reference
                None,
            )
            if slot is None:
                ops.append(op)
                amplitudes.append(ampl)
            else:
                ops[slot] = ops[slot] + op
        elif drift:
            drift[0] = drift[0] + term
        else:
            drift.append(term)
    if not amplitudes:
        if not drift:
            raise ValueError("Generator has no terms")
        return drift[0]
    if all(_is_number(a) for a in amplitudes):
        return Operator(drift + ops, amplitudes)
    return Generator(drift + ops, amplitudes)


def canonical(generator):
    """Tuple generators ``(H0, (H1, ϵ1), ...)`` are canonicalised through ``hamiltonian`` so
    that nothing is materialised per step."""
    if isinstance(generator, (tuple, list)):
        return hamiltonian(*generator, check=False)
    return generator


def get_controls(generator):
    """Unique controls of a generator in order of first appearance (reference
    ``src/generators.jl:711-733``)."""
    generator = canonical(generator)
    if not isinstance(generator, Generator):
        return ()
    found = []
    for ampl in generator.amplitudes:
        for control in _controls.get_controls(ampl):
            if not any(control is c for c in found):
                found.append(control)
    return tuple(found)


def evaluate(generator, *args, vals_dict=None):
    """``evaluate(generator, tlist, n; vals_dict)`` -> static ``Operator`` sharing the
    generator's ops (reference ``src/generators.jl:740-754``); static objects evaluate to
    themselves (``src/controls.jl:309-313``)."""
    generator = canonical(generator)
    if not isinstance(generator, Generator):
        return generator
    coeffs = []
    for i, ampl in enumerate(generator.amplitudes):
        c = _controls.evaluate(ampl, *args, vals_dict=vals_dict)
        if not _is_number(c):
            raise TypeError(f"In `evaluate`, the amplitude {i + 1} evaluates to {type(c)}, not a number")
        coeffs.append(c)
    return Operator(generator.ops, coeffs, _cache=generator._cache)


def evaluate_(op, generator, *args, vals_dict=None):
    """``evaluate!(op, generator, tlist, n; vals_dict)``: rewrites ``op.coeffs`` only
    (reference ``src/generators.jl:757-766``)."""
    if not isinstance(generator, Generator):
        if op is generator:
            return op
        raise TypeError("typeof(op) = typeof(generator), but op ≢ generator")
    if len(op.ops) != len(generator.ops) or any(a is not b for a, b in zip(op.ops, generator.ops)):
        raise AssertionError("op was not evaluated from this generator")
    for i, ampl in enumerate(generator.amplitudes):
        c = _controls.evaluate(ampl, *args, vals_dict=vals_dict)
        if not _is_number(c):
            raise AssertionError("amplitude does not evaluate to a number")
        op.coeffs[i] = c
    return op


def substitute(generator, replacements):
    """``substitute(generator, replacements)`` (reference ``src/controls.jl:476-520``,
    ``src/generators.jl:769-782``): replaces operators and amplitudes / controls by identity;
    tuple generators are substituted term by term."""
    if isinstance(replacements, dict):
        replacements = _controls.IdDict(replacements)
    if generator in replacements:
        return replacements[generator]
    if isinstance(generator, (tuple, list)):
        out = []
        for term in generator:
            if isinstance(term, (tuple, list)):
                op, ampl = term
                out.append((substitute(op, replacements), _controls.substitute(ampl, replacements)))
            else:
                out.append(substitute(term, replacements))
        return tuple(out)
    if isinstance(generator, Generator):
        ops = [substitute(op, replacements) for op in generator.ops]
        amplitudes = [_controls.substitute(a, replacements) for a in generator.amplitudes]
        if all(a is b for a, b in zip(ops, generator.ops)) and all(a is b for a, b in zip(amplitudes, generator.amplitudes)):
            return generator
        return Generator(ops, amplitudes)
    if isinstance(generator, Operator):
        ops = [substitute(op, replacements) for op in generator.ops]
        if all(a is b for a, b in zip(ops, generator.ops)):
            return generator
        return Operator(ops, list(generator.coeffs))
    return generator


# ---------------------------------------------------------------------------------------
# Liouvillian super-operators (reference src/generators.jl:470-632; column-stacking vec)
# ---------------------------------------------------------------------------------------


def _check_convention(convention):
    if convention not in ("TDSE", "LvN"):
        raise ValueError("convention must be TDSE or LvN")


def ham_to_superop(H, convention):
    """𝟙⊗H − Hᵀ⊗𝟙 for ``convention="TDSE"``, times i for ``"LvN"`` (reference
    ``src/generators.jl:470-488``; arXiv:1312.0111 App. B.2)."""
    _check_convention(convention)
    H = sp.csr_matrix(H, dtype=np.complex128)
    ident = sp.identity(H.shape[0], dtype=np.complex128, format="csr")
    Lop = (sp.kron(ident, H, format="csr") - sp.kron(H.T.tocsr(), ident, format="csr")).tocsr()
    if convention == "LvN":
        Lop = (1j * Lop).tocsr()
    Lop.eliminate_zeros()
    Lop.sort_indices()
    return Lop


def lindblad_to_superop(A, convention):
    """(A†)ᵀ⊗A − (𝟙⊗A†A)/2 − ((A†A)ᵀ⊗𝟙)/2 for ``"LvN"``, times i for ``"TDSE"`` (reference
    ``src/generators.jl:491-508``)."""
    _check_convention(convention)
    A = sp.csr_matrix(A, dtype=np.complex128)
    Ad = A.conj().T.tocsr()
    AdA = (Ad @ A).tocsr()
    ident = sp.identity(A.shape[0], dtype=np.complex128, format="csr")
    D = (
        sp.kron(Ad.T.tocsr(), A, format="csr")
        - sp.kron(ident, AdA, format="csr") / 2
        - sp.kron(AdA.T.tocsr(), ident, format="csr") / 2
    ).tocsr()
    if convention == "TDSE":
        D = (1j * D).tocsr()
    D.eliminate_zeros()
    D.sort_indices()
    return D


class LeftRightOperator:
    """Matrix-free super-operator ρ ↦ Σ_t c_t P_t ρ Q_t on column-stacked n × n matrices
    (``vec(P ρ Q) = (Qᵀ ⊗ P) vec ρ``): what ``ham_to_superop`` / ``lindblad_to_superop``
    (reference ``src/generators.jl:470-508``) build explicitly with Kronecker products, kept as
    its n × n factors.  ``terms`` is a list of ``(P, Q, c)`` with ``P`` / ``Q`` sparse matrices or
    ``None`` (identity).  Uploaded with ``qp_op_create_leftright``; the device applies it without
    ever forming the n² × n² matrix (O(nnz) instead of O(n·nnz) memory)."""

    is_leftright = True

    def __init__(self, n, terms):
        self.n = int(n)
        self.terms = []
        left, right = None, None  # all left-only / right-only terms merge into one factor each
        for P, Q, c in terms:
            P = None if P is None else sp.csr_matrix(P, dtype=np.complex128)
            Q = None if Q is None else sp.csr_matrix(Q, dtype=np.complex128)
            for F in (P, Q):
                if F is not None and F.shape != (self.n, self.n):
                    raise ValueError(f"factor of shape {F.shape} in a left/right operator on {self.n} x {self.n} matrices")
            if P is not None and Q is None:
                left = c * P if left is None else left + c * P
            elif P is None and Q is not None:
                right = c * Q if right is None else right + c * Q
            else:
                self.terms.append((P, Q, complex(c)))
        for F, side in ((right, "right"), (left, "left")):
            if F is not None:
                F = F.tocsr()
                F.eliminate_zeros()
                F.sort_indices()
                self.terms.insert(0, (F, None, 1.0 + 0j) if side == "left" else (None, F, 1.0 + 0j))

    @property
    def shape(self):
        return (self.n * self.n, self.n * self.n)

    def tosparse(self):
        """The explicit n² × n² matrix Σ c Qᵀ ⊗ P (for tests and small systems)."""
        ident = sp.identity(self.n, dtype=np.complex128, format="csr")
        out = sp.csr_matrix(self.shape, dtype=np.complex128)
        for P, Q, c in self.terms:
            out = out + c * sp.kron((ident if Q is None else Q).T.tocsr(), ident if P is None else P, format="csr")
        out = out.tocsr()
        out.sort_indices()
        return out

    def toarray(self):
        return self.tosparse().toarray()

    def __add__(self, other):
        if not isinstance(other, LeftRightOperator) or other.n != self.n:
            return NotImplemented
        return LeftRightOperator(self.n, self.terms + other.terms)

    def __mul__(self, alpha):
        if not _is_number(alpha):
            return NotImplemented
        return LeftRightOperator(self.n, [(P, Q, alpha * c) for P, Q, c in self.terms])

    __rmul__ = __mul__

    def __matmul__(self, vec):
        """Host application to a column-stacked matrix (NumPy): Σ c P ρ Q."""
        rho = np.asarray(vec).reshape(self.n, self.n, order="F")
        out = np.zeros_like(rho, dtype=np.complex128)
        for P, Q, c in self.terms:
            t = rho if P is None else P @ rho
            out += c * (t if Q is None else (Q.T @ t.T).T)
        return out.reshape(-1, order="F")

    def __repr__(self):
        return f"LeftRightOperator on {self.n} x {self.n} matrices with {len(self.terms)} terms"


def _ham_lr(H, convention):
    f = 1.0 if convention == "TDSE" else 1j
    H = sp.csr_matrix(H, dtype=np.complex128)
    return LeftRightOperator(H.shape[0], [(H, None, f), (None, H, -f)])


def _lindblad_lr(A, convention):
    f = 1j if convention == "TDSE" else 1.0
    A = sp.csr_matrix(A, dtype=np.complex128)
    Ad = A.conj().T.tocsr()
    AdA = (Ad @ A).tocsr()
    return LeftRightOperator(A.shape[0], [(A, Ad, f), (AdA, None, -0.5 * f), (None, AdA, -0.5 * f)])


def _dissipator(c_ops, convention, matrix_free=False):
    if matrix_free:
        D = _lindblad_lr(c_ops[0], convention)
        for A in c_ops[1:]:
            D = D + _lindblad_lr(A, convention)
        return D
    n = c_ops[0].shape[0]
    if c_ops[0].shape[1] != n:
        raise AssertionError("Lindblad operators must be square")
    D = sp.csr_matrix((n * n, n * n), dtype=np.complex128)
    for A in c_ops:
        D = D + lindblad_to_superop(A, convention)
    D = D.tocsr()
    D.sort_indices()
    return D


def liouvillian(H, c_ops=(), *, convention, check=True, matrix_free=False):
    """``liouvillian(Ĥ, c_ops; convention)`` (reference ``src/generators.jl:520-632``): the
    sparse Liouvillian super-operator of a Hamiltonian (matrix, ``Generator`` / ``Operator`` or
    tuple of terms; ``None`` for a pure dissipator) and Lindblad operators ``c_ops``, acting on
    column-stacked density matrices.  A time-dependent Ĥ gives a ``Generator`` with the same
    amplitudes: drift commutator + dissipator first, then one commutator per control term.
    ``convention`` is mandatory, as in the reference.

    ``matrix_free=True`` (no reference counterpart) returns the same generator
    with every super-operator kept as a :class:`LeftRightOperator` (its n × n factors) instead of
    an n² × n² sparse matrix: identical action on column-stacked density matrices, O(nnz(Ĥ))
    memory."""
    _check_convention(convention)
    c_ops = list(c_ops)
    superop = (lambda M: _ham_lr(M, convention)) if matrix_free else (lambda M: ham_to_superop(M, convention))
    if isinstance(H, (tuple, list)):
        H = hamiltonian(*H, check=check)
    terms = []
    if H is None:
        if not c_ops:
            raise ValueError("Empty Liouvillian, must give at least one of `H` or `c_ops`")
        return hamiltonian(_dissipator(c_ops, convention, matrix_free), check=check)
    if isinstance(H, (Generator, Operator)):
        if c_ops:
            terms.append(_dissipator(c_ops, convention, matrix_free))
        second = H.amplitudes if isinstance(H, Generator) else H.coeffs
        drift = len(H.ops) - len(second)
        for i, op in enumerate(H.ops):
            term = superop(op)
            terms.append(term if i < drift else (term, second[i - drift]))
        return hamiltonian(*terms, check=check)
    terms.append(superop(H))
    if c_ops:
        terms.append(_dissipator(c_ops, convention, matrix_free))
    return hamiltonian(*terms, check=check)
