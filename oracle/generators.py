"""Oracle restatement of the operator-algebra part of ``src/generators.jl``.

Test infrastructure only.  Component operators are ``scipy.sparse`` matrices or dense
``numpy.ndarray``; states are complex128 arrays of shape (N,) or (N, B).
"""

from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .controls import evaluate_control, get_controls_of_amplitude


def _is_number(x) -> bool:
    return isinstance(x, (int, float, complex, np.number)) and not isinstance(x, bool)


class Generator:
    """``Generator(ops, amplitudes)`` -- ``src/generators.jl:44-61``.

    H(t) = Σ_{l<=drift} ops[l] + Σ_l a_l(t) ops[drift+l]; the first
    ``len(ops) - len(amplitudes)`` operators are drift terms.
    """

    def __init__(self, ops, amplitudes):
        ops = list(ops)
        amplitudes = list(amplitudes)
        if len(amplitudes) > len(ops):
            raise ValueError(
                "The number of amplitudes cannot exceed the number of operators in a Generator"
            )
        if len(amplitudes) < 1:
            raise ValueError("A Generator requires at least one amplitude")
        self.ops = ops
        self.amplitudes = amplitudes


class Operator:
    """``Operator(ops, coeffs)`` -- ``src/generators.jl:111-125`` (lazy sum Σ c_l H_l)."""

    def __init__(self, ops, coeffs):
        ops = list(ops)
        coeffs = list(coeffs)
        if len(coeffs) > len(ops):
            raise ValueError(
                "The number of coefficients cannot exceed the number of operators in an Operator"
            )
        self.ops = ops
        self.coeffs = coeffs

    @property
    def shape(self):
        return self.ops[0].shape

    def toarray(self):
        """``Array(O::Operator)``: the dense sum."""
        drift_offset = len(self.ops) - len(self.coeffs)
        A = np.zeros(self.shape, dtype=np.complex128)
        for i, op in enumerate(self.ops):
            c = self.coeffs[i - drift_offset] if i >= drift_offset else 1.0
            A = A + c * (op.toarray() if sp.issparse(op) else np.asarray(op))
        return A

    def __matmul__(self, Psi):
        return op_mul_psi(self, Psi, 1.0)


class ScaledOperator:
    """``ScaledOperator(α, Ĥ)`` -- ``src/generators.jl:238-249``."""

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
        return self.coeff * toarray(self.operator)

    def __matmul__(self, Psi):
        return op_mul_psi(self.operator, Psi, self.coeff)


def toarray(H) -> np.ndarray:
    if isinstance(H, (Operator, ScaledOperator)):
        return H.toarray()
    if sp.issparse(H):
        return H.toarray()
    return np.asarray(H)


def _mul5(C, A, B, alpha, beta):
    """5-argument ``mul!(C, A, B, α, β)``: C ← β·C + α·A·B for a plain matrix A."""
    AB = A @ B
    if beta is False or beta == 0:
        C[...] = alpha * AB
    else:
        C *= beta
        C += alpha * AB
    return C


def op_mul(C, A, B, alpha=True, beta=False):
    """``LinearAlgebra.mul!(C, A::Operator, B, α, β)`` -- ``src/generators.jl:634-645``
    and ``mul!(C, A::ScaledOperator, B, α, β)`` -- ``:701-703``.

    One 5-arg ``mul!`` per component operator; the first uses β, the rest accumulate.
    """
    if isinstance(A, ScaledOperator):
        return op_mul(C, A.operator, B, A.coeff * alpha, beta)
    if not isinstance(A, Operator):
        return _mul5(C, A, B, alpha, beta)
    drift_offset = len(A.ops) - len(A.coeffs)
    c = alpha
    if drift_offset == 0:
        c = c * A.coeffs[0]
    _mul5(C, A.ops[0], B, c, beta)
    for i in range(1, len(A.ops)):
        c = alpha
        if i >= drift_offset:
            c = c * A.coeffs[i - drift_offset]
        _mul5(C, A.ops[i], B, c, True)
    return C


def op_mul_psi(O, Psi, c):
    """``_op_mul_psi(O::Operator, Ψ, c)`` -- ``src/generators.jl:671-684``."""
    drift_offset = len(O.ops) - len(O.coeffs)
    a = c
    if drift_offset == 0:
        a = a * O.coeffs[0]
    Phi = a * (O.ops[0] @ Psi)
    for i in range(1, len(O.ops)):
        a = c
        if i >= drift_offset:
            a = a * O.coeffs[i - drift_offset]
        Phi = Phi + a * (O.ops[i] @ Psi)
    return Phi


def op_dot(x, A, y):
    """3-argument ``dot(x, A, y)`` = ⟨x|A|y⟩ -- ``src/generators.jl:648-660, 706-708``."""
    if isinstance(A, ScaledOperator):
        return A.coeff * op_dot(x, A.operator, y)
    if not isinstance(A, Operator):
        return complex(np.vdot(x, A @ y))
    drift_offset = len(A.ops) - len(A.coeffs)
    result = 0j
    for i, op in enumerate(A.ops):
        if i >= drift_offset:
            result += A.coeffs[i - drift_offset] * np.vdot(x, op @ y)
        else:
            result += np.vdot(x, op @ y)
    return complex(result)


def matvec(H, x):
    """3-argument ``mul!(y, H, x)`` as used in ``src/cheby.jl:176,190`` and
    ``src/arnoldi.jl:82,120`` (returns a fresh array)."""
    if isinstance(H, (Operator, ScaledOperator)):
        y = np.empty_like(x)
        return op_mul(y, H, x, True, False)
    return H @ x


def get_controls(generator):
    """``get_controls(generator)`` -- ``src/generators.jl:711-733``; unique controls in
    order of first appearance (identity comparison, like the reference's ``IdDict``)."""
    if isinstance(generator, (tuple, list)):
        generator = hamiltonian(*generator, check=False)
    if not isinstance(generator, Generator):
        return tuple()
    controls = []
    for ampl in generator.amplitudes:
        for control in get_controls_of_amplitude(ampl):
            if not any(control is c for c in controls):
                controls.append(control)
    return tuple(controls)


def evaluate(generator, *args, vals_dict=None):
    """``evaluate(generator::Generator, args...; vals_dict)`` -- ``src/generators.jl:740-754``;
    bare matrices / Operators evaluate to themselves (``src/controls.jl:309-313``)."""
    if isinstance(generator, (tuple, list)):
        generator = hamiltonian(*generator, check=False)
    if not isinstance(generator, Generator):
        return generator
    coeffs = []
    for i, ampl in enumerate(generator.amplitudes):
        coeff = evaluate_control(ampl, *args, vals_dict=vals_dict)
        if not _is_number(coeff):
            raise TypeError(f"amplitude {i} evaluates to {type(coeff)}, not a number")
        coeffs.append(coeff)
    return Operator(generator.ops, coeffs)


def evaluate_inplace(op, generator, *args, vals_dict=None):
    """``evaluate!(op::Operator, generator::Generator, args...; vals_dict)`` --
    ``src/generators.jl:757-766``: only ``op.coeffs`` is rewritten."""
    if not isinstance(generator, Generator):
        if op is generator:
            return op  # src/controls.jl:466-475
        raise TypeError("typeof(op) = typeof(generator), but op ≢ generator")
    assert len(op.ops) == len(generator.ops)
    assert all(O is P for O, P in zip(op.ops, generator.ops))
    for i, ampl in enumerate(generator.amplitudes):
        coeff = evaluate_control(ampl, *args, vals_dict=vals_dict)
        assert _is_number(coeff)
        op.coeffs[i] = coeff
    return op


def hamiltonian(*terms, check=True):
    """``hamiltonian(terms...)`` / ``_make_generator`` -- ``src/generators.jl:388-469``.

    Drift terms are summed into one operator; terms sharing an amplitude are merged;
    all-numeric amplitudes give an ``Operator``; no amplitudes give the drift itself.
    """
    ops, drift, amplitudes = [], [], []
    for term in terms:
        if isinstance(term, (tuple, list)):
            if len(term) != 2:
                raise ValueError("time-dependent term must be 2-tuple")
            op, ampl = term
            idx = None
            for j, a in enumerate(amplitudes):
                if a is ampl or (_is_number(a) and _is_number(ampl) and a == ampl):
                    idx = j
                    break
            if idx is None:
                ops.append(op)
                amplitudes.append(ampl)
            else:
                ops[idx] = ops[idx] + op
        else:
            if not drift:
                drift.append(term)
            else:
                drift[0] = drift[0] + term
    all_ops = drift + ops
    if not amplitudes:
        if not drift:
            raise ValueError("Generator has no terms")
        return drift[0]
    if all(_is_number(a) for a in amplitudes):
        return Operator(all_ops, amplitudes)
    return Generator(all_ops, amplitudes)
