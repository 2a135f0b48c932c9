"""Device-resident objects: context, operators, lazy-sum generators and states.

``DeviceState`` is the host-side handle of a device-resident state; it implements the verb
set the reference demands of a state (``src/interfaces/state.jl:24-47``: dot, norm, + - *,
copy, zero, similar, copyto!, fill!, lmul!, axpy!) by calling the C ABI, so it plays the
role the ``DeviceState`` type plays in the Julia wrapper.
"""

from __future__ import annotations

import ctypes as C
import weakref

import numpy as np
import scipy.sparse as sp

from . import _lib as L

_default_ctx = {}


class Context:
    """One device + one stream (``qp_ctx_create``)."""

    def __init__(self, device: int = 0):
        lib = L.load()
        h = C.c_void_p()
        L.check(lib.qp_ctx_create(int(device), C.byref(h)), None)
        self.handle = h
        self.device = int(device)
        self._lib = lib
        self._finalizer = weakref.finalize(self, lib.qp_ctx_destroy, h)

    @classmethod
    def borrowed(cls, handle, device: int):
        """Wrap a context owned by someone else (the members of a ``qp_ens_t``): never destroyed here."""
        self = cls.__new__(cls)
        self.handle = handle
        self.device = int(device)
        self._lib = L.load()
        self._finalizer = None
        return self

    def sync(self):
        L.check(self._lib.qp_sync(self.handle), self.handle)

    @property
    def stream(self) -> int:
        s = C.c_void_p()
        L.check(self._lib.qp_ctx_stream(self.handle, C.byref(s)), self.handle)
        return s.value or 0

    @property
    def launch_count(self) -> int:
        n = C.c_int64()
        L.check(self._lib.qp_ctx_launch_count(self.handle, C.byref(n)), self.handle)
        return n.value

    def enable_timings(self, on=True):
        """``QuantumPropagators.enable_timings()`` analogue (reference ``src/timings.jl:31-40``)."""
        L.check(self._lib.qp_timer_enable(self.handle, 1 if on else 0), self.handle)

    def timing(self, label: str):
        n, s = C.c_int64(), C.c_double()
        L.check(self._lib.qp_timer_get(self.handle, label.encode(), C.byref(n), C.byref(s)), self.handle)
        return n.value, s.value

    def reset_timings(self):
        L.check(self._lib.qp_timer_reset(self.handle), self.handle)


def default_context(device: int = 0) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


class DeviceOperator:
    """A component operator uploaded once (``qp_op_upload_sparse`` / ``qp_op_upload_dense``).

    Accepts SciPy CSR/CSC (any index dtype; converted to the Int64 arrays the ABI takes, the
    native layout of Julia's ``SparseMatrixCSC{ComplexF64,Int64}`` up to the index base) or a
    dense ndarray.
    """

    def __init__(self, ctx: Context, A):
        self.ctx = ctx
        lib = ctx._lib
        h = C.c_void_p()
        if getattr(A, "is_leftright", False):
            # matrix-free super-operator Σ c P ρ Q (generators.LeftRightOperator): the n x n factors
            # are uploaded as ordinary sparse operators, the n² x n² matrix is never built
            self.factors = []
            left = (C.c_void_p * len(A.terms))()
            right = (C.c_void_p * len(A.terms))()
            for t, (P, Q, _) in enumerate(A.terms):
                for arr, F in ((left, P), (right, Q)):
                    if F is None:
                        arr[t] = None
                    else:
                        d = DeviceOperator(ctx, sp.csr_matrix(F, dtype=np.complex128))
                        self.factors.append(d)
                        arr[t] = d.handle
            coeffs = L.as_c128_array([c for _, _, c in A.terms])
            L.check(lib.qp_op_create_leftright(ctx.handle, int(A.n), len(A.terms), left, right, L.ptr(coeffs), C.byref(h)), ctx.handle)
            self.dense = False
        elif sp.issparse(A):
            if A.format not in ("csr", "csc"):
                A = A.tocsr()
            layout = L.QP_LAYOUT_CSR if A.format == "csr" else L.QP_LAYOUT_CSC
            indptr = np.ascontiguousarray(A.indptr, dtype=np.int64)
            indices = np.ascontiguousarray(A.indices, dtype=np.int64)
            data = L.as_c128_array(A.data)
            L.check(
                lib.qp_op_upload_sparse(
                    ctx.handle, A.shape[0], A.shape[1], A.nnz, L.ptr(indptr), L.ptr(indices),
                    L.ptr(data), layout, 0, C.byref(h),
                ),
                ctx.handle,
            )
            self.dense = False
        else:
            A = np.asarray(A)
            if A.ndim != 2 or A.shape[0] != A.shape[1]:
                raise ValueError("dense operator must be a square matrix")
            colmajor = np.asfortranarray(A, dtype=np.complex128)
            L.check(lib.qp_op_upload_dense(ctx.handle, A.shape[0], L.ptr(colmajor), C.byref(h)), ctx.handle)
            self.dense = True
        self.handle = h
        self.shape = tuple(A.shape)
        self._finalizer = weakref.finalize(self, lib.qp_op_destroy, h)


class DeviceGenerator:
    """Device form of the lazy sum ``Operator(ops, coeffs)`` (``qp_gen_create``): the first
    ``n_ops - n_coeffs`` operators are drift terms (reference ``src/generators.jl:634-636``)."""

    def __init__(self, ctx: Context, ops, n_coeffs: int, fmt=L.QP_FORMAT_AUTO):
        self.ctx = ctx
        lib = ctx._lib
        self.ops = [op if isinstance(op, DeviceOperator) else DeviceOperator(ctx, op) for op in ops]
        arr = (C.c_void_p * len(self.ops))(*[op.handle for op in self.ops])
        h = C.c_void_p()
        if isinstance(fmt, str):
            fmt = {v: k for k, v in L.FORMAT_NAMES.items()}[fmt]
        L.check(lib.qp_gen_create(ctx.handle, len(self.ops), arr, int(n_coeffs), int(fmt), C.byref(h)), ctx.handle)
        self.handle = h
        self.n_ops = len(self.ops)
        self.n_coeffs = int(n_coeffs)
        self._finalizer = weakref.finalize(self, lib.qp_gen_destroy, h)
        f, n, stored, mbytes = C.c_int32(), C.c_int64(), C.c_int64(), C.c_int64()
        L.check(lib.qp_gen_info(h, C.byref(f), C.byref(n), C.byref(stored), C.byref(mbytes)), ctx.handle)
        self.format = L.FORMAT_NAMES[f.value]
        self.n = n.value
        self.stored_entries = stored.value
        self.matrix_bytes = mbytes.value
        self.shape = (self.n, self.n)
        sb, nd, cb = C.c_int64(), C.c_int32(), C.c_int32()
        L.check(lib.qp_gen_storage(h, C.byref(sb), C.byref(nd), C.byref(cb)), ctx.handle)
        self.stored_bytes = sb.value  # matrix bytes one application actually streams
        self.n_dict = nd.value  # SELL-D: table entries (0 otherwise)
        self.code_bytes = cb.value

    def tile_info(self):
        """The two-pass tiled form used for batched states (``qp_gen_tile_info``; built on first
        use): ``{"available", "split", "blocks", "n_table", "entries": {"A", "B", "other", "diag"}}``."""
        av, sp_, nb, nt = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        ent = np.zeros(4, dtype=np.int64)
        L.check(self.ctx._lib.qp_gen_tile_info(self.handle, C.byref(av), C.byref(sp_), C.byref(nb), C.byref(nt), L.ptr(ent)),
                self.ctx.handle)
        return {"available": bool(av.value), "split": sp_.value, "blocks": nb.value, "n_table": nt.value,
                "entries": dict(zip(("A", "B", "other", "diag"), (int(v) for v in ent)))}

    def _coeffs(self, coeffs):
        c = L.as_c128_array(coeffs if coeffs is not None else [])
        if c.size != self.n_coeffs:
            raise ValueError(f"expected {self.n_coeffs} coefficients, got {c.size}")
        return c

    def mul(self, y: "DeviceState", x: "DeviceState", coeffs, alpha=1.0, beta=0.0):
        """``mul!(y, H, x, α, β)`` with H = Σ c_l H_l (reference ``src/generators.jl:634-645``)."""
        c = self._coeffs(coeffs)
        L.check(
            self.ctx._lib.qp_gen_mul(self.handle, L.ptr(c), L.to_c128(alpha), L.to_c128(beta), x.handle, y.handle),
            self.ctx.handle,
        )
        return y

    def expval(self, x: "DeviceState", coeffs=None):
        """⟨x|H|x⟩ per trajectory in one fused pass (``qp_gen_expval``): the expectation value of
        a matrix observable (reference ``src/storage.jl:100-123``)."""
        c = self._coeffs(coeffs)
        out = np.zeros(x.batch, dtype=np.complex128)
        L.check(self.ctx._lib.qp_gen_expval(self.handle, L.ptr(c), x.handle, L.ptr(out)), self.ctx.handle)
        return complex(out[0]) if x.batch == 1 else out

    def dot(self, x: "DeviceState", y: "DeviceState", coeffs):
        """3-argument ``dot(x, H, y)`` (reference ``src/generators.jl:648-660``)."""
        c = self._coeffs(coeffs)
        out = np.zeros(x.batch, dtype=np.complex128)
        L.check(self.ctx._lib.qp_gen_dot(self.handle, L.ptr(c), x.handle, y.handle, L.ptr(out)), self.ctx.handle)
        return complex(out[0]) if x.batch == 1 else out


class DeviceState:
    """Handle of a device-resident state of dimension ``n`` with ``batch`` trajectories,
    stored [n][batch] (batch fastest)."""

    def __init__(self, ctx: Context, n: int, batch: int = 1):
        self.ctx = ctx
        lib = ctx._lib
        h = C.c_void_p()
        L.check(lib.qp_state_create(ctx.handle, int(n), int(batch), C.byref(h)), ctx.handle)
        self.handle = h
        self.n = int(n)
        self.batch = int(batch)
        self._finalizer = weakref.finalize(self, lib.qp_state_destroy, h)

    # -- construction / transfer ---------------------------------------------------------
    @classmethod
    def from_host(cls, ctx: Context, psi) -> "DeviceState":
        psi = np.asarray(psi)
        if psi.ndim == 1:
            st = cls(ctx, psi.shape[0], 1)
        elif psi.ndim == 2:
            st = cls(ctx, psi.shape[0], psi.shape[1])
        else:
            raise ValueError("state must be a vector (n,) or a batch (n, B)")
        st.upload(psi)
        return st

    def upload(self, psi, b0: int = 0):
        psi = L.as_c128_array(psi)
        nb = 1 if psi.ndim == 1 else psi.shape[1]
        if psi.shape[0] != self.n:
            raise ValueError(f"state dimension {psi.shape[0]} != {self.n}")
        L.check(self.ctx._lib.qp_state_upload(self.handle, L.ptr(psi), int(b0), int(nb)), self.ctx.handle)
        return self

    def download(self, b0: int = 0, nb=None) -> np.ndarray:
        nb = self.batch - b0 if nb is None else nb
        out = np.empty((self.n, nb), dtype=np.complex128)
        L.check(self.ctx._lib.qp_state_download(self.handle, L.ptr(out), int(b0), int(nb)), self.ctx.handle)
        return out[:, 0].copy() if self.batch == 1 else out

    def to_host(self) -> np.ndarray:
        return self.download()

    @property
    def devptr(self) -> int:
        p = C.c_void_p()
        L.check(self.ctx._lib.qp_state_devptr(self.handle, C.byref(p)), self.ctx.handle)
        return p.value

    @property
    def __cuda_array_interface__(self):
        """Zero-copy view for torch (``torch.as_tensor(state, device='cuda')``): used by the
        multi-GPU gather of final states.  Valid until the next ``qp_cheby_step`` on this state
        (which swaps buffers)."""
        shape = (self.n,) if self.batch == 1 else (self.n, self.batch)
        return {"shape": shape, "typestr": "<c16", "data": (self.devptr, False), "version": 3, "strides": None}

    # -- the state verbs (src/interfaces/state.jl:24-47) ---------------------------------
    def similar(self) -> "DeviceState":
        return DeviceState(self.ctx, self.n, self.batch)

    def copy(self) -> "DeviceState":
        out = self.similar()
        out.copyto(self)
        return out

    def zero(self) -> "DeviceState":
        out = self.similar()
        out.fill(0.0)
        return out

    def copyto(self, src) -> "DeviceState":
        """``copyto!(self, src)``; ``src`` may be a DeviceState or a host array."""
        if isinstance(src, DeviceState):
            L.check(self.ctx._lib.qp_copy(self.handle, src.handle), self.ctx.handle)
        else:
            self.upload(src)
        return self

    def fill(self, value) -> "DeviceState":
        L.check(self.ctx._lib.qp_fill(self.handle, L.to_c128(value)), self.ctx.handle)
        return self

    def lmul(self, alpha) -> "DeviceState":
        """``lmul!(α, self)``."""
        L.check(self.ctx._lib.qp_scal(self.handle, L.to_c128(alpha)), self.ctx.handle)
        return self

    def axpy(self, alpha, x: "DeviceState") -> "DeviceState":
        """``axpy!(α, x, self)``: self += α x."""
        L.check(self.ctx._lib.qp_axpy(L.to_c128(alpha), x.handle, self.handle), self.ctx.handle)
        return self

    def dot(self, other: "DeviceState"):
        """``dot(self, other)`` = ⟨self|other⟩ (conjugate-linear in self)."""
        out = np.zeros(self.batch, dtype=np.complex128)
        L.check(self.ctx._lib.qp_dot(self.handle, other.handle, L.ptr(out)), self.ctx.handle)
        return complex(out[0]) if self.batch == 1 else out

    def norm(self):
        out = np.zeros(self.batch, dtype=np.float64)
        L.check(self.ctx._lib.qp_norm(self.handle, L.ptr(out)), self.ctx.handle)
        return float(out[0]) if self.batch == 1 else out

    def __add__(self, other):
        return self.copy().axpy(1.0, other)

    def __sub__(self, other):
        return self.copy().axpy(-1.0, other)

    def __mul__(self, alpha):
        return self.copy().lmul(alpha)

    __rmul__ = __mul__

    def __len__(self):
        return self.n
