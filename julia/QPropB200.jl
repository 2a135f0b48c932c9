# QPropB200.jl -- Julia host side of the B200 engine: the `ccall` bindings of
# include/qprop.h and the method plugins `method=ChebyB200` / `method=NewtonB200` for
# QuantumPropagators.jl.
#
# STATUS: written against QuantumPropagators v0.8.5 (`/root/reference`), NOT executed in the
# build image (no Julia there).  The executed and tested host side is the Python mirror in
# `quantumpropagators.jl_b200/`, which makes the same C-ABI calls in the same order; the
# mapping is documented in INTEGRATION.md.
#
# Usage (unchanged user code, only the `method` differs):
#
#     using QuantumPropagators, QPropB200
#     Ψ = propagate(Ψ₀, generator, tlist; method=ChebyB200, E_min=-10.0, E_max=10.0)
#
# Selection follows the reference's own plugin route: `init_prop(state, generator, tlist,
# ::Val{:ChebyB200}; ...)` is found through `Val(nameof(method))`
# (reference src/propagator.jl:255-258; precedent ext/QuantumPropagatorsExponentialUtilitiesExt.jl).

module QPropB200

using LinearAlgebra
using SparseArrays
using OffsetArrays
import QuantumPropagators
import QuantumPropagators: init_prop, prop_step!, set_t!, set_state!, reinit_prop!, PWCPropagator
import QuantumPropagators: _pwc_get_max_genop, _pwc_process_parameters, _pwc_set_t!,
    _pwc_set_genop!, _pwc_advance_time!, _get_uniform_dt, cheby_get_spectral_envelope
import QuantumPropagators.Interfaces: supports_inplace
using QuantumPropagators.Controls: get_controls, discretize
using QuantumPropagators.Generators: Generator, Operator
using QuantumPropagators.Cheby: cheby_coeffs
using QuantumPropagators.Arnoldi: diagonalize_hessenberg_matrix
using QuantumPropagators.Newton: extend_leja!, extend_newton_coeffs!, leja_radius

export ChebyB200, NewtonB200, DeviceState, to_device, to_host, propagate_on_device!, expval
export EnsembleB200, BatchedState, ensemble_propagate!, gather_states, gather_expvals, device_specrange

const libqprop = get(ENV, "QPROP_B200_LIB", joinpath(@__DIR__, "..", "quantumpropagators.jl_b200",
                                                      "csrc", "libqprop_b200.so"))

# `method=ChebyB200` / `method=NewtonB200`: modules, like `QuantumPropagators.Cheby`
module ChebyB200 end
module NewtonB200 end

const QP_LAYOUT_CSC = Int32(0)
const QP_FORMAT_AUTO = Int32(0)
const QP_FORMAT_BITFLIP = Int32(6)  # diagonal + (conditional) uniform bit flips, detected by the library (AUTO selects it)

# ---------------------------------------------------------------------------------------
# status -> exception (reference error conventions, SURVEY.md §5)
# ---------------------------------------------------------------------------------------
function check(status::Int32, ctx::Ptr{Cvoid}=C_NULL)
    status == 0 && return
    msg = unsafe_string(ccall((:qp_last_error, libqprop), Cstring, (Ptr{Cvoid},), ctx))
    status == -1 && throw(ArgumentError(msg))
    status == -3 && throw(OutOfMemoryError())
    (status == -4 || status == -5) && throw(AssertionError(msg))  # src/newton.jl:375, src/cheby.jl:196
    error("libqprop_b200 [$status]: $msg")
end

mutable struct Context
    handle::Ptr{Cvoid}
    function Context(device::Integer=0)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:qp_ctx_create, libqprop), Int32, (Int32, Ref{Ptr{Cvoid}}), device, h))
        ctx = new(h[])
        finalizer(c -> ccall((:qp_ctx_destroy, libqprop), Int32, (Ptr{Cvoid},), c.handle), ctx)
    end
    # a context owned by someone else (the members of an ensemble): no finalizer
    Context(handle::Ptr{Cvoid}, owned::Bool) = new(handle)
end

const _default_ctx = Ref{Union{Nothing,Context}}(nothing)
default_context() = something(_default_ctx[], (_default_ctx[] = Context(0)))

# ---------------------------------------------------------------------------------------
# DeviceState: satisfies check_state (reference src/interfaces/state.jl:24-47) without being
# an AbstractVector, so supports_vector_interface stays false (SURVEY.md §8b)
# ---------------------------------------------------------------------------------------
mutable struct DeviceState
    ctx::Context
    handle::Ptr{Cvoid}
    n::Int64
    function DeviceState(ctx::Context, n::Integer)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:qp_state_create, libqprop), Int32, (Ptr{Cvoid}, Int64, Int64, Ref{Ptr{Cvoid}}),
                    ctx.handle, n, 1, h), ctx.handle)
        st = new(ctx, h[], n)
        finalizer(s -> ccall((:qp_state_destroy, libqprop), Int32, (Ptr{Cvoid},), s.handle), st)
    end
end

function to_device(Ψ::Vector{ComplexF64}; ctx=default_context())
    st = DeviceState(ctx, length(Ψ))
    check(ccall((:qp_state_upload, libqprop), Int32, (Ptr{Cvoid}, Ptr{ComplexF64}, Int64, Int64),
                st.handle, Ψ, 0, 1), ctx.handle)
    return st
end

function to_host(st::DeviceState)
    Ψ = Vector{ComplexF64}(undef, st.n)
    check(ccall((:qp_state_download, libqprop), Int32, (Ptr{Cvoid}, Ptr{ComplexF64}, Int64, Int64),
                st.handle, Ψ, 0, 1), st.ctx.handle)
    return Ψ
end

supports_inplace(::Type{DeviceState}) = true
Base.length(st::DeviceState) = st.n
Base.similar(st::DeviceState) = DeviceState(st.ctx, st.n)
Base.copy(st::DeviceState) = copyto!(similar(st), st)
Base.zero(st::DeviceState) = fill!(similar(st), 0)
function Base.copyto!(dst::DeviceState, src::DeviceState)
    check(ccall((:qp_copy, libqprop), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), dst.handle, src.handle), dst.ctx.handle)
    return dst
end
function Base.fill!(st::DeviceState, v)
    check(ccall((:qp_fill, libqprop), Int32, (Ptr{Cvoid}, ComplexF64), st.handle, ComplexF64(v)), st.ctx.handle)
    return st
end
function LinearAlgebra.lmul!(α::Number, st::DeviceState)
    check(ccall((:qp_scal, libqprop), Int32, (Ptr{Cvoid}, ComplexF64), st.handle, ComplexF64(α)), st.ctx.handle)
    return st
end
function LinearAlgebra.axpy!(α::Number, x::DeviceState, y::DeviceState)
    check(ccall((:qp_axpy, libqprop), Int32, (ComplexF64, Ptr{Cvoid}, Ptr{Cvoid}), ComplexF64(α), x.handle, y.handle),
          y.ctx.handle)
    return y
end
function LinearAlgebra.dot(x::DeviceState, y::DeviceState)
    out = Ref{ComplexF64}(0)
    check(ccall((:qp_dot, libqprop), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{ComplexF64}), x.handle, y.handle, out),
          x.ctx.handle)
    return out[]
end
function LinearAlgebra.norm(x::DeviceState)
    out = Ref{Float64}(0)
    check(ccall((:qp_norm, libqprop), Int32, (Ptr{Cvoid}, Ref{Float64}), x.handle, out), x.ctx.handle)
    return out[]
end
Base.:+(a::DeviceState, b::DeviceState) = axpy!(1, b, copy(a))
Base.:-(a::DeviceState, b::DeviceState) = axpy!(-1, b, copy(a))
Base.:*(α::Number, a::DeviceState) = lmul!(α, copy(a))
Base.:*(a::DeviceState, α::Number) = α * a

# ---------------------------------------------------------------------------------------
# operators / generators: Julia's native SparseMatrixCSC{ComplexF64,Int64} (1-based CSC) and
# Matrix{ComplexF64} are passed as they are; the library transposes CSC -> CSR and narrows to
# Int32 (SURVEY.md §7 hard part 1)
# ---------------------------------------------------------------------------------------
mutable struct DeviceGenerator
    ctx::Context
    handle::Ptr{Cvoid}
    ops::Vector{Ptr{Cvoid}}
    n_coeffs::Int
end

function upload_op(ctx::Context, A::SparseMatrixCSC{ComplexF64,Int64})
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:qp_op_upload_sparse, libqprop), Int32,
                (Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{ComplexF64}, Int32, Int32, Ref{Ptr{Cvoid}}),
                ctx.handle, size(A, 1), size(A, 2), nnz(A), A.colptr, A.rowval, A.nzval, QP_LAYOUT_CSC, 1, h),
          ctx.handle)
    return h[]
end

function upload_op(ctx::Context, A::Matrix{ComplexF64})
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:qp_op_upload_dense, libqprop), Int32, (Ptr{Cvoid}, Int64, Ptr{ComplexF64}, Ref{Ptr{Cvoid}}),
                ctx.handle, size(A, 1), A, h), ctx.handle)
    return h[]
end

"""Matrix-free super-operator ρ ↦ Σ_t c_t P_t ρ Q_t on column-stacked n × n matrices: what
`ham_to_superop` / `lindblad_to_superop` (reference `src/generators.jl:470-508`) build with
Kronecker products, kept as its n × n factors (`nothing` = identity).  A generator whose
`ops` are all `LeftRight` runs on the matrix-free kernel (`QP_FORMAT_LR`); `size` is n² × n², so
it can stand wherever the explicit sparse super-operator stood in a `Generator`."""
struct LeftRight
    n::Int
    terms::Vector{Tuple{Union{Nothing,SparseMatrixCSC{ComplexF64,Int64}},Union{Nothing,SparseMatrixCSC{ComplexF64,Int64}},ComplexF64}}
end
Base.size(A::LeftRight) = (A.n^2, A.n^2)
Base.size(A::LeftRight, d) = A.n^2

"""`liouvillian(Ĥ, c_ops; convention)` without the Kronecker products: same arguments for a
static Hamiltonian matrix; returns a `LeftRight` (use one per term of a time-dependent Ĥ)."""
function liouvillian_matrix_free(H::AbstractMatrix, c_ops=(); convention)
    f, g = convention == :TDSE ? (1.0 + 0im, 1.0im) : convention == :LvN ? (1.0im, 1.0 + 0im) :
           throw(ArgumentError("convention must be :TDSE or :LvN"))
    S(A) = SparseMatrixCSC{ComplexF64,Int64}(sparse(A))
    terms = Any[(S(H), nothing, f), (nothing, S(H), -f)]
    for A in c_ops
        AdA = S(A' * A)
        push!(terms, (S(A), S(A'), g), (AdA, nothing, -g / 2), (nothing, AdA, -g / 2))
    end
    return LeftRight(size(H, 1), [(P, Q, ComplexF64(c)) for (P, Q, c) in terms])
end

function upload_op(ctx::Context, A::LeftRight)
    # the factors are ordinary sparse operators; they must outlive the left/right operator
    left = Ptr{Cvoid}[P === nothing ? C_NULL : upload_op(ctx, P) for (P, _, _) in A.terms]
    right = Ptr{Cvoid}[Q === nothing ? C_NULL : upload_op(ctx, Q) for (_, Q, _) in A.terms]
    coeffs = ComplexF64[c for (_, _, c) in A.terms]
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:qp_op_create_leftright, libqprop), Int32,
                (Ptr{Cvoid}, Int64, Int32, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{ComplexF64}, Ref{Ptr{Cvoid}}),
                ctx.handle, A.n, length(A.terms), left, right, coeffs, h), ctx.handle)
    return h[]
end

upload_op(ctx::Context, A::AbstractSparseMatrix) = upload_op(ctx, SparseMatrixCSC{ComplexF64,Int64}(A))
upload_op(ctx::Context, A::AbstractMatrix) = upload_op(ctx, Matrix{ComplexF64}(A))

"""Device form of a static `Operator` (ops shared with the `Generator` it was evaluated from)."""
function DeviceGenerator(ctx::Context, G::Operator)
    ops = [upload_op(ctx, A) for A in G.ops]
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:qp_gen_create, libqprop), Int32, (Ptr{Cvoid}, Int32, Ptr{Ptr{Cvoid}}, Int32, Int32, Ref{Ptr{Cvoid}}),
                ctx.handle, length(ops), ops, length(G.coeffs), QP_FORMAT_AUTO, h), ctx.handle)
    gen = DeviceGenerator(ctx, h[], ops, length(G.coeffs))
    finalizer(gen) do g
        ccall((:qp_gen_destroy, libqprop), Int32, (Ptr{Cvoid},), g.handle)
        foreach(o -> ccall((:qp_op_destroy, libqprop), Int32, (Ptr{Cvoid},), o), g.ops)
    end
end
DeviceGenerator(ctx::Context, A::AbstractMatrix) = DeviceGenerator(ctx, Operator([A], Float64[]))

_coeffs(G::Operator) = ComplexF64[c for c in G.coeffs]
_coeffs(::AbstractMatrix) = ComplexF64[]

# ---------------------------------------------------------------------------------------
# method = ChebyB200   (mirrors reference src/cheby_propagator.jl:9-27, 87-175, 243-299, 348-386)
# ---------------------------------------------------------------------------------------
mutable struct ChebyB200Propagator{GT,OT} <: PWCPropagator
    const generator::GT
    state::DeviceState
    t::Float64
    n::Int64
    const tlist::Vector{Float64}
    parameters::AbstractDict
    controls
    control_ranges::AbstractDict
    genop::OT
    devgen::DeviceGenerator
    wrk::Ptr{Cvoid}              # qp_cheby_t
    coeffs::Vector{Float64}
    Δ::Float64
    E_min::Float64
    dt::Float64
    limit::Float64
    backward::Bool
    inplace::Bool
    specrange_method::Symbol
    specrange_buffer::Float64
    check_normalization::Bool
    specrange_options::Dict{Symbol,Any}
end

set_t!(p::ChebyB200Propagator, t) = _pwc_set_t!(p, t)

function _set_spectral_range!(p, Δ, E_min, dt)
    p.coeffs = cheby_coeffs(Δ, dt; limit=p.limit)              # host: Bessel functions, src/cheby.jl:25-39
    p.Δ, p.E_min, p.dt = Δ, E_min, dt
    check(ccall((:qp_cheby_set_coeffs, libqprop), Int32,
                (Ptr{Cvoid}, Ptr{Float64}, Int32, Float64, Float64, Float64, Float64),
                p.wrk, p.coeffs, length(p.coeffs), Δ, E_min, abs(dt), p.limit), p.state.ctx.handle)
end

function init_prop(state, generator, tlist, ::Val{:ChebyB200};
                   inplace=true, backward=false, verbose=false, parameters=nothing,
                   control_ranges=nothing, specrange_method=:auto, specrange_buffer=0.01,
                   cheby_coeffs_limit=1e-12, check_normalization=false, uniform_dt_tolerance=1e-12,
                   specrange_kwargs...)
    tlist = convert(Vector{Float64}, tlist)
    controls = get_controls(generator)
    controlvals = [discretize(c, tlist) for c in controls]
    G = _pwc_get_max_genop(generator, controls, tlist)
    parameters = _pwc_process_parameters(parameters, controls, tlist)
    if isnothing(control_ranges)
        control_ranges = IdDict(c => (minimum(controlvals[i]), maximum(controlvals[i]))
                                for (i, c) in enumerate(controls))
    end
    # spectral envelope exactly as the reference (specrange on host operators; :arnoldi can be
    # redirected to the device by passing DeviceState start vectors)
    dstate = state isa DeviceState ? (inplace ? copy(state) : state) : to_device(Vector{ComplexF64}(state))
    ctx = dstate.ctx
    devgen = DeviceGenerator(ctx, G)
    if specrange_method == :arnoldi_device
        # the envelope of the reference (src/cheby_propagator.jl:331-345: extremal control values)
        # with the Arnoldi runs on the device operators that were just uploaded
        lo = ComplexF64[control_ranges[c][1] for c in controls]
        hi = ComplexF64[control_ranges[c][2] for c in controls]
        r1 = device_specrange(devgen, lo, ctx, dstate.n; specrange_kwargs...)
        r2 = device_specrange(devgen, hi, ctx, dstate.n; specrange_kwargs...)
        E_min, E_max = min(r1[1], r2[1]), max(r1[2], r2[2])
    else
        E_min, E_max = cheby_get_spectral_envelope(generator, tlist, control_ranges, specrange_method;
                                                   specrange_kwargs...)
    end
    Δ = E_max - E_min
    @assert Δ > 0.0
    δ = specrange_buffer * Δ
    E_min -= δ / 2
    Δ += δ
    dt = _get_uniform_dt(tlist; tol=uniform_dt_tolerance, warn=true)
    isnothing(dt) && error("Chebychev propagation only works on a uniform time grid")
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:qp_cheby_create, libqprop), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}),
                devgen.handle, dstate.handle, h), ctx.handle)
    n, t = backward ? (length(tlist) - 1, tlist[end]) : (1, tlist[1])
    p = ChebyB200Propagator{typeof(generator),typeof(G)}(
        generator, dstate, t, n, tlist, parameters, controls, control_ranges, G, devgen, h[],
        Float64[], Δ, E_min, dt, cheby_coeffs_limit, backward, inplace, specrange_method,
        specrange_buffer, check_normalization, Dict{Symbol,Any}(specrange_kwargs))
    finalizer(q -> ccall((:qp_cheby_destroy, libqprop), Int32, (Ptr{Cvoid},), q.wrk), p)
    _set_spectral_range!(p, Δ, E_min, dt)
    return p
end

function prop_step!(p::ChebyB200Propagator)
    n = p.n
    tlist = getfield(p, :tlist)
    (0 < n < length(tlist)) || return nothing
    dt = p.backward ? -p.dt : p.dt
    H = _pwc_set_genop!(p, n)                                   # host: rewrites H.coeffs only
    Ψ = p.inplace ? p.state : copy(p.state)
    check(ccall((:qp_cheby_step, libqprop), Int32,                # ONE ccall per prop_step!
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{ComplexF64}, Int32, Float64, Int32),
                p.wrk, Ψ.handle, _coeffs(H), 0, dt, p.check_normalization), Ψ.ctx.handle)
    p.inplace || setfield!(p, :state, Ψ)
    _pwc_advance_time!(p)
    return p.state
end

"""
    expvals, norms = propagate_on_device!(p::ChebyB200Propagator, observables; norms=false)

The step loop of `propagate` (reference `src/propagate.jl:283-344`) as ONE `ccall`
(`qp_cheby_propagate`): all remaining intervals of the grid, the amplitudes of every interval
taken from `p.parameters` up front (legal when no callback can touch them between steps), and
the expectation values `dot(Ψ, O, Ψ)` of the matrix observables (`src/storage.jl:100-123`)
recorded on the device before the first and after every step -- no state is downloaded.
`expvals[k, i]` is observable `k` at grid point `i` in propagation order.
"""
function propagate_on_device!(p::ChebyB200Propagator, observables::Vector{<:AbstractMatrix}=AbstractMatrix[]; norms::Bool=false)
    tlist = getfield(p, :tlist)
    steps = p.backward ? (p.n:-1:1) : (p.n:length(tlist)-1)
    n_c = length(getfield(p, :genop).coeffs)
    table = Matrix{ComplexF64}(undef, n_c, length(steps))      # column-major: [n_steps][n_coeffs] in C order
    for (s, n) in enumerate(steps)
        table[:, s] .= _coeffs(_pwc_set_genop!(p, n))
    end
    ctx = p.state.ctx
    gens = [DeviceGenerator(ctx, O) for O in observables]
    handles = Ptr{Cvoid}[g.handle for g in gens]
    ev = Matrix{ComplexF64}(undef, length(gens), length(steps) + 1)
    nr = Vector{Float64}(undef, norms ? length(steps) + 1 : 0)
    GC.@preserve gens check(ccall((:qp_cheby_propagate, libqprop), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{ComplexF64}, Int32, Int32, Float64, Int32, Ptr{Ptr{Cvoid}}, Ptr{ComplexF64}, Ptr{Float64}),
                p.wrk, p.state.handle, table, 0, length(steps), p.backward ? -p.dt : p.dt, length(gens),
                isempty(gens) ? C_NULL : handles, isempty(gens) ? C_NULL : ev, norms ? nr : C_NULL), ctx.handle)
    for _ in steps
        _pwc_advance_time!(p)
    end
    return ev, nr
end

"""`dot(Ψ, O, Ψ)` of a device-resident state in one fused pass (`qp_gen_expval`)."""
function expval(G::DeviceGenerator, Ψ::DeviceState, coeffs::Vector{ComplexF64}=ComplexF64[])
    out = Ref{ComplexF64}(0)
    check(ccall((:qp_gen_expval, libqprop), Int32, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{Cvoid}, Ref{ComplexF64}),
                G.handle, coeffs, Ψ.handle, out), Ψ.ctx.handle)
    return out[]
end

function reinit_prop!(p::ChebyB200Propagator, state; transform_control_ranges=(c, lo, hi, check) -> (lo, hi), _...)
    set_state!(p, state isa DeviceState ? state : to_device(Vector{ComplexF64}(state); ctx=p.state.ctx))
    ranges = IdDict(c => (minimum(p.parameters[c]), maximum(p.parameters[c])) for c in p.controls)
    recalc = any(p.controls) do c
        lo, hi = transform_control_ranges(c, ranges[c]..., true)
        lo < p.control_ranges[c][1] || hi > p.control_ranges[c][2]
    end
    if recalc
        for c in p.controls
            ranges[c] = transform_control_ranges(c, ranges[c]..., false)
        end
        E_min, E_max = cheby_get_spectral_envelope(getfield(p, :generator), p.tlist, ranges,
                                                   p.specrange_method; p.specrange_options...)
        Δ = E_max - E_min
        δ = p.specrange_buffer * Δ
        p.control_ranges = ranges
        # new coefficient table only -- the operators stay on the device (SURVEY.md §3.4)
        _set_spectral_range!(p, Δ + δ, E_min - δ / 2, float(p.tlist[2] - p.tlist[1]))
    end
    _pwc_set_t!(p, float(p.backward ? p.tlist[end] : p.tlist[1]))
end

# ---------------------------------------------------------------------------------------
# method = NewtonB200   (mirrors reference src/newton_propagator.jl and src/newton.jl:246-385)
# vectors on the device, Hessenberg / Leja / divided differences on the host
# ---------------------------------------------------------------------------------------
mutable struct NewtonB200Propagator{GT,OT} <: PWCPropagator
    const generator::GT
    state::DeviceState
    t::Float64
    n::Int64
    const tlist::Vector{Float64}
    parameters::AbstractDict
    controls
    genop::OT
    devgen::DeviceGenerator
    krylov::Ptr{Cvoid}           # qp_krylov_t
    v::DeviceState
    m_max::Int64
    backward::Bool
    inplace::Bool
    func::Function
    norm_min::Float64
    relerr::Float64
    max_restarts::Int64
end

set_t!(p::NewtonB200Propagator, t) = _pwc_set_t!(p, t)

function init_prop(state, generator, tlist, ::Val{:NewtonB200};
                   inplace=true, backward=false, verbose=false, parameters=nothing, m_max=10,
                   func=(z -> exp(-1im * z)), norm_min=1e-14, relerr=1e-12, max_restarts=50, _...)
    inplace || error("The Newton propagator is only implemented in-place")
    tlist = convert(Vector{Float64}, tlist)
    controls = get_controls(generator)
    G = _pwc_get_max_genop(generator, controls, tlist)
    parameters = _pwc_process_parameters(parameters, controls, tlist)
    dstate = state isa DeviceState ? copy(state) : to_device(Vector{ComplexF64}(state))
    m_max = min(m_max, length(dstate) - 1)
    m_max > 2 || error("Newton propagation requires m_max > 2")
    devgen = DeviceGenerator(dstate.ctx, G)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:qp_krylov_create, libqprop), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ref{Ptr{Cvoid}}),
                devgen.handle, dstate.handle, m_max, h), dstate.ctx.handle)
    n, t = backward ? (length(tlist) - 1, tlist[end]) : (1, tlist[1])
    p = NewtonB200Propagator{typeof(generator),typeof(G)}(
        generator, dstate, t, n, tlist, parameters, controls, G, devgen, h[], similar(dstate), m_max,
        backward, inplace, func, norm_min, relerr, max_restarts)
    finalizer(q -> ccall((:qp_krylov_destroy, libqprop), Int32, (Ptr{Cvoid},), q.krylov), p)
end

function _combine!(p, w::Vector{ComplexF64}, st::DeviceState, accumulate::Bool)
    check(ccall((:qp_krylov_combine, libqprop), Int32, (Ptr{Cvoid}, Ptr{ComplexF64}, Int32, Int32, Ptr{Cvoid}, Int32),
                p.krylov, w, 0, length(w), st.handle, accumulate), st.ctx.handle)
end

function prop_step!(p::NewtonB200Propagator)
    n = p.n
    tlist = getfield(p, :tlist)
    (0 < n < length(tlist)) || return nothing
    dt = tlist[n+1] - tlist[n]
    p.backward && (dt = -dt)
    H = _pwc_set_genop!(p, n)
    Ψ, ctx, func = p.state, p.state.ctx, p.func
    m = p.m_max
    a = OffsetVector(zeros(ComplexF64, 10 * m + 1), 0:(10*m))
    leja = OffsetVector(zeros(ComplexF64, 10 * m + 1), 0:(10*m))
    Hess = zeros(ComplexF64, m + 1, m + 1)
    n_a = n_leja = s = 0
    radius = 0.0
    copyto!(p.v, Ψ)
    β = norm(p.v)
    lmul!(1 / β, p.v)
    while true
        m_out = Ref{Int32}(0)
        check(ccall((:qp_arnoldi, libqprop), Int32,            # Arnoldi sweep on the device
                    (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{Cvoid}, Int32, Float64, Int32, Float64, Ptr{ComplexF64}, Int32, Ref{Int32}),
                    p.krylov, _coeffs(H), p.v.handle, m, dt, true, p.norm_min, Hess, size(Hess, 1), m_out), ctx.handle)
        m = Int(m_out[])
        if m == 1 && s == 0
            lmul!(func(β * Hess[1, 1]), Ψ)
            break
        end
        ritz = diagonalize_hessenberg_matrix(Hess, m, accumulate=true)        # host, src/arnoldi.jl:143
        s == 0 && (radius = leja_radius(ritz))
        n_s = n_leja
        n_leja = extend_leja!(leja, n_leja, OffsetVector(ritz, 0:(length(ritz)-1)), m)   # host
        n_a = extend_newton_coeffs!(a, n_a, leja, func, n_leja, radius)                  # host
        R = zeros(ComplexF64, m + 1)
        R[1] = β
        P = a[n_s] * R
        Hm = @view Hess[1:(m+1), 1:(m+1)]
        for k = 1:(m-1)
            R = (Hm * R - leja[n_s+k-1] * R) / radius
            P += a[n_s+k] * R
        end
        _combine!(p, P[1:m], Ψ, s > 0)                         # Ψ (+)= Σ P_i q_i on the device
        R = (Hm * R - leja[n_s+m-1] * R) / radius
        β = norm(R)
        _combine!(p, R / β, p.v, false)                        # restart vector on the device
        (β * abs(a[n_a-1]) / (1 + norm(Ψ)) < p.relerr) && break
        s += 1
        @assert s <= p.max_restarts
    end
    _pwc_advance_time!(p)
    return p.state
end

# ---------------------------------------------------------------------------------------
# Device spectral range (SURVEY.md §8f-2): `specrange(H; method=:arnoldi)` (reference
# src/specrad.jl:88-112, 170-220) with the Krylov vectors on the GPU -- the Ritz values of an
# m-step Arnoldi run started from a random state, enlarged by the usual buffer.  Use it through
# `init_prop(...; method=ChebyB200, specrange_method=:arnoldi_device)`; the host route of the
# reference (`cheby_get_spectral_envelope` on SparseMatrixCSC operators) stays the default.
# ---------------------------------------------------------------------------------------
function device_specrange(devgen::DeviceGenerator, coeffs::Vector{ComplexF64}, ctx::Context, n::Integer;
                          m_max::Integer=60, enlarge::Bool=true, norm_min=1e-15)
    Ψ = randn(ComplexF64, n)
    v = to_device(Ψ / norm(Ψ); ctx=ctx)
    m = min(m_max, n - 1)
    K = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:qp_krylov_create, libqprop), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ref{Ptr{Cvoid}}),
                devgen.handle, v.handle, m, K), ctx.handle)
    Hess = zeros(ComplexF64, m + 1, m + 1)
    m_out = Ref{Int32}(0)
    try
        check(ccall((:qp_arnoldi, libqprop), Int32,
                    (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{Cvoid}, Int32, Float64, Int32, Float64, Ptr{ComplexF64}, Int32, Ref{Int32}),
                    K[], coeffs, v.handle, m, 1.0, false, norm_min, Hess, size(Hess, 1), m_out), ctx.handle)
    finally
        ccall((:qp_krylov_destroy, libqprop), Int32, (Ptr{Cvoid},), K[])
    end
    mm = Int(m_out[])
    ritz = sort(real.(eigvals(Hess[1:mm, 1:mm])))
    E_min, E_max = ritz[1], ritz[end]
    if enlarge && mm > 1                      # src/specrad.jl:103-107
        E_min -= ritz[2] - ritz[1]
        E_max += ritz[end] - ritz[end-1]
    end
    return E_min, E_max
end

# ---------------------------------------------------------------------------------------
# Ensembles of trajectories on all GPUs of a box (C ABI group `qp_ens_*`, SURVEY.md §8b / §8e):
# trajectory b scales control l by `scales[l, b]`; the trajectories are cut into contiguous
# blocks, one per GPU; every GPU steps its [N][B_local] batched state with ONE qp_cheby_step per
# time interval (per-trajectory coefficients); nothing is communicated inside the time loop, and
# the final gathers run over NCCL/NVLink inside the library.  The shared spectral envelope is
# the `control_ranges` hook of the reference (src/cheby_propagator.jl:59-66) evaluated over the
# whole ensemble.
# ---------------------------------------------------------------------------------------
mutable struct BatchedState
    ctx::Context
    handle::Ptr{Cvoid}
    n::Int64
    batch::Int64
    function BatchedState(ctx::Context, n::Integer, batch::Integer)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:qp_state_create, libqprop), Int32, (Ptr{Cvoid}, Int64, Int64, Ref{Ptr{Cvoid}}),
                    ctx.handle, n, batch, h), ctx.handle)
        st = new(ctx, h[], n, batch)
        finalizer(s -> ccall((:qp_state_destroy, libqprop), Int32, (Ptr{Cvoid},), s.handle), st)
    end
end

# host layout of a batched state: Matrix{ComplexF64} of size (batch, N) -- trajectory index fastest,
# i.e. the library's [N][B] row-major layout seen by column-major Julia
function upload!(st::BatchedState, Ψ::Matrix{ComplexF64})
    @assert size(Ψ) == (st.batch, st.n)
    check(ccall((:qp_state_upload, libqprop), Int32, (Ptr{Cvoid}, Ptr{ComplexF64}, Int64, Int64),
                st.handle, Ψ, 0, st.batch), st.ctx.handle)
    return st
end

mutable struct EnsembleMember
    ctx::Context
    rank::Int
    b0::Int
    b1::Int
    gen::DeviceGenerator
    state::BatchedState
    wrk::Ptr{Cvoid}
end

mutable struct EnsembleB200
    handle::Ptr{Cvoid}
    members::Vector{EnsembleMember}
    n_total::Int
    tlist::Vector{Float64}
    coeffs::Array{ComplexF64,3}     # (B_total, L, nt-1): u[b, l, n] = scales[l, b] * ε_l(t_n) on the midpoints
    n::Int
    dt::Float64
end

"""
    EnsembleB200(devices, H::Generator, scales, Ψ₀, tlist; E_min, E_max, specrange_buffer=0.01, limit=1e-12)

`devices`: CUDA device indices (all distinct: NCCL; all equal: several ranks on one GPU).
`scales`: `L × B_total` matrix of amplitude scales.  `Ψ₀`: `Vector` (shared) or `B_total × N` matrix.
"""
function EnsembleB200(devices::Vector{<:Integer}, H::Generator, scales::Matrix{Float64}, Ψ₀, tlist;
                      E_min::Float64, E_max::Float64, specrange_buffer=0.01, limit=1e-12)
    tlist = convert(Vector{Float64}, tlist)
    controls = get_controls(H)
    L, B = size(scales)
    @assert L == length(controls)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:qp_ens_create, libqprop), Int32, (Ptr{Int32}, Int32, Ref{Ptr{Cvoid}}),
                Int32.(devices), length(devices), h))
    mid = [QuantumPropagators.Controls.discretize_on_midpoints(c, tlist) for c in controls]
    coeffs = [ComplexF64(scales[l, b] * mid[l][n]) for b = 1:B, l = 1:L, n = 1:(length(tlist)-1)]
    Δ = E_max - E_min
    δ = specrange_buffer * Δ
    dt = tlist[2] - tlist[1]
    a = cheby_coeffs(Δ + δ, dt; limit=limit)
    G = _pwc_get_max_genop(H, controls, tlist)
    N = size(G.ops[1], 1)
    members = EnsembleMember[]
    for i = 0:(length(devices)-1)
        ch, rk = Ref{Ptr{Cvoid}}(C_NULL), Ref{Int32}(0)
        check(ccall((:qp_ens_ctx, libqprop), Int32, (Ptr{Cvoid}, Int32, Ref{Ptr{Cvoid}}, Ref{Int32}), h[], i, ch, rk))
        ctx = Context(ch[], false)             # borrowed: owned by the ensemble
        b0, b1 = Ref{Int64}(0), Ref{Int64}(0)
        check(ccall((:qp_ens_shard, libqprop), Int32, (Int64, Int32, Int32, Ref{Int64}, Ref{Int64}),
                    B, rk[], length(devices), b0, b1))
        gen = DeviceGenerator(ctx, G)
        st = BatchedState(ctx, N, b1[] - b0[])
        local_Ψ = Ψ₀ isa AbstractVector ? repeat(transpose(ComplexF64.(Ψ₀)), b1[] - b0[], 1) :
                  Matrix{ComplexF64}(Ψ₀[(b0[]+1):b1[], :])
        upload!(st, local_Ψ)
        w = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:qp_cheby_create, libqprop), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), gen.handle, st.handle, w), ctx.handle)
        check(ccall((:qp_cheby_set_coeffs, libqprop), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int32, Float64, Float64, Float64, Float64),
                    w[], a, length(a), Δ + δ, E_min - δ / 2, abs(dt), limit), ctx.handle)
        push!(members, EnsembleMember(ctx, rk[], b0[], b1[], gen, st, w[]))
    end
    ens = EnsembleB200(h[], members, B, tlist, coeffs, 1, dt)
    finalizer(ens) do e
        foreach(m -> ccall((:qp_cheby_destroy, libqprop), Int32, (Ptr{Cvoid},), m.wrk), e.members)
        ccall((:qp_ens_destroy, libqprop), Int32, (Ptr{Cvoid},), e.handle)
    end
    return ens
end

"One `prop_step!` of every trajectory: one batched `qp_cheby_step` per GPU, enqueued back to back (the GPUs run concurrently)."
function prop_step!(ens::EnsembleB200)
    (0 < ens.n < length(ens.tlist)) || return nothing
    for m in ens.members
        c = ens.coeffs[(m.b0+1):m.b1, :, ens.n]                 # (B_local, L): [l][b] with b fastest
        check(ccall((:qp_cheby_step, libqprop), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{ComplexF64}, Int32, Float64, Int32),
                    m.wrk, m.state.handle, c, 1, ens.dt, 0), m.ctx.handle)
    end
    ens.n += 1
    return ens
end

function ensemble_propagate!(ens::EnsembleB200)
    while !isnothing(prop_step!(ens)) end
    return ens
end

"Final states of the whole ensemble as a `B_total × N` matrix (NCCL gather inside the library)."
function gather_states(ens::EnsembleB200)
    N = ens.members[1].state.n
    out = Matrix{ComplexF64}(undef, ens.n_total, N)
    hs = [m.state.handle for m in ens.members]
    check(ccall((:qp_ens_gather_states, libqprop), Int32, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Int64, Ptr{Ptr{Cvoid}}, Ptr{ComplexF64}),
                ens.handle, hs, ens.n_total, C_NULL, out), ens.members[1].ctx.handle)
    return out
end

"Per-trajectory numbers: `values[i]` is member i's `B_local × n_values` matrix; returns `B_total × n_values`."
function gather_expvals(ens::EnsembleB200, values::Vector{Matrix{ComplexF64}})
    nv = size(values[1], 2)
    out = Matrix{ComplexF64}(undef, ens.n_total, nv)
    ptrs = [pointer(v) for v in values]
    GC.@preserve values check(ccall((:qp_ens_gather_expvals, libqprop), Int32,
                                    (Ptr{Cvoid}, Ptr{Ptr{ComplexF64}}, Int32, Int64, Ptr{ComplexF64}),
                                    ens.handle, ptrs, nv, ens.n_total, out), ens.members[1].ctx.handle)
    return out
end

end # module
