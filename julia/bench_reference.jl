# Times the REAL reference package on BASELINE configs[1] (TFIM chain, H0 + 2 PWC controls,
# Cheby prop_step!) -- the procedure of BASELINE.md §2.1 / docs/src/benchmarks/profiling.md.
# Called by `bench.py --impl reference` when a Julia install is found (PATH or baseline/_ref);
# prints one JSON line.  NOT executed in this image (no Julia): bench.py falls back to the C
# restatement oracle/cheby_ref.c when this script cannot run.
#
#   JULIA_NUM_THREADS=1 julia julia/bench_reference.jl <n_spins> <steps> <warmup>
using LinearAlgebra, SparseArrays
using QuantumPropagators
using QuantumPropagators: init_prop, prop_step!, Cheby
using QuantumPropagators.Generators: hamiltonian

n = parse(Int, get(ARGS, 1, "20"))
steps = parse(Int, get(ARGS, 2, "20"))
warmup = parse(Int, get(ARGS, 3, "5"))
BLAS.set_num_threads(Sys.CPU_THREADS)

N = 1 << n
idx = collect(0:(N-1))
z = [1.0 .- 2.0 .* ((idx .>> i) .& 1) for i = 0:(n-1)]
d0 = zeros(N); for i = 1:(n-1); d0 .-= z[i] .* z[i+1]; end
d2 = zeros(N); for i = 1:n; d2 .+= z[i]; end
H0 = spdiagm(0 => ComplexF64.(d0))
H2 = spdiagm(0 => ComplexF64.(d2))
rows = repeat(idx .+ 1, n)
cols = vcat([(idx .⊻ (1 << i)) .+ 1 for i = 0:(n-1)]...)
H1 = sparse(rows, cols, ones(ComplexF64, N * n), N, N)

T = 0.1 * 100
tlist = collect(range(0.0, T, length = 101))
u1(t) = sin(pi * t / T)^2
u2(t) = 0.5 * sin(4pi * t / T)
H = hamiltonian(H0, (H1, u1), (H2, u2))
bound = (n - 1) + n + 0.5n
using Random
Random.seed!(2000)
psi = randn(ComplexF64, N); psi ./= norm(psi)

p = init_prop(psi, H, tlist; method = Cheby, E_min = -bound, E_max = bound)
for _ = 1:max(warmup, 1); prop_step!(p); end
t = @elapsed for _ = 1:steps; prop_step!(p); end
println("{\"prop_steps_per_s\": $(steps / t), \"n_coeffs\": $(p.wrk.n_coeffs), \"blas_threads\": $(BLAS.get_num_threads()), \"julia_threads\": $(Threads.nthreads())}")
