/*
 * qprop.h -- C ABI of libqprop_b200: B200-native (sm_100a) Chebyshev and Newton/Arnoldi
 * propagation kernels behind QuantumPropagators.jl's propagator API.
 *
 * This is the drop-in boundary.  The reference has no FFI: its method
 * plugin interface is Julia multiple dispatch (init_prop/prop_step!/...), and the
 * numerical kernels below it (module Cheby, Newton, Arnoldi, SpectralRange, and the
 * Operator mul!) are what each entry point here replaces.  The Julia host
 * (julia/QPropB200.jl) binds these with `ccall`; the Python host
 * mirror (quantumpropagators.jl_b200/) binds them with ctypes.
 *
 * Conventions
 *   - every function returns an int32 status (QP_OK == 0, negative == error class);
 *     qp_last_error(ctx) gives the message.  No exceptions cross the ABI.
 *   - host pointers are borrowed for the duration of the call only; the library copies.
 *   - handles are opaque and freed explicitly.  Calls on different contexts are
 *     thread-safe; calls on one context are not (one CUDA stream per context).
 *   - complex numbers are (re, im) pairs of float64 (Julia ComplexF64 / numpy complex128).
 *   - a state holds B >= 1 vectors of dimension N, stored [N][B] with the batch index
 *     fastest ("trajectory-batched"); B == 1 is a plain vector.
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with
 *     QP_ERR_CUDA.
 */
#ifndef QPROP_H
#define QPROP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QP_VERSION 100 /* 0.1.0 */

/* status codes */
#define QP_OK 0
#define QP_ERR_INVALID_ARG -1   /* Julia: ArgumentError / AssertionError on sizes */
#define QP_ERR_CUDA -2          /* CUDA runtime / launch failure, no device */
#define QP_ERR_OOM -3           /* device or host allocation failed */
#define QP_ERR_NOT_CONVERGED -4 /* `@assert s <= max_restarts`, src/newton.jl:375 */
#define QP_ERR_NORMALIZATION -5 /* "Incorrect normalization", src/cheby.jl:196 */
#define QP_ERR_UNSUPPORTED -6   /* valid request outside the engine's limits */
#define QP_ERR_INTERNAL -7

/* sparse layouts accepted by qp_op_upload_sparse */
#define QP_LAYOUT_CSC 0 /* Julia SparseMatrixCSC: colptr/rowval/nzval */
#define QP_LAYOUT_CSR 1

/* device storage formats of a generator (qp_gen_create / qp_gen_info) */
#define QP_FORMAT_AUTO 0
#define QP_FORMAT_CSR 1   /* merged multi-operator CSR, sub-warp per row        */
#define QP_FORMAT_SELL 2  /* merged sliced-ELL (C = 32), thread per row           */
#define QP_FORMAT_DENSE 3 /* dense row-major ComplexF64 operators                 */
#define QP_FORMAT_SELLD 4 /* dictionary-compressed sliced-ELL: one 8/16-bit code per entry
                             indexing a table of distinct (operator, value, column-row)
                             triples; chosen when the table has < 4096 entries            */
#define QP_FORMAT_LR 5    /* matrix-free left/right products on column-stacked matrices
                             (qp_op_create_leftright); chosen automatically for such operators */
#define QP_FORMAT_BITFLIP 6 /* no matrix stream: every operator is a diagonal (a vector or 16-bit codes)
                             plus uniform bit flips -- row r couples to r XOR mask_t with the same value
                             v_t, on all rows or on the sub-cube (r AND cmask_t) == cval_t (sums of
                             Pauli-X strings, sigma+/sigma- terms, Liouvillians of such systems).
                             Detected from the uploaded matrices; single states use it, state batches
                             of such a generator keep the tiled / dictionary kernels */

typedef struct qp_ctx_s* qp_ctx_t;
typedef struct qp_op_s* qp_op_t;
typedef struct qp_gen_s* qp_gen_t;
typedef struct qp_state_s* qp_state_t;
typedef struct qp_cheby_s* qp_cheby_t;
typedef struct qp_krylov_s* qp_krylov_t;
typedef struct qp_ens_s* qp_ens_t;

typedef struct {
  double re, im;
} qp_c128;

/* ------------------------------------------------------------------ library / context */

int32_t qp_version(void);
const char* qp_status_string(int32_t status);

/* One context = one device + one CUDA stream + error/timer state. */
int32_t qp_ctx_create(int32_t device, qp_ctx_t* ctx);
int32_t qp_ctx_destroy(qp_ctx_t ctx);
int32_t qp_sync(qp_ctx_t ctx);
const char* qp_last_error(qp_ctx_t ctx);
/* raw cudaStream_t of the context (for event timing / interop by the host) */
int32_t qp_ctx_stream(qp_ctx_t ctx, void** stream);
/* number of kernels this library has launched on the context since creation */
int32_t qp_ctx_launch_count(qp_ctx_t ctx, int64_t* n_launches);

/* Timers under the reference's TimerOutputs labels (src/timings.jl; labels
 * "prop_step!", "matrix-vector product", "arnoldi!", src/cheby.jl:175,
 * src/arnoldi.jl:81, src/cheby_propagator.jl:349 -- CUDA-event based -- and the host-side
 * sections of newton!, "diagonalize_hessenberg_matrix", "get Leja points",
 * "get Newton coeffs", "evaluate polynomial", src/newton.jl:297-343 -- wall clock).
 * Off by default. */
int32_t qp_timer_enable(qp_ctx_t ctx, int32_t on);
int32_t qp_timer_get(qp_ctx_t ctx, const char* label, int64_t* ncalls, double* seconds);
int32_t qp_timer_reset(qp_ctx_t ctx);

/* ------------------------------------------------------------------ operators
 * Replaces: the component operators `A.ops[i]` of a reference `Operator`
 * (src/generators.jl:111-125), i.e. SparseMatrixCSC{ComplexF64,Int64} or
 * Matrix{ComplexF64}.  Sparse input is converted to device CSR with Int32 indices
 * (CSC input is transposed: the arrays of a CSC matrix are the CSR arrays of its
 * transpose -- SURVEY.md §7 hard part 1). */
int32_t qp_op_upload_sparse(qp_ctx_t ctx, int64_t nrows, int64_t ncols, int64_t nnz,
                            const int64_t* ptr, const int64_t* idx, const qp_c128* val,
                            int32_t layout, int32_t index_base, qp_op_t* op);
/* column-major (Julia Matrix) n x n */
int32_t qp_op_upload_dense(qp_ctx_t ctx, int64_t n, const qp_c128* colmajor, qp_op_t* op);
/* Matrix-free super-operator on column-stacked n x n matrices (vec index i + n*j):
 *     rho  ->  sum_t  c_t * P_t rho Q_t ,     P_t / Q_t sparse n x n operators or NULL (= identity).
 * Replaces the explicit sparse super-operators the reference builds with Kronecker products in
 * ham_to_superop / lindblad_to_superop (src/generators.jl:470-508; vec(P rho Q) = (Q^T (x) P) vec rho):
 *     1 (x) H  = (H, NULL),   H^T (x) 1 = (NULL, H),   (A^+)^T (x) A = (A, A^+),
 * so a Liouvillian of an n-dimensional system costs O(nnz(H)) memory instead of O(n * nnz(H)).
 * The result is an operator of dimension n^2; a generator is either made of such operators only
 * (format QP_FORMAT_LR -- or QP_FORMAT_BITFLIP when the factors are diagonals plus bit flips that compose,
 * which QP_FORMAT_AUTO detects -- single states) or of matrices only.  `left` / `right` must be sparse
 * operators of this context and outlive the new operator. */
int32_t qp_op_create_leftright(qp_ctx_t ctx, int64_t n, int32_t n_terms, const qp_op_t* left,
                               const qp_op_t* right, const qp_c128* coeffs, qp_op_t* op);
int32_t qp_op_destroy(qp_op_t op);
int32_t qp_op_info(qp_op_t op, int64_t* nrows, int64_t* ncols, int64_t* nnz, int32_t* is_dense);

/* Lazy sum H = sum_{l<drift} ops[l] + sum_l c_l ops[drift+l], drift = n_ops - n_coeffs
 * (src/generators.jl:111-125, 634-636).  The operators are shared, immutable, and must
 * outlive the generator; only the n_coeffs numbers change from step to step
 * (evaluate!, src/generators.jl:757-766).  `format` is QP_FORMAT_AUTO or a forced one. */
int32_t qp_gen_create(qp_ctx_t ctx, int32_t n_ops, const qp_op_t* ops, int32_t n_coeffs,
                      int32_t format, qp_gen_t* gen);
int32_t qp_gen_destroy(qp_gen_t gen);
/* format chosen, total stored entries (incl. padding), algorithmic matrix bytes M
 * (sum_ops 20*nnz + 4*(N+1), or 16 N^2 per dense operator) */
int32_t qp_gen_info(qp_gen_t gen, int32_t* format, int64_t* n, int64_t* stored_entries,
                    int64_t* matrix_bytes);
/* what one application of the generator actually streams from memory for the matrix:
 * bytes of the chosen device format (incl. padding); n_dict / code_bytes: entries of the
 * dictionary of distinct (operator, value, offset) triples and the code width, when one could
 * be built (used by QP_FORMAT_SELLD and by the trajectory-batched kernel; 0 otherwise).  No reference counterpart:
 * the reference stores SparseMatrixCSC{ComplexF64,Int64} (24 B per entry). */
int32_t qp_gen_storage(qp_gen_t gen, int64_t* stored_bytes, int32_t* n_dict, int32_t* code_bytes);

/* The two-pass tiled form of the generator used for trajectory-batched states (batch a multiple of
 * 32; csrc/tile_format.h): rows split as r = hi * split + lo, entries served from a shared-memory
 * tile of the state when row and column lie in the same block (class A) or at the same position of
 * different blocks (class B).  Built on first use.  available = 0: the generator does not qualify
 * (dense / matrix-free, more than 3 operators, complex-valued operators, no split of N with tiles
 * of <= 256 rows, or more than a quarter of the entries in neither class) and batched states use
 * the one-pass kernels.  entries[4] = merged entries of class A, B, other, diagonal.  No reference
 * counterpart (the reference applies one SparseMatrixCSC per operator and state,
 * src/generators.jl:634-645). */
int32_t qp_gen_tile_info(qp_gen_t gen, int32_t* available, int32_t* split, int32_t* blocks, int32_t* n_table,
                         int64_t* entries);

/* ------------------------------------------------------------------ states
 * Replaces: Vector{ComplexF64} states and the level-1 verbs the reference's kernels use
 * on them (src/interfaces/state.jl:24-47; call sites src/cheby.jl:171-211,
 * src/arnoldi.jl:79-96, src/newton.jl:268-367). */
int32_t qp_state_create(qp_ctx_t ctx, int64_t n, int64_t batch, qp_state_t* st);
int32_t qp_state_destroy(qp_state_t st);
int32_t qp_state_info(qp_state_t st, int64_t* n, int64_t* batch);
/* current device pointer of the [n][batch] ComplexF64 buffer (may change after
 * qp_cheby_step, which swaps buffers instead of copying) */
int32_t qp_state_devptr(qp_state_t st, void** devptr);
/* host layout [n][nb] (batch fastest), columns b0 .. b0+nb-1 of the state */
int32_t qp_state_upload(qp_state_t st, const qp_c128* host, int64_t b0, int64_t nb);
int32_t qp_state_download(qp_state_t st, qp_c128* host, int64_t b0, int64_t nb);
/* The same transfers enqueued on the context's stream WITHOUT waiting for them: the host buffer
 * (page-locked, or the copy degrades to a synchronous one) must stay untouched until qp_sync.
 * Lets a host that keeps its states in host memory (the reference's Vector{ComplexF64} states,
 * src/propagate.jl:283-344) overlap the copies of one trajectory with the step of another
 * (one context = one stream per trajectory). */
int32_t qp_state_upload_async(qp_state_t st, const qp_c128* host, int64_t b0, int64_t nb);
int32_t qp_state_download_async(qp_state_t st, qp_c128* host, int64_t b0, int64_t nb);

int32_t qp_copy(qp_state_t dst, qp_state_t src);                    /* copyto!(dst, src) */
int32_t qp_fill(qp_state_t st, qp_c128 value);                      /* fill!            */
int32_t qp_scal(qp_state_t st, qp_c128 alpha);                      /* lmul!(alpha, st)  */
int32_t qp_axpy(qp_c128 alpha, qp_state_t x, qp_state_t y);         /* axpy!(alpha,x,y)  */
int32_t qp_dot(qp_state_t x, qp_state_t y, qp_c128* out /*[batch]*/);  /* dot(x,y) = x^H y  */
int32_t qp_norm(qp_state_t x, double* out /*[batch]*/);             /* norm(x) per column */

/* y <- beta*y + alpha * (sum_l c_l H_l) x : mul!(C, A::Operator, B, alpha, beta),
 * src/generators.jl:634-645; `coeffs` has n_coeffs entries (shared by the batch). */
int32_t qp_gen_mul(qp_gen_t gen, const qp_c128* coeffs, qp_c128 alpha, qp_c128 beta,
                   qp_state_t x, qp_state_t y);
/* out[b] = <x_b| sum_l c_l H_l |y_b> : dot(x, A::Operator, y), src/generators.jl:648-660 */
int32_t qp_gen_dot(qp_gen_t gen, const qp_c128* coeffs, qp_state_t x, qp_state_t y,
                   qp_c128* out /*[batch]*/);

/* out[b] = <x_b| sum_l c_l H_l |x_b> in one fused pass (no temporary, no stored H x):
 * expectation values of matrix observables, src/storage.jl:100-123 */
int32_t qp_gen_expval(qp_gen_t gen, const qp_c128* coeffs, qp_state_t x, qp_c128* out /*[batch]*/);

/* ------------------------------------------------------------------ Chebyshev
 * Replaces: ChebyWrk (src/cheby.jl:87-124) and cheby! (src/cheby.jl:150-213).  The
 * coefficient table a_k comes from the host (cheby_coeffs, src/cheby.jl:25-39). */
int32_t qp_cheby_create(qp_gen_t gen, qp_state_t like, qp_cheby_t* wrk);
int32_t qp_cheby_destroy(qp_cheby_t wrk);
/* (re-)upload coefficients without touching the operators (reinit_prop!,
 * src/cheby_propagator.jl:243-299); `limit` is ChebyWrk.limit (threshold of the
 * normalization check, src/cheby.jl:164,197) */
int32_t qp_cheby_set_coeffs(qp_cheby_t wrk, const double* a, int32_t n_a, double Delta,
                            double E_min, double dt_abs, double limit);
/* One prop_step!: st <- exp(-i H dt) st with H = sum_l c_l H_l.
 *   op_coeffs: n_coeffs numbers if coeffs_per_traj == 0 (shared by all trajectories),
 *              else [n_coeffs][batch] (trajectory b uses op_coeffs[l*batch + b]).
 *   dt_signed: +-dt_abs (backward propagation negates, src/cheby_propagator.jl:353-356);
 *              |dt| must match dt_abs of qp_cheby_set_coeffs (src/cheby.jl:157).
 *   check_normalization: src/cheby.jl:194-200; failing returns QP_ERR_NORMALIZATION. */
int32_t qp_cheby_step(qp_cheby_t wrk, qp_state_t st, const qp_c128* op_coeffs,
                      int32_t coeffs_per_traj, double dt_signed, int32_t check_normalization);
/* The step loop of propagate (src/propagate.jl:283-344) in ONE call: n_steps consecutive
 * prop_step!s with the amplitudes of every interval given up front (the caller opts in to
 * this: nothing can change `parameters` between the steps of one propagate call without a
 * callback), and -- instead of downloading the state for storage -- the expectation values
 * <psi|O_k|psi> of n_obs observables (generators without free coefficients; matrix
 * observables, src/storage.jl:100-123) and the norms recorded on the device before the first
 * and after every step.
 *   coeff_table: [n_steps][n_coeffs] (coeffs_per_traj == 0) or [n_steps][n_coeffs][batch]
 *   expvals:     [n_steps + 1][n_obs][batch] or NULL when n_obs == 0
 *   norms:       [n_steps + 1][batch] or NULL
 * No host synchronisation happens between the steps. */
int32_t qp_cheby_propagate(qp_cheby_t wrk, qp_state_t st, const qp_c128* coeff_table,
                           int32_t coeffs_per_traj, int32_t n_steps, double dt_signed, int32_t n_obs,
                           const qp_gen_t* observables, qp_c128* expvals, double* norms);
/* algorithmic bytes of one prop_step! as defined in SURVEY.md §8d:
 * (n_a - 1) * (M + 80 N B) */
int32_t qp_cheby_step_bytes(qp_cheby_t wrk, int64_t* bytes);

/* ------------------------------------------------------------------ Arnoldi / Newton
 * Replaces: arnoldi! (src/arnoldi.jl:60-100), extend_arnoldi! (:115-129) and the vector
 * work of newton! (src/newton.jl:346-367).  The small dense step (Ritz values, Leja
 * points, divided differences, polynomial in the Hessenberg matrix; src/newton.jl:297-343)
 * stays on the host, where `func` is an arbitrary closure.  A Krylov workspace holds m_max+1
 * vectors of the shape of `like`.  batch == 1: classical Gram-Schmidt with a fused multi-dot and
 * selective re-orthogonalisation, enqueued without host round trips.  batch > 1 (a bundle of
 * states sharing the generator, e.g. the forward / backward states of GRAPE and Krotov, reference
 * src/cheby_propagator.jl:147-152, docs/src/overview.md:193-195): the B Arnoldi processes run in
 * lock step with the reference's own modified Gram-Schmidt, every state with its own Hessenberg
 * matrix:  hess = [batch][ld * ld] column-major matrices, m_out = [batch] dimensions, weights of
 * qp_krylov_combine = [n_w][batch]. */
int32_t qp_krylov_create(qp_gen_t gen, qp_state_t like, int32_t m_max, qp_krylov_t* K);
int32_t qp_krylov_destroy(qp_krylov_t K);
/* q_1 <- v; for j = 1..m: q_{j+1} = H q_j, orthogonalised against q_1..q_j; fills the
 * column-major (ld x ld) host matrix `hess` exactly like the reference (entries scaled by
 * dt, zero elsewhere).  Returns the possibly reduced dimension in *m_out. */
int32_t qp_arnoldi(qp_krylov_t K, const qp_c128* op_coeffs, qp_state_t v, int32_t m, double dt,
                   int32_t extended, double norm_min, qp_c128* hess, int32_t ld, int32_t* m_out);
/* extend an (m-1)x(m-1) non-extended decomposition to m x m; *extended_out = 0 if the
 * Krylov space was exhausted (reference returns early, src/arnoldi.jl:116-117). */
int32_t qp_arnoldi_extend(qp_krylov_t K, const qp_c128* op_coeffs, int32_t m, double dt,
                          double norm_min, qp_c128* hess, int32_t ld, int32_t* extended_out);
/* st <- (accumulate ? st : 0) + sum_{i<n_w} w[i] q_{first+i}   (src/newton.jl:346-352) */
int32_t qp_krylov_combine(qp_krylov_t K, const qp_c128* w, int32_t first, int32_t n_w,
                          qp_state_t st, int32_t accumulate);
/* copy Krylov vector q_{index} (0-based) to/from a state.  (Single states: the workspace keeps a new vector
 * unnormalised and its factor 1/|w| apart -- applied wherever the vector is used -- so that no rescaling pass
 * runs per column; qp_krylov_get returns the normalised q_{index}, qp_krylov_set stores a vector with factor 1.) */
int32_t qp_krylov_get(qp_krylov_t K, int32_t index, qp_state_t dst);
int32_t qp_krylov_set(qp_krylov_t K, int32_t index, qp_state_t src);

/* newton! (src/newton.jl:246-385) as ONE call: psi <- func(H dt) psi by restarted Newton
 * interpolation at Leja-ordered Ritz values.  The Arnoldi process and the vector updates run
 * on the device, the small dense step (Ritz values of the Hessenberg blocks, Leja ordering,
 * divided differences, polynomial in the Hessenberg matrix) on the host inside the library.
 *   v:        work state of the same shape (the reference's NewtonWrk.v)
 *   func_id:  QP_FUNC_EXPMI  exp(-i z)  (default of the reference, TDSE)
 *             QP_FUNC_EXP    exp(z)     (`func = exp` for Liouvillians in the :LvN convention)
 *             QP_FUNC_CALLBACK  `func` is called at the Leja points only (any analytic function)
 *   K:        Krylov workspace with m_max > 2 (NewtonWrk, src/newton.jl:23-60)
 * Fails with QP_ERR_NOT_CONVERGED after max_restarts (the `@assert` of src/newton.jl:375).
 * The fine-grained entry points above remain for hosts that keep the dense step themselves. */
#define QP_FUNC_EXPMI 0
#define QP_FUNC_EXP 1
#define QP_FUNC_CALLBACK 2
typedef void (*qp_newton_func_t)(const qp_c128* z, qp_c128* f_of_z, void* user);
int32_t qp_newton_step(qp_krylov_t K, qp_state_t psi, qp_state_t v, const qp_c128* op_coeffs, double dt,
                       int32_t func_id, qp_newton_func_t func, void* user, double norm_min, double relerr,
                       int32_t max_restarts, int32_t* restarts_out);

/* NewtonWrk bookkeeping after the last qp_newton_step on this workspace (the reference sets
 * wrk.n_a, wrk.n_leja, wrk.radius and fills wrk.a, wrk.leja; src/newton.jl:381-383): counts,
 * radius and up to `capacity` coefficients / Leja points (a, leja may be NULL). */
int32_t qp_newton_last(qp_krylov_t K, int32_t* n_a, int32_t* n_leja, double* radius, qp_c128* a, qp_c128* leja,
                       int32_t capacity);

/* The host-side pieces of newton! on their own (pure host code, no device needed):
 *   diagonalize_hessenberg_matrix(Hess, m; accumulate)   src/arnoldi.jl:143-170
 *     hess column-major ld x ld; out receives m values, or m(m+1)/2 when accumulate != 0
 *     (blocks 1..m concatenated); each block sorted by (real, imag) like Julia's eigvals
 *   extend_leja!(leja, n, newpoints, n_use)              src/newton.jl:97-148  (newpoints is clobbered)
 *   extend_newton_coeffs!(a, n_a, leja, func, n_leja, radius)   src/newton.jl:176-214 */
int32_t qp_diagonalize_hessenberg(const qp_c128* hess, int32_t ld, int32_t m, int32_t accumulate,
                                  qp_c128* out, int32_t* n_out);
int32_t qp_extend_leja(qp_c128* leja, int32_t capacity, int32_t* n, qp_c128* newpoints, int32_t n_new,
                       int32_t n_use);
int32_t qp_extend_newton_coeffs(qp_c128* a, int32_t capacity, int32_t* n_a, const qp_c128* leja,
                                int32_t n_leja, int32_t func_id, qp_newton_func_t func, void* user,
                                double radius);

/* ------------------------------------------------------------------ ensembles / multi-GPU
 * An ensemble of n_total independent trajectories (same system, different control amplitudes) is
 * cut into contiguous blocks, one per rank (qp_ens_shard); every rank holds the operators and a
 * [N][B_local] batched state and steps it with qp_cheby_step / qp_cheby_propagate -- there is NO
 * communication inside the time loop.  The only collectives are the final gathers below; they run
 * over NCCL (NVLink / NVSwitch) INSIDE the library, so a host without its own communication layer
 * (Julia) can use every GPU of a box.  Replaces: the caller-side loop over one propagator per
 * trajectory that shares a spectral envelope through `control_ranges`
 * (src/cheby_propagator.jl:59-66) and collects the results.
 *
 *   single process, many GPUs:  qp_ens_create(devices, n, &ens); rank i's objects are created on
 *                               the context qp_ens_ctx(ens, i, ...) returns.  All entries of
 *                               `devices` distinct: NCCL (ncclCommInitAll).  All entries equal:
 *                               "fake ranks" on one device, the gathers use device copies (the
 *                               layout logic can be exercised on a single-GPU box).
 *   one process per GPU:        rank 0 calls qp_ens_unique_id and hands the 128 bytes to the other
 *                               processes (MPI, the launcher, a file ...); every process then
 *                               calls qp_ens_create_rank with its own context.
 * NCCL (libnccl.so.2) is loaded at first use; without it only single-rank and fake-rank ensembles
 * work (QP_ERR_UNSUPPORTED otherwise). */
int32_t qp_ens_shard(int64_t n_total, int32_t rank, int32_t n_ranks, int64_t* b0, int64_t* b1);
int32_t qp_ens_create(const int32_t* devices, int32_t n_ranks, qp_ens_t* ens);
int32_t qp_ens_unique_id(uint8_t* id /*[128]*/);
int32_t qp_ens_create_rank(qp_ctx_t ctx, int32_t rank, int32_t n_ranks, const uint8_t* id /*[128]*/, qp_ens_t* ens);
int32_t qp_ens_destroy(qp_ens_t ens);
/* transport: 0 = device copies (one rank, or fake ranks), 1 = NCCL */
int32_t qp_ens_info(qp_ens_t ens, int32_t* n_ranks, int32_t* n_local, int32_t* transport);
/* context and world rank of the local_index-th member held by this process */
int32_t qp_ens_ctx(qp_ens_t ens, int32_t local_index, qp_ctx_t* ctx, int32_t* rank);
/* Final states of the whole ensemble: local_states[i] is the [N][B_local] state of local member i.
 * full_states (or NULL): one [N][n_total] state per local member, filled on its device;
 * host_out (or NULL): [N][n_total] on the host (from the first local member).  Collective: every
 * process of the ensemble must call it. */
int32_t qp_ens_gather_states(qp_ens_t ens, const qp_state_t* local_states, int64_t n_total,
                             const qp_state_t* full_states, qp_c128* host_out);
/* Per-trajectory numbers (expectation values, norms ...): local_values[i] is member i's host array
 * [n_values][B_local]; out receives [n_values][n_total] in every process.  Collective. */
int32_t qp_ens_gather_expvals(qp_ens_t ens, const qp_c128* const* local_values, int32_t n_values,
                              int64_t n_total, qp_c128* out);

#ifdef __cplusplus
}
#endif
#endif /* QPROP_H */
