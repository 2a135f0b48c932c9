// Internal declarations of libqprop_b200 (not part of the ABI).
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdlib>
#include <utility>
#include <cstdint>
#include <cstdio>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "qprop.h"

// ---------------------------------------------------------------------------------------
// limits of the packed column word: low 28 bits = column, high 4 bits = operator index
// ---------------------------------------------------------------------------------------
constexpr int QP_COL_BITS = 28;
constexpr uint32_t QP_COL_MASK = (1u << QP_COL_BITS) - 1u;
constexpr int QP_MAX_OPS = 16;
constexpr int QP_SELL_C = 32;  // slice height of the sliced-ELL format = warp size

struct qp_timer_rec {
  int64_t ncalls = 0;
  double seconds = 0.0;
};

struct qp_ctx_s {
  // objects created on the context keep it alive: qp_ctx_destroy on a context that still has
  // objects only marks it, the resources go with the last object (hosts with finalizers -- Julia,
  // Python -- destroy things in no particular order)
  int refs = 0;
  bool destroy_requested = false;
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  std::string err;
  int64_t launches = 0;
  bool timers_on = false;
  std::map<std::string, qp_timer_rec> timers;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_stage = nullptr;  // fences reuse of the pinned coefficient staging buffer
  // reduction scratch (device partials + pinned host landing zone)
  double* d_red = nullptr;
  double* h_red = nullptr;
  size_t red_doubles = 0;
  // per-warp / per-tile partial sums of the fused expectation value and the normalization check
  double* d_part = nullptr;
  size_t part_doubles = 0;
  double* d_gemv = nullptr;  // partial dot products of the flat dense GEMV (spmv.cuh: k_gemv_flat)
  size_t gemv_doubles = 0;
  // small pinned staging area for per-step coefficient uploads
  qp_c128* h_stage = nullptr;
  size_t stage_elems = 0;
  std::set<const void*> smem_configured;  // kernels already opted in to large dynamic smem
};

void qp_ctx_release(qp_ctx_t ctx);  // drops one object reference; frees a context whose destruction was requested

// the `ctx` member of every object: a context pointer that holds a reference
struct CtxHandle {
  qp_ctx_t c = nullptr;
  CtxHandle() = default;
  CtxHandle(const CtxHandle& o) : c(o.c) { if (c) c->refs++; }
  CtxHandle& operator=(const CtxHandle& o) { return *this = o.c; }
  CtxHandle& operator=(qp_ctx_t ctx) {
    if (ctx) ctx->refs++;
    if (c) qp_ctx_release(c);
    c = ctx;
    return *this;
  }
  ~CtxHandle() { if (c) qp_ctx_release(c); }
  operator qp_ctx_t() const { return c; }
  qp_ctx_t operator->() const { return c; }
};

// one term c * P rho Q of a matrix-free left/right operator (qp_op_create_leftright)
struct LRTermHost {
  qp_op_t left = nullptr;       // borrowed sparse operator or nullptr (identity)
  uint32_t* d_rptr = nullptr;   // Q^T as CSR (row j lists (b, Q[b, j])), owned; nullptr = identity
  uint32_t* d_rcol = nullptr;
  double2* d_rval = nullptr;
  int64_t r_nnz = 0;
  double2 c = {1.0, 0.0};
};

// device form of a term inside a generator
struct LRTerm {
  const uint32_t* lptr;
  const uint32_t* lcol;
  const double2* lval;
  const uint32_t* rptr;
  const uint32_t* rcol;
  const double2* rval;
  double2 c;
  int op;  // operator index within the generator (selects the coefficient)
  int pad;
};

struct qp_op_s {
  CtxHandle ctx;
  bool dense = false;
  bool leftright = false;        // matrix-free: nrows = ncols = lr_n^2, no matrix arrays
  int64_t lr_n = 0;
  std::vector<LRTermHost> lr_terms;
  int64_t nrows = 0, ncols = 0, nnz = 0;
  // sparse: canonical CSR, Int32 indices
  uint32_t* d_ptr = nullptr;
  uint32_t* d_col = nullptr;
  double2* d_val = nullptr;
  // dense: row-major n x n
  double2* d_dense = nullptr;
};

// device view of a merged multi-operator matrix (CSR or SELL-32)
struct MatView {
  const uint32_t* ptr;   // CSR: row pointers [n+1];  SELL: slice offsets [n_slices+1] (in entries)
  const uint32_t* colop; // packed (op << 28 | col)
  const double2* val;
  int64_t n;             // rows
};

// device view of a dictionary-compressed SELL-32 matrix (QP_FORMAT_SELLD): every stored entry
// is an 8- or 16-bit code into a table of distinct (operator, value, column - row) triples;
// code 0 is the padding entry (value 0, offset 0).  A slice holds `width` chunks; chunk c of
// lane l is the 16-byte word codes[off + c*32 + l] carrying 16 (8-bit) or 8 (16-bit) codes.
struct DictView {
  const uint32_t* sptr;   // slice offsets [n_slices+1] in 16-byte words (unused when uniform_words > 0)
  const uint4* codes;
  const double2* dval;    // [n_dict]
  const int32_t* ddelta;  // [n_dict] column - row
  const uint8_t* dop;     // [n_dict] operator index
  int n_dict;
  uint32_t uniform_words; // > 0: every slice has this many words (offset = slice * uniform_words)
  int64_t n;
  // operators whose main diagonal has too many distinct values for the table keep it as an
  // explicit vector diag[i*n + row] (their diagonal entries are padding codes in the stream)
  const double2* diag;
  int n_diag;
  unsigned long long diag_ops;  // operator index of vector i in bits [4i, 4i+4)
};

// two-pass tiled format of the trajectory-batched path (tile.cu / tile_format.h), built lazily
struct qp_tile_s;
void qp_tile_free(qp_tile_s* t);
// bit-flip (XOR-stencil) form of a generator for single states (bitflip.cu), detected at qp_gen_create
struct qp_bitflip_s;
void qp_bitflip_free(qp_bitflip_s* b);

constexpr int QP_DICT_MAX = 4096;       // table entries incl. the padding entry
constexpr int QP_DICT_HASH_CAP = 16384; // open-addressing capacity used while building

struct qp_gen_s {
  CtxHandle ctx;
  int n_ops = 0, n_coeffs = 0, drift = 0;
  int64_t n = 0;
  int format = QP_FORMAT_CSR;
  std::vector<qp_op_t> ops;
  int64_t nnz_total = 0;       // true nonzeros over all operators
  int64_t stored_entries = 0;  // entries stored in the chosen format (incl. padding)
  int64_t matrix_bytes = 0;    // algorithmic M of SURVEY.md §8
  int lanes = 8;               // CSR: lanes per row (power of two <= 32)
  int sell_kernel = 1;         // SELL: 0 = LDG kernel, 1 = TMA-staged kernel (env QPROP_SELL_KERNEL)
  int tma_cfg = 2;             // TMA kernel shape (env QPROP_TMA_CFG), see launch_epi; 2 = 16 warps x 2 stages x 8 entries (best on B200, profiles/r1_variants.txt)
  // merged CSR (always built for sparse generators)
  uint32_t* d_mptr = nullptr;
  uint32_t* d_mcolop = nullptr;
  double2* d_mval = nullptr;
  // SELL-32 (built when selected)
  uint32_t* d_sptr = nullptr;
  uint32_t* d_scolop = nullptr;
  double2* d_sval = nullptr;
  int64_t n_slices = 0;
  // SELL-D (built when the dictionary of distinct entries is small enough)
  uint32_t* d_dptr = nullptr;
  uint4* d_dcodes = nullptr;
  double2* d_dval = nullptr;
  int32_t* d_ddelta = nullptr;
  uint8_t* d_dop = nullptr;
  double* d_dvalr = nullptr;  // table values as one real number each (valid when dict_realv)
  bool pair_ordered = false;  // rows ordered "columns shared by two operators first" (spmm_pairs.cuh)
  unsigned dict_has_re = 0, dict_has_im = 0;  // bit l: operator l has table values with a real / imaginary part
  // host copy of the effective per-operator coefficients (valid when they were set from the host
  // and are not per-trajectory): decides the real-table variant of the SELL-D kernel per launch
  std::vector<double2> h_coef;
  bool h_coef_valid = false;
  bool dict_realv = false;    // every operator purely real or purely imaginary
  unsigned imag_ops = 0;      // bit l: operator l is purely imaginary (value stored = Im)
  double2* d_diag = nullptr;  // explicit diagonals [n_diag][n] (see DictView)
  int n_diag = 0;
  uint8_t diag_op[QP_MAX_OPS] = {0};
  int n_dict = 0;          // 0: no dictionary
  int code_bytes = 0;      // 1 or 2
  uint32_t uniform_words = 0;
  int tail_codes = 0;      // uniform_words > 0: codes of the longest row in its LAST word (0: full / unknown)
  int64_t dict_words = 0;  // 16-byte words stored
  int64_t stored_bytes = 0;  // bytes of the matrix stream actually read per application
  // QP_FORMAT_LR: terms of all operators
  LRTerm* d_lr_terms = nullptr;
  int n_lr_terms = 0;
  int64_t lr_n = 0;
  // dense: pointers to the row-major operators
  const double2** d_dense_ops = nullptr;
  qp_bitflip_s* bitflip = nullptr;  // bit-flip (XOR-stencil) form for single states (bitflip.cu), QP_FORMAT_BITFLIP
  qp_tile_s* tile = nullptr;   // two-pass tiled format for batched states (nullptr: not tried yet)
  // device copy of the effective per-operator coefficients (drift ops = 1), [n_ops][B]
  double2* d_coef = nullptr;
  size_t coef_elems = 0;
};

struct qp_state_s {
  CtxHandle ctx;
  int64_t n = 0, batch = 1;
  double2* d = nullptr;
};

struct qp_cheby_s {
  CtxHandle ctx;
  qp_gen_t gen = nullptr;
  int64_t n = 0, batch = 1;
  double2* w1 = nullptr;
  double2* w2 = nullptr;
  std::vector<double> a;
  double Delta = 0, E_min = 0, dt = 0, limit = 1e-12;
  double* d_chk = nullptr;  // normalization-check accumulators [n_a][batch][3]
  size_t chk_doubles = 0;
};

struct qp_krylov_s {
  CtxHandle ctx;
  qp_gen_t gen = nullptr;
  int64_t n = 0, batch = 1;
  int m_max = 0;
  double2* d_pb = nullptr;  // per-trajectory numbers of a state bundle (batch > 1)
  size_t pb_elems = 0;
  double2* q = nullptr;    // (m_max + 1) vectors of length n * batch, contiguous
  double2* d_h = nullptr;  // device Hessenberg column (m_max + 2 complex): correction of a second GS round
  double2* d_hall = nullptr;      // all columns, [m_max + 1][m_max + 2]
  struct ColCtl* d_ctl = nullptr; // per-column norms / DGKS flag, [m_max + 1] (krylov.cu)
  // diagnostics of the last qp_newton_step (NewtonWrk.n_a / n_leja / radius / a / leja)
  int last_n_a = 0, last_n_leja = 0;
  double last_radius = 0.0;
  std::vector<qp_c128> last_a, last_leja;
  // lazy normalisation (single states): q[i] is stored UNNORMALISED, its factor is h_scale[i] on the host and
  // d_ctl[i - 1].inv on the device (q[0]: 1); krylov.cu
  std::vector<double> h_scale;
};

// ---------------------------------------------------------------------------------------
// error handling
// ---------------------------------------------------------------------------------------
int32_t qp_fail(qp_ctx_t ctx, int32_t code, const char* fmt, ...);

#define QP_CUDA(ctx, call)                                                                  \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      cudaGetLastError();                                                                   \
      return qp_fail((ctx), e__ == cudaErrorMemoryAllocation ? QP_ERR_OOM : QP_ERR_CUDA,    \
                     "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__,     \
                     __LINE__);                                                             \
    }                                                                                       \
  } while (0)

#define QP_CHECK(expr)               \
  do {                               \
    int32_t s__ = (expr);            \
    if (s__ != QP_OK) return s__;    \
  } while (0)

#define QP_REQUIRE(ctx, cond, ...)                                   \
  do {                                                               \
    if (!(cond)) return qp_fail((ctx), QP_ERR_INVALID_ARG, __VA_ARGS__); \
  } while (0)

// count + check a kernel launch
#define QP_LAUNCHED(ctx)                    \
  do {                                      \
    (ctx)->launches++;                      \
    QP_CUDA((ctx), cudaGetLastError());     \
  } while (0)

int32_t qp_ctx_bind(qp_ctx_t ctx);  // cudaSetDevice
int32_t qp_ctx_reserve_red(qp_ctx_t ctx, size_t doubles);
int32_t qp_ctx_reserve_gemv(qp_ctx_t ctx, size_t doubles);
int32_t qp_ctx_reserve_stage(qp_ctx_t ctx, size_t elems);
int32_t qp_ctx_reserve_part(qp_ctx_t ctx, size_t doubles);
// chk[b][k] = sum over the n_slots slots of part[slot][b][k] in a fixed order (deterministic)
int32_t qp_part_reduce(qp_ctx_t ctx, const double* part, int64_t n_slots, int64_t batch, double* chk);

// timers ("matrix-vector product", "prop_step!", ...): CUDA events on the ctx stream
struct QpScopedTimer {
  qp_ctx_t ctx;
  const char* label;
  bool active;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  QpScopedTimer(qp_ctx_t c, const char* l);
  ~QpScopedTimer();
};

// ---------------------------------------------------------------------------------------
// programmatic dependent launch: the launch of a kernel overlaps the execution of the previous
// kernel of the stream; kernels launched this way call pdl_sync() before touching memory
// (QPROP_PDL=0: plain stream order)
// ---------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

template <typename... KArgs, typename... Args>
inline cudaError_t qp_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                 Args&&... args) {
  static const int pdl = getenv("QPROP_PDL") ? atoi(getenv("QPROP_PDL")) : 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...);
}
#endif

// ---------------------------------------------------------------------------------------
// internal kernels shared between translation units
// ---------------------------------------------------------------------------------------

// upload effective per-operator coefficients: drift operators get 1, the rest op_coeffs.
// per_traj == 0: n_coeffs numbers broadcast; else [n_coeffs][batch].
int32_t qp_gen_set_coeffs(qp_gen_t gen, const qp_c128* op_coeffs, int per_traj, int64_t batch,
                          int* coef_stride_out);

// y <- beta*y + alpha*H x using the coefficients currently in gen->d_coef
// y <- beta y + alpha * (*alpha_dev) * H x : alpha_dev (or nullptr) is a real factor in device memory
int32_t qp_gen_apply_scaled(qp_gen_t gen, int coef_stride, double2 alpha, const double* alpha_dev, double2 beta, const double2* x,
                            double2* y, int64_t batch);
int32_t qp_gen_apply(qp_gen_t gen, int coef_stride, double2 alpha, double2 beta, const double2* x,
                     double2* y, int64_t batch);

// deterministic reductions; results land in ctx->h_red after a stream sync
int32_t qp_reduce_dot(qp_ctx_t ctx, const double2* x, const double2* y, int64_t n, int64_t batch,
                      qp_c128* out);
int32_t qp_reduce_norm2(qp_ctx_t ctx, const double2* x, int64_t n, int64_t batch, double* out);
