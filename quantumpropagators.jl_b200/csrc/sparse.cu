// Operators and generators: upload (CSC/CSR Int64 -> device CSR Int32), device-side merge of
// the component operators into one tagged matrix, SELL-32 construction, format selection,
// and the host-side dispatcher of the fused SpMV kernels.
#include <algorithm>
#include <cmath>
#include <cstring>

#include <cub/device/device_scan.cuh>

#include "spmv.cuh"
#include "spmv_lr.cuh"
#include "spmm_pairs.cuh"
#include "dense_mma.cuh"

// ---------------------------------------------------------------------------------------
// operator upload
// ---------------------------------------------------------------------------------------

extern "C" int32_t qp_op_upload_sparse(qp_ctx_t ctx, int64_t nrows, int64_t ncols, int64_t nnz,
                                       const int64_t* ptr, const int64_t* idx, const qp_c128* val,
                                       int32_t layout, int32_t index_base, qp_op_t* out) {
  QP_CHECK(qp_ctx_bind(ctx));
  QP_REQUIRE(ctx, out != nullptr, "qp_op_upload_sparse: null output pointer");
  *out = nullptr;
  QP_REQUIRE(ctx, nrows >= 1 && ncols >= 1 && nnz >= 0, "qp_op_upload_sparse: bad shape");
  QP_REQUIRE(ctx, ptr && (nnz == 0 || (idx && val)), "qp_op_upload_sparse: null array");
  QP_REQUIRE(ctx, layout == QP_LAYOUT_CSC || layout == QP_LAYOUT_CSR, "qp_op_upload_sparse: bad layout %d", layout);
  QP_REQUIRE(ctx, index_base == 0 || index_base == 1, "qp_op_upload_sparse: index_base must be 0 or 1");
  if (nrows >= (int64_t(1) << QP_COL_BITS) || ncols >= (int64_t(1) << QP_COL_BITS))
    return qp_fail(ctx, QP_ERR_UNSUPPORTED, "qp_op_upload_sparse: dimension %lld x %lld exceeds 2^%d",
                   (long long)nrows, (long long)ncols, QP_COL_BITS);
  if (nnz >= (int64_t(1) << 32))
    return qp_fail(ctx, QP_ERR_UNSUPPORTED, "qp_op_upload_sparse: nnz=%lld exceeds 2^32", (long long)nnz);

  const int64_t n_major = layout == QP_LAYOUT_CSR ? nrows : ncols;
  const int64_t n_minor = layout == QP_LAYOUT_CSR ? ncols : nrows;
  QP_REQUIRE(ctx, ptr[0] == index_base && ptr[n_major] == nnz + index_base,
             "qp_op_upload_sparse: pointer array does not span [base, nnz+base]");
  for (int64_t i = 0; i < n_major; ++i)
    QP_REQUIRE(ctx, ptr[i + 1] >= ptr[i], "qp_op_upload_sparse: pointer array not monotone at %lld", (long long)i);
  for (int64_t k = 0; k < nnz; ++k) {
    int64_t j = idx[k] - index_base;
    QP_REQUIRE(ctx, j >= 0 && j < n_minor, "qp_op_upload_sparse: index %lld out of range at entry %lld",
               (long long)idx[k], (long long)k);
  }

  std::vector<uint32_t> h_ptr((size_t)nrows + 1), h_col((size_t)nnz);
  std::vector<double2> h_val((size_t)nnz);
  if (layout == QP_LAYOUT_CSR) {
    for (int64_t i = 0; i <= nrows; ++i) h_ptr[i] = (uint32_t)(ptr[i] - index_base);
    for (int64_t k = 0; k < nnz; ++k) {
      h_col[k] = (uint32_t)(idx[k] - index_base);
      h_val[k] = make_double2(val[k].re, val[k].im);
    }
  } else {
    // CSC -> CSR by counting sort over rows (the arrays of a CSC matrix are the CSR arrays of
    // its transpose, so a real transposition is required -- not a reinterpretation)
    std::vector<uint32_t> count((size_t)nrows + 1, 0);
    for (int64_t k = 0; k < nnz; ++k) count[(size_t)(idx[k] - index_base) + 1]++;
    h_ptr[0] = 0;
    for (int64_t i = 0; i < nrows; ++i) h_ptr[i + 1] = h_ptr[i] + count[i + 1];
    std::vector<uint32_t> fill(h_ptr.begin(), h_ptr.end() - 1);
    for (int64_t c = 0; c < ncols; ++c) {
      for (int64_t k = ptr[c] - index_base; k < ptr[c + 1] - index_base; ++k) {
        int64_t r = idx[k] - index_base;
        uint32_t dst = fill[r]++;
        h_col[dst] = (uint32_t)c;
        h_val[dst] = make_double2(val[k].re, val[k].im);
      }
    }
  }

  qp_op_t op = new qp_op_s();
  op->ctx = ctx;
  op->nrows = nrows;
  op->ncols = ncols;
  op->nnz = nnz;
  auto cleanup = [&](int32_t rc) {
    cudaFree(op->d_ptr);
    cudaFree(op->d_col);
    cudaFree(op->d_val);
    delete op;
    return rc;
  };
  cudaError_t e;
  if ((e = cudaMalloc(&op->d_ptr, sizeof(uint32_t) * (nrows + 1))) != cudaSuccess ||
      (e = cudaMalloc(&op->d_col, sizeof(uint32_t) * std::max<int64_t>(nnz, 1))) != cudaSuccess ||
      (e = cudaMalloc(&op->d_val, sizeof(double2) * std::max<int64_t>(nnz, 1))) != cudaSuccess) {
    cudaGetLastError();
    return cleanup(qp_fail(ctx, QP_ERR_OOM, "qp_op_upload_sparse: cudaMalloc failed: %s", cudaGetErrorString(e)));
  }
  if ((e = cudaMemcpyAsync(op->d_ptr, h_ptr.data(), sizeof(uint32_t) * (nrows + 1), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess ||
      (nnz > 0 && (e = cudaMemcpyAsync(op->d_col, h_col.data(), sizeof(uint32_t) * nnz, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) ||
      (nnz > 0 && (e = cudaMemcpyAsync(op->d_val, h_val.data(), sizeof(double2) * nnz, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) ||
      (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) {
    cudaGetLastError();
    return cleanup(qp_fail(ctx, QP_ERR_CUDA, "qp_op_upload_sparse: upload failed: %s", cudaGetErrorString(e)));
  }
  *out = op;
  return QP_OK;
}

__global__ void k_transpose_colmajor(const double2* __restrict__ in, double2* __restrict__ out, int64_t n) {
  __shared__ double2 tile[32][33];
  int64_t bx = blockIdx.x * 32, by = blockIdx.y * 32;
  // in is column-major: element (r, c) at in[c*n + r]; out row-major: out[r*n + c]
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int64_t c = by + j, r = bx + threadIdx.x;
    if (r < n && c < n) tile[j][threadIdx.x] = in[c * n + r];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int64_t r = bx + j, c = by + threadIdx.x;
    if (r < n && c < n) out[r * n + c] = tile[threadIdx.x][j];
  }
}

extern "C" int32_t qp_op_upload_dense(qp_ctx_t ctx, int64_t n, const qp_c128* colmajor, qp_op_t* out) {
  QP_CHECK(qp_ctx_bind(ctx));
  QP_REQUIRE(ctx, out != nullptr && colmajor != nullptr, "qp_op_upload_dense: null argument");
  *out = nullptr;
  QP_REQUIRE(ctx, n >= 1, "qp_op_upload_dense: n must be >= 1");
  if (n >= (int64_t(1) << QP_COL_BITS)) return qp_fail(ctx, QP_ERR_UNSUPPORTED, "qp_op_upload_dense: n too large");
  qp_op_t op = new qp_op_s();
  op->ctx = ctx;
  op->dense = true;
  op->nrows = op->ncols = n;
  op->nnz = n * n;
  double2* d_tmp = nullptr;
  cudaError_t e;
  if ((e = cudaMalloc(&op->d_dense, sizeof(double2) * n * n)) != cudaSuccess ||
      (e = cudaMalloc(&d_tmp, sizeof(double2) * n * n)) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(op->d_dense);
    delete op;
    return qp_fail(ctx, QP_ERR_OOM, "qp_op_upload_dense: cudaMalloc failed: %s", cudaGetErrorString(e));
  }
  e = cudaMemcpyAsync(d_tmp, colmajor, sizeof(double2) * n * n, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) {
    dim3 grid((unsigned)((n + 31) / 32), (unsigned)((n + 31) / 32)), block(32, 8);
    k_transpose_colmajor<<<grid, block, 0, ctx->stream>>>(d_tmp, op->d_dense, n);
    ctx->launches++;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_tmp);
  if (e != cudaSuccess) {
    cudaGetLastError();
    cudaFree(op->d_dense);
    delete op;
    return qp_fail(ctx, QP_ERR_CUDA, "qp_op_upload_dense: upload failed: %s", cudaGetErrorString(e));
  }
  *out = op;
  return QP_OK;
}

extern "C" int32_t qp_op_destroy(qp_op_t op) {
  if (!op) return QP_OK;
  cudaSetDevice(op->ctx->device);
  cudaStreamSynchronize(op->ctx->stream);
  cudaFree(op->d_ptr);
  cudaFree(op->d_col);
  cudaFree(op->d_val);
  cudaFree(op->d_dense);
  for (LRTermHost& t : op->lr_terms) {
    cudaFree(t.d_rptr);
    cudaFree(t.d_rcol);
    cudaFree(t.d_rval);
  }
  delete op;
  return QP_OK;
}

// ---------------------------------------------------------------------------------------
// matrix-free left/right super-operators (spmv_lr.cuh)
// ---------------------------------------------------------------------------------------
extern "C" int32_t qp_op_create_leftright(qp_ctx_t ctx, int64_t n, int32_t n_terms, const qp_op_t* left,
                                          const qp_op_t* right, const qp_c128* coeffs, qp_op_t* out) {
  QP_CHECK(qp_ctx_bind(ctx));
  QP_REQUIRE(ctx, out != nullptr, "qp_op_create_leftright: null output pointer");
  *out = nullptr;
  QP_REQUIRE(ctx, n >= 1 && n_terms >= 1 && left && right && coeffs, "qp_op_create_leftright: bad argument");
  if (n_terms > QP_LR_MAX_TERMS)
    return qp_fail(ctx, QP_ERR_UNSUPPORTED, "qp_op_create_leftright: %d terms exceed the limit of %d", n_terms, QP_LR_MAX_TERMS);
  if (n * n >= (int64_t(1) << QP_COL_BITS))
    return qp_fail(ctx, QP_ERR_UNSUPPORTED, "qp_op_create_leftright: dimension n^2 = %lld exceeds 2^%d", (long long)(n * n), QP_COL_BITS);
  for (int t = 0; t < n_terms; ++t)
    for (qp_op_t f : {left[t], right[t]}) {
      if (!f) continue;
      QP_REQUIRE(ctx, f->ctx == ctx, "qp_op_create_leftright: factor of term %d belongs to another context", t);
      QP_REQUIRE(ctx, !f->dense && !f->leftright, "qp_op_create_leftright: factors must be sparse operators (term %d)", t);
      QP_REQUIRE(ctx, f->nrows == n && f->ncols == n, "qp_op_create_leftright: factor of term %d is %lld x %lld, expected %lld x %lld",
                 t, (long long)f->nrows, (long long)f->ncols, (long long)n, (long long)n);
    }
  qp_op_t op = new qp_op_s();
  op->ctx = ctx;
  op->leftright = true;
  op->lr_n = n;
  op->nrows = op->ncols = n * n;
  auto fail = [&](int32_t rc) {
    qp_op_destroy(op);
    return rc;
  };
  for (int t = 0; t < n_terms; ++t) {
    LRTermHost T;
    T.left = left[t];
    T.c = make_double2(coeffs[t].re, coeffs[t].im);
    op->nnz += left[t] ? left[t]->nnz : 0;
    if (qp_op_t Q = right[t]) {
      // Q^T in CSR: row j lists the entries (b, Q[b, j]) of column j of Q (counting sort on the host:
      // the factors are n x n, tiny next to the n^2-dimensional state)
      const int64_t nnz = Q->nnz;
      std::vector<uint32_t> h_ptr((size_t)n + 1), h_col((size_t)std::max<int64_t>(nnz, 1));
      std::vector<double2> h_val((size_t)std::max<int64_t>(nnz, 1));
      cudaError_t e = cudaMemcpy(h_ptr.data(), Q->d_ptr, sizeof(uint32_t) * (n + 1), cudaMemcpyDeviceToHost);
      if (e == cudaSuccess && nnz > 0) e = cudaMemcpy(h_col.data(), Q->d_col, sizeof(uint32_t) * nnz, cudaMemcpyDeviceToHost);
      if (e == cudaSuccess && nnz > 0) e = cudaMemcpy(h_val.data(), Q->d_val, sizeof(double2) * nnz, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) return fail(qp_fail(ctx, QP_ERR_CUDA, "qp_op_create_leftright: download failed: %s", cudaGetErrorString(e)));
      std::vector<uint32_t> t_ptr((size_t)n + 1, 0), t_col((size_t)std::max<int64_t>(nnz, 1));
      std::vector<double2> t_val((size_t)std::max<int64_t>(nnz, 1));
      for (int64_t k = 0; k < nnz; ++k) t_ptr[h_col[k] + 1]++;
      for (int64_t c = 0; c < n; ++c) t_ptr[c + 1] += t_ptr[c];
      std::vector<uint32_t> fill(t_ptr.begin(), t_ptr.end() - 1);
      for (int64_t r = 0; r < n; ++r)
        for (uint32_t k = h_ptr[r]; k < h_ptr[r + 1]; ++k) {
          const uint32_t dst = fill[h_col[k]]++;
          t_col[dst] = (uint32_t)r;
          t_val[dst] = h_val[k];
        }
      if ((e = cudaMalloc(&T.d_rptr, sizeof(uint32_t) * (n + 1))) != cudaSuccess ||
          (e = cudaMalloc(&T.d_rcol, sizeof(uint32_t) * std::max<int64_t>(nnz, 1))) != cudaSuccess ||
          (e = cudaMalloc(&T.d_rval, sizeof(double2) * std::max<int64_t>(nnz, 1))) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(T.d_rptr);
        cudaFree(T.d_rcol);
        cudaFree(T.d_rval);
        return fail(qp_fail(ctx, QP_ERR_OOM, "qp_op_create_leftright: cudaMalloc failed: %s", cudaGetErrorString(e)));
      }
      op->lr_terms.push_back(T);  // owned from here on (freed by qp_op_destroy)
      if ((e = cudaMemcpy(T.d_rptr, t_ptr.data(), sizeof(uint32_t) * (n + 1), cudaMemcpyHostToDevice)) != cudaSuccess ||
          (nnz > 0 && (e = cudaMemcpy(T.d_rcol, t_col.data(), sizeof(uint32_t) * nnz, cudaMemcpyHostToDevice)) != cudaSuccess) ||
          (nnz > 0 && (e = cudaMemcpy(T.d_rval, t_val.data(), sizeof(double2) * nnz, cudaMemcpyHostToDevice)) != cudaSuccess))
        return fail(qp_fail(ctx, QP_ERR_CUDA, "qp_op_create_leftright: upload failed: %s", cudaGetErrorString(e)));
      op->lr_terms.back().r_nnz = nnz;
      op->nnz += nnz;
    } else {
      op->lr_terms.push_back(T);
    }
  }
  *out = op;
  return QP_OK;
}

extern "C" int32_t qp_op_info(qp_op_t op, int64_t* nrows, int64_t* ncols, int64_t* nnz, int32_t* is_dense) {
  if (!op) return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_op_info: null operator");
  if (nrows) *nrows = op->nrows;
  if (ncols) *ncols = op->ncols;
  if (nnz) *nnz = op->nnz;
  if (is_dense) *is_dense = op->dense ? 1 : 0;
  return QP_OK;
}

// ---------------------------------------------------------------------------------------
// merge kernels
// ---------------------------------------------------------------------------------------

struct OpPtrs {
  const uint32_t* ptr[QP_MAX_OPS];
  const uint32_t* col[QP_MAX_OPS];
  const double2* val[QP_MAX_OPS];
};

__global__ void k_merged_len(OpPtrs ops, int n_ops, int64_t n, uint32_t* __restrict__ len) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n) return;
  uint32_t s = 0;
  for (int l = 0; l < n_ops; ++l) s += ops.ptr[l][r + 1] - ops.ptr[l][r];
  len[r] = s;
}

// one warp per row: copy the row's entries operator by operator, tagging the column word
__global__ void k_merge_rows(OpPtrs ops, int n_ops, int64_t n, const uint32_t* __restrict__ mptr,
                             uint32_t* __restrict__ mcolop, double2* __restrict__ mval) {
  int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (r >= n) return;
  uint32_t dst = mptr[r];
  for (int l = 0; l < n_ops; ++l) {
    uint32_t p0 = ops.ptr[l][r], p1 = ops.ptr[l][r + 1];
    for (uint32_t k = p0 + lane; k < p1; k += 32) {
      mcolop[dst + (k - p0)] = ops.col[l][k] | ((uint32_t)l << QP_COL_BITS);
      mval[dst + (k - p0)] = ops.val[l][k];
    }
    dst += p1 - p0;
  }
}

// Orders every merged row "columns hit by >= 2 operators first (column by column, operator by
// operator), then the rest by column": the trajectory-batched pair kernel (spmm_pairs.cuh) then
// finds the entries that share a gather next to each other and aligned to even positions.
// Thread per row, rows of up to 64 entries (longer rows are left in operator order: every kernel
// accepts any order).  *n_pairs counts the columns found shared.
__global__ void k_pair_order_rows(const uint32_t* __restrict__ mptr, uint32_t* __restrict__ mcolop,
                                  double2* __restrict__ mval, int64_t n, unsigned long long* __restrict__ n_pairs) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n) return;
  const uint32_t p0 = mptr[r];
  const int len = (int)(mptr[r + 1] - p0);
  if (len < 2 || len > 64) return;
  uint32_t co[64];
  double2 v[64];
  unsigned long long key[64];
  for (int a = 0; a < len; ++a) {
    co[a] = mcolop[p0 + a];
    v[a] = mval[p0 + a];
  }
  int shared_cols = 0;
  for (int a = 0; a < len; ++a) {
    const uint32_t col = co[a] & QP_COL_MASK;
    int same = 0, first = 1;
    for (int b = 0; b < len; ++b)
      if ((co[b] & QP_COL_MASK) == col) {
        ++same;
        if (b < a) first = 0;
      }
    // exactly two operators on the column: a pair (three or more stay with the singles so that
    // pairs remain aligned to even positions)
    const unsigned long long single = same == 2 ? 0ull : 1ull;
    key[a] = (single << 40) | ((unsigned long long)col << 4) | (unsigned long long)(co[a] >> QP_COL_BITS);
    if (same == 2 && first) ++shared_cols;
  }
  for (int a = 1; a < len; ++a) {  // insertion sort by key
    const unsigned long long k = key[a];
    const uint32_t c = co[a];
    const double2 val = v[a];
    int b = a - 1;
    while (b >= 0 && key[b] > k) {
      key[b + 1] = key[b];
      co[b + 1] = co[b];
      v[b + 1] = v[b];
      --b;
    }
    key[b + 1] = k;
    co[b + 1] = c;
    v[b + 1] = val;
  }
  for (int a = 0; a < len; ++a) {
    mcolop[p0 + a] = co[a];
    mval[p0 + a] = v[a];
  }
  if (shared_cols) atomicAdd(n_pairs, (unsigned long long)shared_cols);
}

// SELL-32: slice width = longest merged row in the slice
__global__ void k_sell_widths(const uint32_t* __restrict__ mptr, int64_t n, int64_t n_slices,
                              uint32_t* __restrict__ slice_entries) {
  int64_t s = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (s >= n_slices) return;
  int64_t r = s * QP_SELL_C + lane;
  uint32_t len = r < n ? mptr[r + 1] - mptr[r] : 0u;
  for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
  if (lane == 0) slice_entries[s] = len * QP_SELL_C;
}

__global__ void k_sell_fill(const uint32_t* __restrict__ mptr, const uint32_t* __restrict__ mcolop,
                            const double2* __restrict__ mval, int64_t n, int64_t n_slices,
                            const uint32_t* __restrict__ sptr, uint32_t* __restrict__ scolop,
                            double2* __restrict__ sval) {
  int64_t s = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (s >= n_slices) return;
  int64_t r = s * QP_SELL_C + lane;
  uint32_t base = sptr[s];
  uint32_t width = (sptr[s + 1] - base) / QP_SELL_C;
  uint32_t p0 = 0, len = 0;
  if (r < n) {
    p0 = mptr[r];
    len = mptr[r + 1] - p0;
  }
  // padding: value 0 on the row's own column (any valid column works), operator 0
  uint32_t pad_col = (uint32_t)(r < n ? r : n - 1);
  for (uint32_t j = 0; j < width; ++j) {
    uint32_t dst = base + j * QP_SELL_C + lane;
    if (j < len) {
      scolop[dst] = mcolop[p0 + j];
      sval[dst] = mval[p0 + j];
    } else {
      scolop[dst] = pad_col;
      sval[dst] = make_double2(0.0, 0.0);
    }
  }
}

// ---------------------------------------------------------------------------------------
// dictionary compression (QP_FORMAT_SELLD)
//
// Hamiltonians assembled from tensor products of few-level operators have very few distinct
// (operator, value, column - row) triples (TFIM n = 20: 81; transmon chain: a few hundred),
// so an entry can be stored as one small code instead of 16 B value + 4 B column.  The table
// is built on the device with an exact-match open-addressing hash (no lossy fingerprints),
// compacted and ordered on the host (deterministic code numbering), and the codes are laid
// out slice by slice in 16-byte words so that a warp-wide load is one 512 B segment.
// ---------------------------------------------------------------------------------------

constexpr uint16_t QP_DICT_DIAG_SLOT = 0xfffeu;  // slot_of marker: entry lives in an explicit diagonal

struct DictSlot {
  unsigned long long a, b, c;  // (op << 32 | uint32(col - row)), bits of re, bits of im
  unsigned int tag;            // 0 empty, 1 being written, 2 ready
  unsigned int pad;
};

__device__ __forceinline__ unsigned long long dict_mix(unsigned long long a, unsigned long long b,
                                                       unsigned long long c) {
  unsigned long long h = a * 0x9E3779B97F4A7C15ull;
  h ^= (b + 0xBF58476D1CE4E5B9ull + (h << 6) + (h >> 2));
  h *= 0x94D049BB133111EBull;
  h ^= (c + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
  h ^= h >> 29;
  h *= 0xBF58476D1CE4E5B9ull;
  h ^= h >> 32;
  return h;
}

// returns the slot of the key, inserting it if absent; 0xffffffff after overflow
__device__ uint32_t dict_find_or_insert(DictSlot* tab, unsigned long long a, unsigned long long b,
                                        unsigned long long c, unsigned int* ctl /*[0]=count,[1]=overflow*/,
                                        unsigned int max_count) {
  const uint32_t mask = QP_DICT_HASH_CAP - 1;
  uint32_t s = (uint32_t)dict_mix(a, b, c) & mask;
  for (uint32_t probe = 0; probe < (uint32_t)QP_DICT_HASH_CAP; ++probe) {
    if (*(volatile unsigned int*)(ctl + 1)) return 0xffffffffu;
    volatile unsigned int* tag = &tab[s].tag;
    unsigned int t = *tag;
    if (t == 0u) {
      t = atomicCAS(&tab[s].tag, 0u, 1u);
      if (t == 0u) {  // slot claimed
        if (atomicAdd(ctl, 1u) >= max_count) atomicExch(ctl + 1, 1u);
        tab[s].a = a;
        tab[s].b = b;
        tab[s].c = c;
        __threadfence();
        atomicExch(&tab[s].tag, 2u);
        return s;
      }
    }
    while (t == 1u) {  // another thread is publishing this slot
      if (*(volatile unsigned int*)(ctl + 1)) return 0xffffffffu;
      t = *tag;
    }
    __threadfence();
    const volatile DictSlot* q = tab + s;
    if (q->a == a && q->b == b && q->c == c) return s;
    s = (s + 1) & mask;
  }
  atomicExch(ctl + 1, 1u);
  return 0xffffffffu;
}

// warp per row of the merged CSR matrix; slot_of[k] = hash slot of entry k
__global__ void k_dict_scan(const uint32_t* __restrict__ mptr, const uint32_t* __restrict__ mcolop,
                            const double2* __restrict__ mval, int64_t n, DictSlot* tab, unsigned int* ctl,
                            unsigned int max_count, uint16_t* __restrict__ slot_of, int skip_diag) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; r < n; r += warps) {
    if (*(volatile unsigned int*)(ctl + 1)) return;
    const uint32_t p0 = mptr[r], p1 = mptr[r + 1];
    for (uint32_t k0 = p0; k0 < p1; k0 += 32) {
      const uint32_t k = k0 + lane;
      bool have = k < p1;
      unsigned long long a = 0, b = 0, c = 0;
      if (have) {
        const uint32_t co = mcolop[k];
        const double2 v = mval[k];
        const int32_t delta = (int32_t)((int64_t)(co & QP_COL_MASK) - r);
        if (skip_diag && delta == 0) {  // kept as an explicit diagonal vector instead
          slot_of[k] = QP_DICT_DIAG_SLOT;
          have = false;
        }
        a = ((unsigned long long)(co >> QP_COL_BITS) << 32) | (unsigned long long)(uint32_t)delta;
        b = (unsigned long long)__double_as_longlong(v.x);
        c = (unsigned long long)__double_as_longlong(v.y);
      }
      // one insertion per distinct key of the warp (cuts contention on the few hot slots)
      const unsigned act = __ballot_sync(0xffffffffu, have);
      if (have) {
        const unsigned long long h = dict_mix(a, b, c);
        const unsigned peers = __match_any_sync(act, h);
        const int leader = __ffs(peers) - 1;
        uint32_t slot = 0;
        if (lane == leader) slot = dict_find_or_insert(tab, a, b, c, ctl, max_count);
        slot = __shfl_sync(peers, slot, leader);
        // same 64-bit mix but a different key (vanishingly rare): look it up individually
        const unsigned long long la = __shfl_sync(peers, a, leader), lb = __shfl_sync(peers, b, leader),
                                 lc = __shfl_sync(peers, c, leader);
        if (la != a || lb != b || lc != c) slot = dict_find_or_insert(tab, a, b, c, ctl, max_count);
        slot_of[k] = (uint16_t)(slot & 0xffffu);
      }
    }
  }
}

// which operators have entries on the main diagonal (flags[op] = 1)
__global__ void k_diag_flags(const uint32_t* __restrict__ mptr, const uint32_t* __restrict__ mcolop, int64_t n,
                             int* __restrict__ flags) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x)
    for (uint32_t k = mptr[r]; k < mptr[r + 1]; ++k) {
      const uint32_t co = mcolop[k];
      if ((int64_t)(co & QP_COL_MASK) == r && flags[co >> QP_COL_BITS] == 0) flags[co >> QP_COL_BITS] = 1;
    }
}

// diag[index_of[op] * n + r] = sum of operator op's entries at (r, r)
__global__ void k_diag_extract(const uint32_t* __restrict__ mptr, const uint32_t* __restrict__ mcolop,
                               const double2* __restrict__ mval, int64_t n, const int* __restrict__ index_of,
                               double2* __restrict__ diag) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x)
    for (uint32_t k = mptr[r]; k < mptr[r + 1]; ++k) {
      const uint32_t co = mcolop[k];
      if ((int64_t)(co & QP_COL_MASK) != r) continue;
      double2* d = diag + (int64_t)index_of[co >> QP_COL_BITS] * n + r;
      const double2 v = mval[k];
      d->x += v.x;
      d->y += v.y;
    }
}

__global__ void k_max_row_len(const uint32_t* __restrict__ mptr, int64_t n, unsigned int* __restrict__ out) {
  unsigned int m = 0;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x)
    m = max(m, mptr[r + 1] - mptr[r]);
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

// 16-byte words per slice: ceil(longest row / codes per word) * 32
__global__ void k_selld_widths(const uint32_t* __restrict__ mptr, int64_t n, int64_t n_slices, int cpw,
                               uint32_t* __restrict__ slice_words) {
  int64_t s = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (s >= n_slices) return;
  int64_t r = s * QP_SELL_C + lane;
  uint32_t len = r < n ? mptr[r + 1] - mptr[r] : 0u;
  for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
  if (lane == 0) slice_words[s] = ((len + cpw - 1) / cpw) * QP_SELL_C;
}

template <int CB>
__global__ void k_selld_fill(const uint32_t* __restrict__ mptr, const uint16_t* __restrict__ slot_of,
                             const uint16_t* __restrict__ remap, int64_t n, int64_t n_slices,
                             const uint32_t* __restrict__ dptr, uint4* __restrict__ codes) {
  constexpr int CPW = 16 / CB;
  int64_t s = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (s >= n_slices) return;
  int64_t r = s * QP_SELL_C + lane;
  const uint32_t base = dptr[s];
  const uint32_t chunks = (dptr[s + 1] - base) / QP_SELL_C;
  uint32_t p0 = 0, len = 0;
  if (r < n) {
    p0 = mptr[r];
    len = mptr[r + 1] - p0;
  }
  for (uint32_t ch = 0; ch < chunks; ++ch) {
    uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int t = 0; t < CPW; ++t) {
      const uint32_t j = ch * CPW + t;
      uint32_t code = 0u;  // 0 = padding (also stands in for entries kept in the explicit diagonals)
      if (j < len) {
        const uint16_t slot = slot_of[p0 + j];
        if (slot != QP_DICT_DIAG_SLOT) code = remap[slot];
      }
      if (CB == 1) w[t >> 2] |= code << (8 * (t & 3));
      else w[t >> 1] |= code << (16 * (t & 1));
    }
    codes[base + ch * QP_SELL_C + lane] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

static int32_t exclusive_scan_u32(qp_ctx_t ctx, const uint32_t* d_in, uint32_t* d_out, int64_t count) {
  void* d_temp = nullptr;
  size_t temp_bytes = 0;
  QP_CUDA(ctx, cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, d_in, d_out, (int)count, ctx->stream));
  QP_CUDA(ctx, cudaMalloc(&d_temp, temp_bytes ? temp_bytes : 1));
  cudaError_t e = cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, d_in, d_out, (int)count, ctx->stream);
  ctx->launches += 2;
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_temp);
  QP_CUDA(ctx, e);
  return QP_OK;
}

static int pow2_floor(int64_t v) {
  int p = 1;
  while ((int64_t)p * 2 <= v) p *= 2;
  return p;
}

static void gen_free(qp_gen_t g) {
  cudaFree(g->d_mptr);
  cudaFree(g->d_mcolop);
  cudaFree(g->d_mval);
  cudaFree(g->d_sptr);
  cudaFree(g->d_scolop);
  cudaFree(g->d_sval);
  cudaFree(g->d_dptr);
  cudaFree(g->d_dcodes);
  cudaFree(g->d_dval);
  cudaFree(g->d_ddelta);
  cudaFree(g->d_dop);
  cudaFree(g->d_diag);
  cudaFree(g->d_dvalr);
  cudaFree((void*)g->d_dense_ops);
  cudaFree(g->d_lr_terms);
  cudaFree(g->d_coef);
  qp_tile_free(g->tile);
  qp_bitflip_free(g->bitflip);
  delete g;
}

// Builds the SELL-D storage of a generator from its merged CSR arrays.  *ok = false (and
// nothing allocated) when the matrix has more than QP_DICT_MAX-1 distinct entries.
static int32_t build_dict(qp_ctx_t ctx, qp_gen_t g, bool* ok, bool skip_diag) {
  *ok = false;
  const int64_t n = g->n, nnz = g->nnz_total, n_slices = g->n_slices;
  if (nnz == 0) return QP_OK;
  DictSlot* d_tab = nullptr;
  unsigned int* d_ctl = nullptr;
  uint16_t* d_slot_of = nullptr;
  uint16_t* d_remap = nullptr;
  uint32_t* d_words = nullptr;
  auto release = [&]() {
    cudaFree(d_tab);
    cudaFree(d_ctl);
    cudaFree(d_slot_of);
    cudaFree(d_remap);
    cudaFree(d_words);
  };
  auto drop = [&]() {
    cudaFree(g->d_dptr);
    cudaFree(g->d_dcodes);
    cudaFree(g->d_dval);
    cudaFree(g->d_ddelta);
    cudaFree(g->d_dop);
    cudaFree(g->d_diag);
    cudaFree(g->d_dvalr);
    g->d_dvalr = nullptr;
    g->d_dptr = nullptr;
    g->d_dcodes = nullptr;
    g->d_dval = nullptr;
    g->d_ddelta = nullptr;
    g->d_dop = nullptr;
    g->d_diag = nullptr;
    g->n_diag = 0;
    g->n_dict = 0;
  };
#define D_CUDA(call)                                                                              \
  do {                                                                                            \
    cudaError_t e__ = (call);                                                                     \
    if (e__ != cudaSuccess) {                                                                     \
      cudaGetLastError();                                                                         \
      release();                                                                                  \
      drop();                                                                                     \
      return qp_fail(ctx, e__ == cudaErrorMemoryAllocation ? QP_ERR_OOM : QP_ERR_CUDA,            \
                     "qp_gen_create (dictionary): %s failed: %s", #call, cudaGetErrorString(e__)); \
    }                                                                                             \
  } while (0)

  D_CUDA(cudaMalloc(&d_tab, sizeof(DictSlot) * QP_DICT_HASH_CAP));
  D_CUDA(cudaMalloc(&d_ctl, sizeof(unsigned int) * 2));
  D_CUDA(cudaMalloc(&d_slot_of, sizeof(uint16_t) * (size_t)nnz));
  D_CUDA(cudaMemsetAsync(d_tab, 0, sizeof(DictSlot) * QP_DICT_HASH_CAP, ctx->stream));
  D_CUDA(cudaMemsetAsync(d_ctl, 0, sizeof(unsigned int) * 2, ctx->stream));
  {
    int64_t blocks = std::min<int64_t>((n * 32 + 255) / 256, (int64_t)ctx->sm_count * 8);
    k_dict_scan<<<(unsigned)blocks, 256, 0, ctx->stream>>>(g->d_mptr, g->d_mcolop, g->d_mval, n, d_tab, d_ctl,
                                                             (unsigned)(QP_DICT_MAX - 1), d_slot_of, skip_diag ? 1 : 0);
    ctx->launches++;
    D_CUDA(cudaGetLastError());
  }
  unsigned int h_ctl[2] = {0, 0};
  D_CUDA(cudaMemcpyAsync(h_ctl, d_ctl, sizeof(h_ctl), cudaMemcpyDeviceToHost, ctx->stream));
  D_CUDA(cudaStreamSynchronize(ctx->stream));
  if (h_ctl[1] != 0 || h_ctl[0] > (unsigned)(QP_DICT_MAX - 1)) {  // too many distinct entries
    release();
    return QP_OK;
  }

  // compact + order the table on the host: codes are deterministic (sorted by operator,
  // offset, value bits) although the hash slots are not
  std::vector<DictSlot> h_tab(QP_DICT_HASH_CAP);
  D_CUDA(cudaMemcpy(h_tab.data(), d_tab, sizeof(DictSlot) * QP_DICT_HASH_CAP, cudaMemcpyDeviceToHost));
  std::vector<int> used;
  for (int i = 0; i < QP_DICT_HASH_CAP; ++i)
    if (h_tab[i].tag == 2u) used.push_back(i);
  std::sort(used.begin(), used.end(), [&](int x, int y) {
    const DictSlot &p = h_tab[x], &q = h_tab[y];
    if (p.a != q.a) return p.a < q.a;
    if (p.b != q.b) return p.b < q.b;
    return p.c < q.c;
  });
  const int n_dict = (int)used.size() + 1;
  std::vector<uint16_t> h_remap(QP_DICT_HASH_CAP, 0);
  std::vector<double2> h_val(n_dict, make_double2(0.0, 0.0));
  std::vector<int32_t> h_delta(n_dict, 0);
  std::vector<uint8_t> h_op(n_dict, 0);
  for (int i = 0; i < (int)used.size(); ++i) {
    const DictSlot& p = h_tab[used[i]];
    h_remap[used[i]] = (uint16_t)(i + 1);
    double re, im;
    memcpy(&re, &p.b, 8);
    memcpy(&im, &p.c, 8);
    h_val[i + 1] = make_double2(re, im);
    h_delta[i + 1] = (int32_t)(uint32_t)(p.a & 0xffffffffull);
    h_op[i + 1] = (uint8_t)(p.a >> 32);
  }
  const int cb = n_dict <= 256 ? 1 : 2;
  const int cpw = 16 / cb;
  // operator kinds: purely real / purely imaginary operators let the batched kernel store one
  // real number per entry (the factor i moves into the coefficient)
  unsigned has_re = 0, has_im = 0;
  for (int i = 1; i < n_dict; ++i) {
    if (h_val[i].x != 0.0) has_re |= 1u << h_op[i];
    if (h_val[i].y != 0.0) has_im |= 1u << h_op[i];
  }
  const bool realv = (has_re & has_im) == 0u;
  std::vector<double> h_valr(n_dict, 0.0);
  for (int i = 1; i < n_dict; ++i) h_valr[i] = ((has_im >> h_op[i]) & 1u) ? h_val[i].y : h_val[i].x;

  D_CUDA(cudaMalloc(&d_remap, sizeof(uint16_t) * QP_DICT_HASH_CAP));
  D_CUDA(cudaMemcpy(d_remap, h_remap.data(), sizeof(uint16_t) * QP_DICT_HASH_CAP, cudaMemcpyHostToDevice));
  D_CUDA(cudaMalloc(&g->d_dval, sizeof(double2) * n_dict));
  D_CUDA(cudaMalloc(&g->d_ddelta, sizeof(int32_t) * n_dict));
  D_CUDA(cudaMalloc(&g->d_dop, sizeof(uint8_t) * n_dict));
  D_CUDA(cudaMemcpy(g->d_dval, h_val.data(), sizeof(double2) * n_dict, cudaMemcpyHostToDevice));
  D_CUDA(cudaMemcpy(g->d_ddelta, h_delta.data(), sizeof(int32_t) * n_dict, cudaMemcpyHostToDevice));
  D_CUDA(cudaMemcpy(g->d_dop, h_op.data(), sizeof(uint8_t) * n_dict, cudaMemcpyHostToDevice));
  D_CUDA(cudaMalloc(&g->d_dvalr, sizeof(double) * n_dict));
  D_CUDA(cudaMemcpy(g->d_dvalr, h_valr.data(), sizeof(double) * n_dict, cudaMemcpyHostToDevice));
  g->dict_realv = realv;
  g->imag_ops = realv ? has_im : 0u;
  g->dict_has_re = has_re;
  g->dict_has_im = has_im;

  // slice widths in 16-byte words
  D_CUDA(cudaMalloc(&d_words, sizeof(uint32_t) * (n_slices + 1)));
  D_CUDA(cudaMemsetAsync(d_words, 0, sizeof(uint32_t) * (n_slices + 1), ctx->stream));
  k_selld_widths<<<(unsigned)((n_slices * 32 + 255) / 256), 256, 0, ctx->stream>>>(g->d_mptr, n, n_slices, cpw, d_words);
  ctx->launches++;
  D_CUDA(cudaGetLastError());
  std::vector<uint32_t> h_words((size_t)n_slices);
  D_CUDA(cudaMemcpyAsync(h_words.data(), d_words, sizeof(uint32_t) * n_slices, cudaMemcpyDeviceToHost, ctx->stream));
  D_CUDA(cudaStreamSynchronize(ctx->stream));
  uint64_t total_words = 0;
  bool uniform = true;
  for (uint32_t v : h_words) {
    total_words += v;
    uniform &= (v == h_words[0]);
  }
  if (total_words >= (uint64_t(1) << 32)) {  // offsets are 32-bit
    release();
    drop();
    return QP_OK;
  }
  D_CUDA(cudaMalloc(&g->d_dptr, sizeof(uint32_t) * (n_slices + 1)));
  {
    int32_t rc = exclusive_scan_u32(ctx, d_words, g->d_dptr, n_slices + 1);
    if (rc != QP_OK) {
      release();
      drop();
      return rc;
    }
  }
  D_CUDA(cudaMalloc(&g->d_dcodes, sizeof(uint4) * std::max<uint64_t>(total_words, 1)));
  if (cb == 1)
    k_selld_fill<1><<<(unsigned)((n_slices * 32 + 255) / 256), 256, 0, ctx->stream>>>(g->d_mptr, d_slot_of, d_remap, n, n_slices, g->d_dptr, g->d_dcodes);
  else
    k_selld_fill<2><<<(unsigned)((n_slices * 32 + 255) / 256), 256, 0, ctx->stream>>>(g->d_mptr, d_slot_of, d_remap, n, n_slices, g->d_dptr, g->d_dcodes);
  ctx->launches++;
  D_CUDA(cudaGetLastError());
  if (skip_diag) {
    int* d_flags = nullptr;  // [0..15] has-diagonal flags, [16..31] index of the operator's vector
    D_CUDA(cudaMalloc(&d_flags, sizeof(int) * 2 * QP_MAX_OPS));
    cudaMemsetAsync(d_flags, 0, sizeof(int) * 2 * QP_MAX_OPS, ctx->stream);
    const unsigned blocks = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 16);
    k_diag_flags<<<blocks, 256, 0, ctx->stream>>>(g->d_mptr, g->d_mcolop, n, d_flags);
    ctx->launches++;
    int h_flags[2 * QP_MAX_OPS] = {0};
    cudaError_t e1 = cudaMemcpyAsync(h_flags, d_flags, sizeof(int) * QP_MAX_OPS, cudaMemcpyDeviceToHost, ctx->stream);
    if (e1 == cudaSuccess) e1 = cudaStreamSynchronize(ctx->stream);
    g->n_diag = 0;
    for (int l = 0; l < g->n_ops && e1 == cudaSuccess; ++l)
      if (h_flags[l]) {
        h_flags[QP_MAX_OPS + l] = g->n_diag;
        g->diag_op[g->n_diag++] = (uint8_t)l;
      }
    if (e1 == cudaSuccess && g->n_diag > 0) {
      e1 = cudaMalloc(&g->d_diag, sizeof(double2) * (size_t)g->n_diag * (size_t)n);
      if (e1 == cudaSuccess) e1 = cudaMemsetAsync(g->d_diag, 0, sizeof(double2) * (size_t)g->n_diag * (size_t)n, ctx->stream);
      if (e1 == cudaSuccess)
        e1 = cudaMemcpyAsync(d_flags + QP_MAX_OPS, h_flags + QP_MAX_OPS, sizeof(int) * QP_MAX_OPS, cudaMemcpyHostToDevice, ctx->stream);
      if (e1 == cudaSuccess) {
        k_diag_extract<<<blocks, 256, 0, ctx->stream>>>(g->d_mptr, g->d_mcolop, g->d_mval, n, d_flags + QP_MAX_OPS, g->d_diag);
        ctx->launches++;
        e1 = cudaGetLastError();
      }
    }
    if (e1 == cudaSuccess) e1 = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_flags);
    D_CUDA(e1);
  }
  D_CUDA(cudaStreamSynchronize(ctx->stream));
#undef D_CUDA
  release();
  g->n_dict = n_dict;
  g->code_bytes = cb;
  g->uniform_words = uniform && n_slices > 0 ? h_words[0] : 0u;
  g->tail_codes = 0;
  if (g->uniform_words > 0) {  // codes of the longest row in its last word, rounded up to even
    unsigned int* d_max = nullptr;
    unsigned int h_max = 0;
    if (cudaMalloc(&d_max, sizeof(unsigned int)) == cudaSuccess) {
      cudaMemsetAsync(d_max, 0, sizeof(unsigned int), ctx->stream);
      k_max_row_len<<<(unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 8), 256, 0, ctx->stream>>>(g->d_mptr, n, d_max);
      ctx->launches++;
      if (cudaMemcpyAsync(&h_max, d_max, sizeof(h_max), cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess &&
          cudaStreamSynchronize(ctx->stream) == cudaSuccess && h_max > 0) {
        const unsigned words_per_row = g->uniform_words / QP_SELL_C;
        if ((h_max + cpw - 1) / cpw == words_per_row) {
          const int tail = (int)(h_max - (words_per_row - 1) * cpw);
          g->tail_codes = tail >= cpw ? 0 : (tail + 1) / 2 * 2;
          if (g->tail_codes >= cpw || g->tail_codes == 8) g->tail_codes = 0;
        }
      }
      cudaGetLastError();
      cudaFree(d_max);
    }
  }
  g->dict_words = (int64_t)total_words;
  *ok = true;
  return QP_OK;
}

extern "C" int32_t qp_gen_create(qp_ctx_t ctx, int32_t n_ops, const qp_op_t* ops, int32_t n_coeffs,
                                 int32_t format, qp_gen_t* out) {
  QP_CHECK(qp_ctx_bind(ctx));
  QP_REQUIRE(ctx, out != nullptr && ops != nullptr, "qp_gen_create: null argument");
  *out = nullptr;
  QP_REQUIRE(ctx, n_ops >= 1, "qp_gen_create: need at least one operator");
  if (n_ops > QP_MAX_OPS)
    return qp_fail(ctx, QP_ERR_UNSUPPORTED, "qp_gen_create: %d operators exceed the limit of %d", n_ops, QP_MAX_OPS);
  // "The number of coefficients cannot exceed the number of operators" src/generators.jl:116-121
  QP_REQUIRE(ctx, n_coeffs >= 0 && n_coeffs <= n_ops,
             "qp_gen_create: the number of coefficients (%d) cannot exceed the number of operators (%d)", n_coeffs, n_ops);
  QP_REQUIRE(ctx, format >= QP_FORMAT_AUTO && format <= QP_FORMAT_BITFLIP, "qp_gen_create: bad format %d", format);
  bool any_dense = false, all_dense = true, any_lr = false, all_lr = true;
  for (int l = 0; l < n_ops; ++l) {
    QP_REQUIRE(ctx, ops[l] != nullptr, "qp_gen_create: operator %d is null", l);
    QP_REQUIRE(ctx, ops[l]->ctx == ctx, "qp_gen_create: operator %d belongs to another context", l);
    QP_REQUIRE(ctx, ops[l]->nrows == ops[l]->ncols, "qp_gen_create: operator %d is not square", l);
    QP_REQUIRE(ctx, ops[l]->nrows == ops[0]->nrows, "qp_gen_create: operator %d has a different size", l);
    any_dense |= ops[l]->dense;
    all_dense &= ops[l]->dense;
    any_lr |= ops[l]->leftright;
    all_lr &= ops[l]->leftright;
  }
  if (any_lr && !all_lr)
    return qp_fail(ctx, QP_ERR_UNSUPPORTED, "qp_gen_create: mixing matrix-free (left/right) operators and matrices is not supported");
  if ((format == QP_FORMAT_LR) != all_lr && !(all_lr && (format == QP_FORMAT_AUTO || format == QP_FORMAT_BITFLIP)))
    return qp_fail(ctx, QP_ERR_INVALID_ARG, "qp_gen_create: QP_FORMAT_LR is the format of left/right operators (and only theirs)");
  if (any_dense && !all_dense)
    return qp_fail(ctx, QP_ERR_UNSUPPORTED, "qp_gen_create: mixing dense and sparse operators is not supported");

  qp_gen_t g = new qp_gen_s();
  g->ctx = ctx;
  g->n_ops = n_ops;
  g->n_coeffs = n_coeffs;
  g->drift = n_ops - n_coeffs;
  g->n = ops[0]->nrows;
  g->ops.assign(ops, ops + n_ops);
  const int64_t n = g->n;

  auto bail = [&](int32_t rc) {
    gen_free(g);
    return rc;
  };
#define G_CUDA(call)                                                                             \
  do {                                                                                           \
    cudaError_t e__ = (call);                                                                    \
    if (e__ != cudaSuccess) {                                                                    \
      cudaGetLastError();                                                                        \
      return bail(qp_fail(ctx, e__ == cudaErrorMemoryAllocation ? QP_ERR_OOM : QP_ERR_CUDA,      \
                          "qp_gen_create: %s failed: %s", #call, cudaGetErrorString(e__)));      \
    }                                                                                            \
  } while (0)

  if (all_lr) {  // matrix-free: one table of terms over all operators
    g->format = QP_FORMAT_LR;
    g->lr_n = ops[0]->lr_n;
    std::vector<LRTerm> h;
    for (int l = 0; l < n_ops; ++l) {
      if (ops[l]->lr_n != g->lr_n) return bail(qp_fail(ctx, QP_ERR_INVALID_ARG, "qp_gen_create: operator %d acts on a different matrix size", l));
      for (const LRTermHost& T : ops[l]->lr_terms) {
        LRTerm d;
        memset(&d, 0, sizeof(d));
        if (T.left) {
          d.lptr = T.left->d_ptr;
          d.lcol = T.left->d_col;
          d.lval = T.left->d_val;
          g->matrix_bytes += 20 * T.left->nnz + 4 * (g->lr_n + 1);
          g->nnz_total += T.left->nnz;
        }
        if (T.d_rptr) {
          d.rptr = T.d_rptr;
          d.rcol = T.d_rcol;
          d.rval = T.d_rval;
          g->matrix_bytes += 20 * T.r_nnz + 4 * (g->lr_n + 1);
          g->nnz_total += T.r_nnz;
        }
        d.c = T.c;
        d.op = l;
        h.push_back(d);
      }
    }
    if ((int)h.size() > QP_LR_MAX_TERMS)
      return bail(qp_fail(ctx, QP_ERR_UNSUPPORTED, "qp_gen_create: %d left/right terms exceed the limit of %d", (int)h.size(), QP_LR_MAX_TERMS));
    g->n_lr_terms = (int)h.size();
    G_CUDA(cudaMalloc(&g->d_lr_terms, sizeof(LRTerm) * h.size()));
    G_CUDA(cudaMemcpy(g->d_lr_terms, h.data(), sizeof(LRTerm) * h.size(), cudaMemcpyHostToDevice));
    g->stored_entries = g->nnz_total;
    g->stored_bytes = g->matrix_bytes;
    // factors that are diagonals + (conditional) bit flips: the n^2 x n^2 operator has the bit-flip form too
    // (bitflip.cu), without ever being built -- thread-per-row kernel instead of the generic factor walk
    if ((format == QP_FORMAT_AUTO && g->n >= (int64_t)ctx->sm_count * 1024 && !getenv("QPROP_NO_BITFLIP")) || format == QP_FORMAT_BITFLIP) {
      bool bitflip_ok = false;
      int32_t rc = qp_bitflip_build(g, &bitflip_ok);
      if (rc == QP_OK && format == QP_FORMAT_BITFLIP && !bitflip_ok)
        rc = qp_fail(ctx, QP_ERR_UNSUPPORTED,
                     "qp_gen_create: QP_FORMAT_BITFLIP needs left/right factors that are diagonals plus uniform bit flips, no flip "
                     "multiplied by a non-constant diagonal, n a power of two");
      if (rc != QP_OK) return bail(rc);
      if (bitflip_ok) {
        g->format = QP_FORMAT_BITFLIP;
        g->stored_bytes = qp_bitflip_stored_bytes(g->bitflip);
      }
    }
    *out = g;
    return QP_OK;
  }
  if (all_dense) {
    if (format != QP_FORMAT_AUTO && format != QP_FORMAT_DENSE)
      return bail(qp_fail(ctx, QP_ERR_INVALID_ARG, "qp_gen_create: dense operators need QP_FORMAT_DENSE"));
    g->format = QP_FORMAT_DENSE;
    std::vector<const double2*> h(n_ops);
    for (int l = 0; l < n_ops; ++l) h[l] = ops[l]->d_dense;
    G_CUDA(cudaMalloc((void**)&g->d_dense_ops, sizeof(double2*) * n_ops));
    G_CUDA(cudaMemcpy((void*)g->d_dense_ops, h.data(), sizeof(double2*) * n_ops, cudaMemcpyHostToDevice));
    g->nnz_total = (int64_t)n_ops * n * n;
    g->stored_entries = g->nnz_total;
    g->matrix_bytes = 16 * g->nnz_total;
    *out = g;
    return QP_OK;
  }
  if (format == QP_FORMAT_DENSE)
    return bail(qp_fail(ctx, QP_ERR_INVALID_ARG, "qp_gen_create: QP_FORMAT_DENSE needs dense operators"));

  int64_t nnz_total = 0;
  OpPtrs P;
  memset(&P, 0, sizeof(P));
  for (int l = 0; l < n_ops; ++l) {
    nnz_total += ops[l]->nnz;
    P.ptr[l] = ops[l]->d_ptr;
    P.col[l] = ops[l]->d_col;
    P.val[l] = ops[l]->d_val;
    g->matrix_bytes += 20 * ops[l]->nnz + 4 * (n + 1);
  }
  if (nnz_total >= (int64_t(1) << 31))
    return bail(qp_fail(ctx, QP_ERR_UNSUPPORTED, "qp_gen_create: total nnz %lld exceeds 2^31", (long long)nnz_total));
  g->nnz_total = nnz_total;

  // ---- merged CSR
  uint32_t* d_len = nullptr;
  G_CUDA(cudaMalloc(&d_len, sizeof(uint32_t) * (n + 1)));
  G_CUDA(cudaMemsetAsync(d_len, 0, sizeof(uint32_t) * (n + 1), ctx->stream));
  G_CUDA(cudaMalloc(&g->d_mptr, sizeof(uint32_t) * (n + 1)));
  G_CUDA(cudaMalloc(&g->d_mcolop, sizeof(uint32_t) * std::max<int64_t>(nnz_total, 1)));
  G_CUDA(cudaMalloc(&g->d_mval, sizeof(double2) * std::max<int64_t>(nnz_total, 1)));
  k_merged_len<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(P, n_ops, n, d_len);
  ctx->launches++;
  G_CUDA(cudaGetLastError());
  {
    int32_t rc = exclusive_scan_u32(ctx, d_len, g->d_mptr, n + 1);
    if (rc != QP_OK) {
      cudaFree(d_len);
      return bail(rc);
    }
  }
  k_merge_rows<<<(unsigned)((n * 32 + 255) / 256), 256, 0, ctx->stream>>>(P, n_ops, n, g->d_mptr, g->d_mcolop, g->d_mval);
  ctx->launches++;
  G_CUDA(cudaGetLastError());
  // operators that share columns (quadrature control pairs): order the rows for the pair kernel
  if (n_ops >= 2 && n_ops <= 3 && n <= (int64_t(1) << 21) && !getenv("QPROP_NO_PAIRS")) {
    unsigned long long* d_np = nullptr;
    G_CUDA(cudaMalloc(&d_np, sizeof(unsigned long long)));
    G_CUDA(cudaMemsetAsync(d_np, 0, sizeof(unsigned long long), ctx->stream));
    k_pair_order_rows<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(g->d_mptr, g->d_mcolop, g->d_mval, n, d_np);
    ctx->launches++;
    unsigned long long h_np = 0;
    cudaError_t e1 = cudaMemcpyAsync(&h_np, d_np, sizeof(h_np), cudaMemcpyDeviceToHost, ctx->stream);
    if (e1 == cudaSuccess) e1 = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_np);
    G_CUDA(e1);
    g->pair_ordered = 8 * (int64_t)h_np >= nnz_total;  // worth it when >= 1/4 of the entries are paired
  }

  // ---- SELL-32 slice widths (cheap; also tells us the padding overhead)
  const int64_t n_slices = (n + QP_SELL_C - 1) / QP_SELL_C;
  g->n_slices = n_slices;
  uint32_t* d_slice_entries = nullptr;
  G_CUDA(cudaMalloc(&d_slice_entries, sizeof(uint32_t) * (n_slices + 1)));
  G_CUDA(cudaMemsetAsync(d_slice_entries, 0, sizeof(uint32_t) * (n_slices + 1), ctx->stream));
  k_sell_widths<<<(unsigned)((n_slices * 32 + 255) / 256), 256, 0, ctx->stream>>>(g->d_mptr, n, n_slices, d_slice_entries);
  ctx->launches++;
  G_CUDA(cudaGetLastError());
  G_CUDA(cudaMalloc(&g->d_sptr, sizeof(uint32_t) * (n_slices + 1)));
  uint64_t sell_entries = 0;  // 64-bit total on the host: the padded size may exceed 2^32
  {
    std::vector<uint32_t> h_slice((size_t)n_slices);
    G_CUDA(cudaMemcpyAsync(h_slice.data(), d_slice_entries, sizeof(uint32_t) * n_slices, cudaMemcpyDeviceToHost, ctx->stream));
    G_CUDA(cudaStreamSynchronize(ctx->stream));
    for (uint32_t v : h_slice) sell_entries += v;
  }

  // ---- format selection
  const double mean_len = (double)nnz_total / (double)n;
  const double pad = nnz_total > 0 ? (double)sell_entries / (double)nnz_total : 1.0;
  int chosen = format;
  // thread-per-row needs enough rows to fill the machine; otherwise use sub-warps per row
  const bool rows_ok = n >= (int64_t)ctx->sm_count * 1024;
  bool dict_ok = false;
  // attempted for every AUTO generator: besides the B = 1 SELL-D kernel (large N only) the
  // dictionary also serves the trajectory-batched kernel at any N
  if ((format == QP_FORMAT_AUTO && !getenv("QPROP_NO_DICT")) || format == QP_FORMAT_SELLD || format == QP_FORMAT_BITFLIP) {
    int32_t rc = build_dict(ctx, g, &dict_ok, false);
    // typical second chance: a diagonal drift term with thousands of distinct energies on top of
    // couplings with a handful of values -- keep the diagonals as explicit vectors
    if (rc == QP_OK && !dict_ok) rc = build_dict(ctx, g, &dict_ok, true);
    if (rc != QP_OK) {
      cudaFree(d_len);
      cudaFree(d_slice_entries);
      return bail(rc);
    }
    if (format == QP_FORMAT_SELLD && !dict_ok) {
      cudaFree(d_len);
      cudaFree(d_slice_entries);
      return bail(qp_fail(ctx, QP_ERR_UNSUPPORTED,
                          "qp_gen_create: QP_FORMAT_SELLD needs fewer than %d distinct (operator, value, offset) entries",
                          QP_DICT_MAX));
    }
  }
  // bit-flip form (bitflip.cu): diagonal operators + uniform XOR stencils, no matrix stream at all
  bool bitflip_ok = false;
  if ((format == QP_FORMAT_AUTO && rows_ok && !getenv("QPROP_NO_BITFLIP")) || format == QP_FORMAT_BITFLIP) {
    int32_t rc = qp_bitflip_build(g, &bitflip_ok);
    if (rc == QP_OK && format == QP_FORMAT_BITFLIP && !bitflip_ok)
      rc = qp_fail(ctx, QP_ERR_UNSUPPORTED,
                   "qp_gen_create: QP_FORMAT_BITFLIP needs operators that are diagonal or the same set of (row XOR mask, value) "
                   "entries in every row, N a multiple of 32");
    if (rc != QP_OK) {
      cudaFree(d_len);
      cudaFree(d_slice_entries);
      return bail(rc);
    }
  }
  if (chosen == QP_FORMAT_AUTO && bitflip_ok) chosen = QP_FORMAT_BITFLIP;
  if (chosen == QP_FORMAT_AUTO) {
    const bool sell_ok = rows_ok && pad <= 1.25 && sell_entries < (uint64_t(1) << 32);
    // padded code slots per true nonzero (the code stream is 1-2 B/entry, so padding is cheap in
    // bytes; the bound limits wasted lookups)
    const double dpad = dict_ok ? (double)g->dict_words * (16 / g->code_bytes) / (double)nnz_total : 1e30;
    chosen = (dict_ok && rows_ok && dpad <= 2.0) ? QP_FORMAT_SELLD : sell_ok ? QP_FORMAT_SELL : QP_FORMAT_CSR;
  }
  if (chosen == QP_FORMAT_AUTO) chosen = QP_FORMAT_CSR;
  g->format = chosen;
  {
    int lanes = pow2_floor((int64_t)std::max(1.0, mean_len / 2.0));
    lanes = std::min(32, std::max(1, lanes));
    // small problems: trade lane efficiency for parallelism
    while (lanes < 32 && n * lanes < (int64_t)ctx->sm_count * 2048) lanes *= 2;
    g->lanes = lanes;
    // SELL kernel selection (tuning knobs; defaults chosen from measurements on B200)
    if (const char* k = getenv("QPROP_SELL_KERNEL")) g->sell_kernel = (strcmp(k, "ldg") == 0) ? 0 : 1;
    if (const char* c = getenv("QPROP_TMA_CFG")) g->tma_cfg = atoi(c);
  }

  if (chosen == QP_FORMAT_SELL) {
    int32_t rc = exclusive_scan_u32(ctx, d_slice_entries, g->d_sptr, n_slices + 1);
    if (rc != QP_OK) {
      cudaFree(d_len);
      cudaFree(d_slice_entries);
      return bail(rc);
    }
    G_CUDA(cudaMalloc(&g->d_scolop, sizeof(uint32_t) * std::max<uint64_t>(sell_entries, 1)));
    G_CUDA(cudaMalloc(&g->d_sval, sizeof(double2) * std::max<uint64_t>(sell_entries, 1)));
    k_sell_fill<<<(unsigned)((n_slices * 32 + 255) / 256), 256, 0, ctx->stream>>>(
        g->d_mptr, g->d_mcolop, g->d_mval, n, n_slices, g->d_sptr, g->d_scolop, g->d_sval);
    ctx->launches++;
    G_CUDA(cudaGetLastError());
    g->stored_entries = (int64_t)sell_entries;
  } else {
    cudaFree(g->d_sptr);
    g->d_sptr = nullptr;
    g->stored_entries = nnz_total;
  }
  if (chosen == QP_FORMAT_BITFLIP) {
    g->stored_entries = nnz_total;
    g->stored_bytes = qp_bitflip_stored_bytes(g->bitflip);
  } else if (chosen == QP_FORMAT_SELLD) {
    g->stored_entries = g->dict_words * (16 / g->code_bytes);
    g->stored_bytes = g->dict_words * 16 + (g->uniform_words ? 0 : 4 * (n_slices + 1)) + 16 * (int64_t)g->n_diag * n;
  } else {
    g->stored_bytes = 20 * g->stored_entries + 4 * ((chosen == QP_FORMAT_SELL ? n_slices : n) + 1);
  }
  G_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_len);
  cudaFree(d_slice_entries);
#undef G_CUDA
  *out = g;
  return QP_OK;
}

extern "C" int32_t qp_gen_destroy(qp_gen_t gen) {
  if (!gen) return QP_OK;
  cudaSetDevice(gen->ctx->device);
  cudaStreamSynchronize(gen->ctx->stream);
  gen_free(gen);
  return QP_OK;
}

extern "C" int32_t qp_gen_info(qp_gen_t gen, int32_t* format, int64_t* n, int64_t* stored_entries,
                               int64_t* matrix_bytes) {
  if (!gen) return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_gen_info: null generator");
  if (format) *format = gen->format;
  if (n) *n = gen->n;
  if (stored_entries) *stored_entries = gen->stored_entries;
  if (matrix_bytes) *matrix_bytes = gen->matrix_bytes;
  return QP_OK;
}

extern "C" int32_t qp_gen_storage(qp_gen_t gen, int64_t* stored_bytes, int32_t* n_dict, int32_t* code_bytes) {
  if (!gen) return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_gen_storage: null generator");
  if (stored_bytes) *stored_bytes = gen->format == QP_FORMAT_DENSE ? gen->matrix_bytes : gen->stored_bytes;
  if (n_dict) *n_dict = gen->n_dict;
  if (code_bytes) *code_bytes = gen->n_dict > 0 ? gen->code_bytes : 0;
  return QP_OK;
}

// ---------------------------------------------------------------------------------------
// per-step coefficients
// ---------------------------------------------------------------------------------------

int32_t qp_gen_set_coeffs(qp_gen_t gen, const qp_c128* op_coeffs, int per_traj, int64_t batch,
                          int* coef_stride_out) {
  qp_ctx_t ctx = gen->ctx;
  QP_REQUIRE(ctx, gen->n_coeffs == 0 || op_coeffs != nullptr, "operator coefficients: null pointer");
  const int64_t width = per_traj ? batch : 1;
  const size_t elems = (size_t)gen->n_ops * (size_t)width;
  if (gen->coef_elems < elems) {
    QP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(gen->d_coef);
    gen->d_coef = nullptr;
    gen->coef_elems = 0;
    QP_CUDA(ctx, cudaMalloc(&gen->d_coef, sizeof(double2) * elems));
    gen->coef_elems = elems;
  }
  // the staging buffer is reused by every step: make sure the previous upload has drained
  // (cheap: the copy is the first thing in the stream of the previous step)
  QP_CHECK(qp_ctx_reserve_stage(ctx, elems));
  QP_CUDA(ctx, cudaEventSynchronize(ctx->ev_stage));
  qp_c128* h = ctx->h_stage;
  for (int l = 0; l < gen->n_ops; ++l)
    for (int64_t b = 0; b < width; ++b) {
      if (l < gen->drift)
        h[l * width + b] = qp_c128{1.0, 0.0};
      else
        h[l * width + b] = op_coeffs[(size_t)(l - gen->drift) * width + b];
    }
  gen->h_coef_valid = !per_traj;
  if (!per_traj) {
    gen->h_coef.resize(gen->n_ops);
    for (int l = 0; l < gen->n_ops; ++l) gen->h_coef[l] = make_double2(h[l].re, h[l].im);
  }
  QP_CUDA(ctx, cudaMemcpyAsync(gen->d_coef, h, sizeof(double2) * elems, cudaMemcpyHostToDevice, ctx->stream));
  QP_CUDA(ctx, cudaEventRecord(ctx->ev_stage, ctx->stream));
  if (coef_stride_out) *coef_stride_out = per_traj ? 1 : 0;
  return QP_OK;
}

// ---------------------------------------------------------------------------------------
// kernel dispatch
// ---------------------------------------------------------------------------------------

// Deterministic sums of the single-state kernels (normalization check, fused expectation value): the
// launch gets a zeroed buffer of one slot per warp, every warp stores its partial sums there, and
// part_end adds the slots up in a fixed order into e.chk.  QPROP_ATOMIC_SUMS=1 keeps the atomics.
template <int EPI>
static int32_t part_begin(qp_gen_t gen, EpiArgs& e, int64_t n_warps) {
  e.part = nullptr;
  if (!epi_has_sums(EPI) || e.chk == nullptr) return QP_OK;
  static const int atomic = getenv("QPROP_ATOMIC_SUMS") ? atoi(getenv("QPROP_ATOMIC_SUMS")) : 0;
  if (atomic) return QP_OK;
  qp_ctx_t ctx = gen->ctx;
  QP_CHECK(qp_ctx_reserve_part(ctx, (size_t)3 * (size_t)n_warps));
  QP_CUDA(ctx, cudaMemsetAsync(ctx->d_part, 0, sizeof(double) * 3 * (size_t)n_warps, ctx->stream));
  e.part = ctx->d_part;
  return QP_OK;
}
template <int EPI>
static int32_t part_end(qp_gen_t gen, const EpiArgs& e, int64_t n_warps) {
  if (e.part == nullptr) return QP_OK;
  return qp_part_reduce(gen->ctx, e.part, n_warps, 1, e.chk);
}

// TMA-staged SELL kernel: one CTA per SM (or a small multiple when a CTA's slice range would
// exceed the staged offset table), dynamic shared memory = stages + barriers + slice offsets.
template <int EPI, int WARPS, int STAGES, int CH>
static int32_t launch_sell_tma(qp_gen_t gen, const MatView& m, const double2* x, const EpiArgs& e) {
  qp_ctx_t ctx = gen->ctx;
  auto kern = k_spmv_sell_tma<EPI, WARPS, STAGES, CH>;
  int64_t ctas = ctx->sm_count;
  while ((gen->n_slices + ctas - 1) / ctas > TMA_MAX_LOCAL_SLICES) ctas += ctx->sm_count;
  if (ctas > gen->n_slices) ctas = gen->n_slices;
  const int spc = (int)((gen->n_slices + ctas - 1) / ctas);
  const size_t smem = sizeof(TmaStage<CH>) * WARPS * STAGES + sizeof(uint64_t) * WARPS * STAGES +
                      sizeof(uint32_t) * (size_t)(spc + 1);
  if (!ctx->smem_configured.count((const void*)kern)) {  // opt in to > 48 KB dynamic smem once per context
    QP_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    ctx->smem_configured.insert((const void*)kern);
  }
  if (smem > 226 * 1024) return qp_fail(ctx, QP_ERR_INTERNAL, "TMA kernel needs %zu bytes of shared memory", smem);
  EpiArgs e2 = e;
  QP_CHECK(part_begin<EPI>(gen, e2, ctas * WARPS));
  kern<<<(unsigned)ctas, WARPS * 32, smem, ctx->stream>>>(m, gen->d_coef, gen->n_ops, x, e2, spc);
  QP_LAUNCHED(ctx);
  return part_end<EPI>(gen, e2, ctas * WARPS);
}

static DictView make_dict_view(qp_gen_t gen) {
  DictView m{gen->d_dptr, gen->d_dcodes, gen->d_dval, gen->d_ddelta, gen->d_dop, gen->n_dict, gen->uniform_words,
             gen->n,      gen->d_diag,   gen->n_diag, 0ull};
  for (int i = 0; i < gen->n_diag; ++i) m.diag_ops |= (unsigned long long)(gen->diag_op[i] & 15) << (4 * i);
  return m;
}

// SELL-D kernel: CTAs = SMs x resident CTAs per SM (or fewer for small matrices), each owning a
// contiguous slice range; dynamic shared memory = the coefficient-scaled table.
template <int EPI, int CB, int TAIL, int REALT>
static int32_t launch_selld(qp_gen_t gen, const DictView& m, const double2* x, const EpiArgs& e) {
  qp_ctx_t ctx = gen->ctx;
  auto kern = k_spmv_selld<EPI, CB, TAIL, REALT>;
  const size_t smem = (size_t)m.n_dict * (sizeof(double2) + sizeof(int32_t));
  if (!ctx->smem_configured.count((const void*)kern)) {
    QP_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    ctx->smem_configured.insert((const void*)kern);
  }
  // one CTA of 512 threads per SM: all 16 resident warps sweep ONE contiguous slice range (measured
  // 32.0 us per term on config 2 against 32.8 us for 2 x 256 and 33.5 us for 4 x 128)
  static const int threads_env = getenv("QPROP_SELLD_THREADS") ? atoi(getenv("QPROP_SELLD_THREADS")) : SELLD_THREADS;
  const int threads = (threads_env == 128 || threads_env == 256) ? threads_env : SELLD_THREADS;
  static const int per_sm = getenv("QPROP_SELLD_CTAS") ? atoi(getenv("QPROP_SELLD_CTAS")) : 0;
  int occ = per_sm;
  if (occ <= 0) {
    QP_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
    if (occ < 1) occ = 1;
  }
  const int wpc = threads / 32;
  int64_t ctas = (int64_t)ctx->sm_count * occ;
  int64_t spc = (gen->n_slices + ctas - 1) / ctas;
  spc = (spc + wpc - 1) / wpc * wpc;  // whole rounds of the CTA's warps
  // alternative: small aligned tiles handed out by the hardware block scheduler
  static const int spc_env = getenv("QPROP_SELLD_SPC") ? atoi(getenv("QPROP_SELLD_SPC")) : 0;
  if (spc_env > 0) spc = spc_env;
  ctas = (gen->n_slices + spc - 1) / spc;
  static const int pdl = getenv("QPROP_PDL") ? atoi(getenv("QPROP_PDL")) : 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)ctas);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  const double2* coef = gen->d_coef;
  const int spc_i = (int)spc;
  EpiArgs e2 = e;
  QP_CHECK(part_begin<EPI>(gen, e2, ctas * wpc));
  QP_CUDA(ctx, cudaLaunchKernelEx(&cfg, kern, m, coef, x, e2, spc_i));
  QP_LAUNCHED(ctx);
  return part_end<EPI>(gen, e2, ctas * wpc);
}

template <int EPI, int CB, int REALV, int T, int G, int LATE>
static int32_t launch_spmm_selld_t(qp_gen_t gen, int coef_stride, const double2* x, int64_t batch, const EpiArgs& e) {
  qp_ctx_t ctx = gen->ctx;
  const DictView m = make_dict_view(gen);
  const int64_t chunks = (batch + 32 * T - 1) / (32 * T);
  if (chunks > 65535) return qp_fail(ctx, QP_ERR_UNSUPPORTED, "batch too large for one launch");
  const size_t smem = ((size_t)gen->n_dict * ((REALV ? 8 : 16) + sizeof(DeltaOp)) + 127) / 128 * 128;
  auto kern = k_spmm_selld<EPI, CB, REALV, T, G, LATE>;
  if (!ctx->smem_configured.count((const void*)kern)) {
    QP_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    ctx->smem_configured.insert((const void*)kern);
  }
  static const int spc_env = getenv("QPROP_SPMM_SPC") ? atoi(getenv("QPROP_SPMM_SPC")) : 1;
  const int spc = spc_env > 0 ? spc_env : 1;
  dim3 grid((unsigned)((gen->n_slices + spc - 1) / spc), (unsigned)chunks);
  kern<<<grid, 256, smem, ctx->stream>>>(m, gen->d_dvalr, gen->imag_ops, gen->d_coef, coef_stride, batch, x, e, spc);
  QP_LAUNCHED(ctx);
  return QP_OK;
}

template <int EPI>
static int32_t launch_spmm_selld(qp_gen_t gen, int coef_stride, const double2* x, int64_t batch, const EpiArgs& e) {
  // The row decode (code words, table lookups, operator changes) is paid once per warp and row:
  // T trajectory chunks per lane amortise it.  Measured on config 3 (N = 2^16, B = 1024, 45 entries
  // per row): 5.72 / 3.59 / 2.86 ms per term for T = 1 / 2 / 4, i.e. ~4.3 ms / T of decode on top of
  // ~1.5 ms of per-trajectory work (the 16-byte gathers through L1: 46 GB per term).  T = 4 (128
  // trajectories per warp) when the batch is large enough and the chunk's slice of x
  // (N x 128 x 16 B; 134 MB on config 3, measured fine) does not exceed L2 by much;
  // QPROP_SPMM_T forces 1 / 2 / 4.
  static const int t_env = getenv("QPROP_SPMM_T") ? atoi(getenv("QPROP_SPMM_T")) : 0;
  int tsel = batch > 64 && (int64_t)gen->n * 128 * 16 <= (int64_t)144 << 20 ? 4 : batch > 32 ? 2 : 1;
  if (t_env == 1 || t_env == 2 || t_env == 4) tsel = t_env;
  if (tsel == 4 && gen->pair_ordered && gen->dict_realv && gen->n_ops <= 3) {
    const int64_t chunks = (batch + 127) / 128;
    if (chunks > 65535) return qp_fail(gen->ctx, QP_ERR_UNSUPPORTED, "batch too large for one launch");
    const DictView m = make_dict_view(gen);
    const size_t smem = ((size_t)gen->n_dict * (8 + sizeof(DeltaOp)) + 127) / 128 * 128;
    dim3 grid((unsigned)gen->n_slices, (unsigned)chunks);
    qp_ctx_t ctx = gen->ctx;
#define QP_PAIRS(CB, NOPS)                                                                                          \
  do {                                                                                                              \
    auto kern = k_spmm_selld_pairs<EPI, CB, NOPS>;                                                                  \
    if (!ctx->smem_configured.count((const void*)kern)) {                                                           \
      QP_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));            \
      ctx->smem_configured.insert((const void*)kern);                                                               \
    }                                                                                                               \
    kern<<<grid, 256, smem, ctx->stream>>>(m, gen->d_dvalr, gen->imag_ops, gen->d_coef, coef_stride, batch, x, e);  \
  } while (0)
    if (gen->code_bytes == 1) {
      if (gen->n_ops == 2) QP_PAIRS(1, 2);
      else QP_PAIRS(1, 3);
    } else {
      if (gen->n_ops == 2) QP_PAIRS(2, 2);
      else QP_PAIRS(2, 3);
    }
#undef QP_PAIRS
    QP_LAUNCHED(ctx);
    return QP_OK;
  }
#define QP_SPMM(CB, RV) \
  (tsel == 4 ? launch_spmm_selld_t<EPI, CB, RV, 4, 2, 1>(gen, coef_stride, x, batch, e) \
   : tsel == 2 ? launch_spmm_selld_t<EPI, CB, RV, 2, 4, 0>(gen, coef_stride, x, batch, e) \
               : launch_spmm_selld_t<EPI, CB, RV, 1, 8, 0>(gen, coef_stride, x, batch, e))
  if (gen->code_bytes == 1) return gen->dict_realv ? QP_SPMM(1, 1) : QP_SPMM(1, 0);
  return gen->dict_realv ? QP_SPMM(2, 1) : QP_SPMM(2, 0);
#undef QP_SPMM
}

template <int EPI>
static int32_t launch_epi(qp_gen_t gen, int coef_stride, const double2* x, int64_t batch, const EpiArgs& e) {
  qp_ctx_t ctx = gen->ctx;
  const int64_t n = gen->n;
  cudaStream_t st = ctx->stream;
  if (gen->format == QP_FORMAT_BITFLIP && gen->d_mptr == nullptr) {  // derived from matrix-free left/right operators
    if (batch != 1) return qp_fail(ctx, QP_ERR_UNSUPPORTED, "matrix-free left/right generators take single states (batch = %lld)", (long long)batch);
    return qp_launch_bitflip(gen, EPI, x, e);
  }
  if (gen->format == QP_FORMAT_LR) {
    if (batch != 1) return qp_fail(ctx, QP_ERR_UNSUPPORTED, "matrix-free left/right generators take single states (batch = %lld)", (long long)batch);
    // CTA = 256 rows of rho x a range of columns swept JB at a time; enough column ranges to fill
    // the machine a few times over
    constexpr int JB = 4;
    const int64_t nh = gen->lr_n;
    const int64_t row_blocks = (nh + 255) / 256;
    int64_t col_ranges = std::max<int64_t>(1, ((int64_t)ctx->sm_count * 6 + row_blocks - 1) / row_blocks);
    int64_t cols = (nh + col_ranges - 1) / col_ranges;
    cols = std::max<int64_t>(JB, (cols + JB - 1) / JB * JB);
    col_ranges = (nh + cols - 1) / cols;
    if (col_ranges > 65535) return qp_fail(ctx, QP_ERR_UNSUPPORTED, "left/right generator too large for one launch");
    dim3 grid((unsigned)row_blocks, (unsigned)col_ranges);
    EpiArgs e2 = e;
    QP_CHECK(part_begin<EPI>(gen, e2, row_blocks * col_ranges * 8));
    k_spmv_lr<EPI, JB><<<grid, 256, sizeof(LRTerm) * gen->n_lr_terms, st>>>(gen->d_lr_terms, gen->n_lr_terms, gen->n_ops, nh,
                                                                               gen->d_coef, x, e2, (int)cols);
    QP_LAUNCHED(ctx);
    return part_end<EPI>(gen, e2, row_blocks * col_ranges * 8);
  }
  if (gen->format == QP_FORMAT_DENSE) {
    if (batch != 1) return launch_dense_batched<EPI>(gen, coef_stride, x, batch, e);  // FP64 tensor cores
    EpiArgs e2 = e;
    static const int gemv_flat = getenv("QPROP_GEMV_FLAT") ? atoi(getenv("QPROP_GEMV_FLAT")) : 1;
    if (gemv_flat && n >= 2048 && n % 128 == 0) {  // the matrix as one contiguous stream + a row-sum pass (spmv.cuh)
      const int64_t ppr = n >> 7;
      QP_CHECK(qp_ctx_reserve_gemv(ctx, (size_t)2 * (size_t)gen->n_ops * (size_t)n * (size_t)ppr));
      double2* part = reinterpret_cast<double2*>(ctx->d_gemv);
      const int64_t fblocks = (int64_t)ctx->sm_count * 8;
      k_gemv_flat<0><<<(unsigned)fblocks, 256, 0, st>>>(gen->d_dense_ops, gen->n_ops, n, x, part);
      QP_LAUNCHED(ctx);
      const int64_t rblocks = (n + 255) / 256;
      QP_CHECK(part_begin<EPI>(gen, e2, rblocks * 8));
      k_gemv_flat_fin<EPI><<<(unsigned)rblocks, 256, 0, st>>>(part, gen->n_ops, n, gen->d_coef, x, e2);
      QP_LAUNCHED(ctx);
      return part_end<EPI>(gen, e2, rblocks * 8);
    }
    static const int gemv_cta = getenv("QPROP_GEMV_CTA") ? atoi(getenv("QPROP_GEMV_CTA")) : 1;
    if (gemv_cta && n >= 2048) {  // CTA per row, grid-stride: few long contiguous streams (spmv.cuh)
      const int64_t gblocks = std::min<int64_t>(n, (int64_t)ctx->sm_count * 8);
      QP_CHECK(part_begin<EPI>(gen, e2, gblocks * 8));
      k_gemv_dense_cta<EPI><<<(unsigned)gblocks, 256, 0, st>>>(gen->d_dense_ops, gen->n_ops, n, gen->d_coef, x, e2);
      QP_LAUNCHED(ctx);
      return part_end<EPI>(gen, e2, gblocks * 8);
    }
    const int64_t gblocks = (n * 32 + 255) / 256;
    QP_CHECK(part_begin<EPI>(gen, e2, gblocks * 8));
    k_gemv_dense<EPI><<<(unsigned)gblocks, 256, 0, st>>>(gen->d_dense_ops, gen->n_ops, n, gen->d_coef, x, e2);
    QP_LAUNCHED(ctx);
    return part_end<EPI>(gen, e2, gblocks * 8);
  }
  if (batch >= 32) {  // two-pass tiled path (tile.cu): structured generators, batch a multiple of 32
    bool handled = false;
    QP_CHECK(qp_launch_tile(gen, EPI, coef_stride, x, batch, e, &handled));
    if (handled) return QP_OK;
  }
  if (batch >= 16 && gen->n_dict > 0)  // dictionary available: warp per (row, 32 T trajectories)
    return launch_spmm_selld<EPI>(gen, coef_stride, x, batch, e);
  if (batch > 1) {
    MatView m{gen->d_mptr, gen->d_mcolop, gen->d_mval, n};
    int64_t blocks = (n * batch + 255) / 256;
    if (blocks >= (int64_t(1) << 31)) return qp_fail(ctx, QP_ERR_UNSUPPORTED, "n * batch too large for one launch");
    k_spmm_csr<EPI><<<(unsigned)blocks, 256, 0, st>>>(m, gen->d_coef, coef_stride, batch, x, e);
    QP_LAUNCHED(ctx);
    return QP_OK;
  }
  if (gen->format == QP_FORMAT_BITFLIP) return qp_launch_bitflip(gen, EPI, x, e);
  if (gen->format == QP_FORMAT_SELLD) {
    const DictView m = make_dict_view(gen);
    // compiled for 16 resident warps per SM (<= 128 registers): measured best on B200 against
    // 24 / 32 warps and a 16-gather variant (profiles/r1_variants.txt)
    // every slice has the same number of words and the longest row ends with `tail_codes` codes in
    // its last word: that word is decoded without its padding (even tails 2 / 4 / 6 of a group)
    static const int no_tail = getenv("QPROP_SELLD_TAIL") ? atoi(getenv("QPROP_SELLD_TAIL")) == 0 : 0;
    int tail = no_tail ? 0 : gen->tail_codes;
    if (tail > 6) tail = 0;
    // real table: every (coefficient x value) product of this step is real -- known on the host
    // when the coefficients were set from the host: operator by operator, real values with a
    // real coefficient or purely imaginary values with a purely imaginary coefficient
    static const int no_real = getenv("QPROP_SELLD_REAL") ? atoi(getenv("QPROP_SELLD_REAL")) == 0 : 0;
    bool real_tab = !no_real && gen->h_coef_valid && (int)gen->h_coef.size() == gen->n_ops && gen->n_diag == 0;
    for (int l = 0; real_tab && l < gen->n_ops; ++l) {
      const bool re = (gen->dict_has_re >> l) & 1u, im = (gen->dict_has_im >> l) & 1u;
      const double2 u = gen->h_coef[l];
      if (re && im) real_tab = false;
      else if (re) real_tab = u.y == 0.0;
      else if (im) real_tab = u.x == 0.0;
    }
#define QP_SELLD_TAILS(CB, RT)                                         \
  switch (tail) {                                                      \
    case 2: return launch_selld<EPI, CB, 2, RT>(gen, m, x, e);         \
    case 4: return launch_selld<EPI, CB, 4, RT>(gen, m, x, e);         \
    case 6: return launch_selld<EPI, CB, 6, RT>(gen, m, x, e);         \
    default: return launch_selld<EPI, CB, 0, RT>(gen, m, x, e);        \
  }
    if (gen->code_bytes == 1) {
      if (real_tab) QP_SELLD_TAILS(1, 1) else QP_SELLD_TAILS(1, 0)
    }
    if (real_tab) QP_SELLD_TAILS(2, 1) else QP_SELLD_TAILS(2, 0)
#undef QP_SELLD_TAILS
  }
  if (gen->format == QP_FORMAT_SELL) {
    MatView m{gen->d_sptr, gen->d_scolop, gen->d_sval, n};
    if (gen->sell_kernel == 1) {
      switch (gen->tma_cfg) {
        case 1: return launch_sell_tma<EPI, 8, 3, 8>(gen, m, x, e);
        case 2: return launch_sell_tma<EPI, 16, 2, 8>(gen, m, x, e);
        case 3: return launch_sell_tma<EPI, 12, 3, 8>(gen, m, x, e);
        case 4: return launch_sell_tma<EPI, 8, 6, 4>(gen, m, x, e);
        case 5: return launch_sell_tma<EPI, 16, 3, 4>(gen, m, x, e);
        case 6: return launch_sell_tma<EPI, 32, 2, 4>(gen, m, x, e);
        case 7: return launch_sell_tma<EPI, 24, 2, 4>(gen, m, x, e);
        case 8: return launch_sell_tma<EPI, 20, 2, 8>(gen, m, x, e);
        case 9: return launch_sell_tma<EPI, 24, 3, 4>(gen, m, x, e);
        case 10: return launch_sell_tma<EPI, 16, 4, 4>(gen, m, x, e);
        default: return launch_sell_tma<EPI, 8, 4, 8>(gen, m, x, e);
      }
    }
    // persistent grid: exactly the number of CTAs that are resident at once (SMs x occupancy),
    // each striding over the slices -- no partial second wave
    static int occ = 0;
    if (occ == 0) {
      QP_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_spmv_sell<EPI>, 256, 0));
      if (occ < 1) occ = 1;
    }
    int64_t blocks = (gen->n_slices + 7) / 8;
    int64_t cap = (int64_t)ctx->sm_count * occ;
    if (blocks > cap) blocks = cap;
    EpiArgs e2 = e;
    QP_CHECK(part_begin<EPI>(gen, e2, blocks * 8));
    k_spmv_sell<EPI><<<(unsigned)blocks, 256, 0, st>>>(m, gen->d_coef, gen->n_ops, x, e2);
    QP_LAUNCHED(ctx);
    return part_end<EPI>(gen, e2, blocks * 8);
  }
  MatView m{gen->d_mptr, gen->d_mcolop, gen->d_mval, n};
  const int lanes = gen->lanes;
  const unsigned blocks = (unsigned)((n * lanes + 255) / 256);
  // programmatic dependent launch (see k_spmv_csr): the next term's launch and matrix prologue
  // overlap the tail of this one; QPROP_PDL=0 disables it
  static const int pdl = getenv("QPROP_PDL") ? atoi(getenv("QPROP_PDL")) : 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(256);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  const double2* coef = gen->d_coef;
  const int n_ops = gen->n_ops;
  cudaError_t le = cudaSuccess;
  EpiArgs e2 = e;
  QP_CHECK(part_begin<EPI>(gen, e2, (int64_t)blocks * 8));
  switch (lanes) {
    case 1: le = cudaLaunchKernelEx(&cfg, k_spmv_csr<1, EPI>, m, coef, n_ops, x, e2); break;
    case 2: le = cudaLaunchKernelEx(&cfg, k_spmv_csr<2, EPI>, m, coef, n_ops, x, e2); break;
    case 4: le = cudaLaunchKernelEx(&cfg, k_spmv_csr<4, EPI>, m, coef, n_ops, x, e2); break;
    case 8: le = cudaLaunchKernelEx(&cfg, k_spmv_csr<8, EPI>, m, coef, n_ops, x, e2); break;
    case 16: le = cudaLaunchKernelEx(&cfg, k_spmv_csr<16, EPI>, m, coef, n_ops, x, e2); break;
    default: le = cudaLaunchKernelEx(&cfg, k_spmv_csr<32, EPI>, m, coef, n_ops, x, e2); break;
  }
  QP_CUDA(ctx, le);
  QP_LAUNCHED(ctx);
  return part_end<EPI>(gen, e2, (int64_t)blocks * 8);
}

int32_t qp_launch_fused(qp_gen_t gen, int epi, int coef_stride, const double2* x, int64_t batch,
                        const EpiArgs& e) {
  QpScopedTimer t(gen->ctx, "matrix-vector product");
  switch (epi) {
    case EPI_MUL: return launch_epi<EPI_MUL>(gen, coef_stride, x, batch, e);
    case EPI_CHEB_FIRST: return launch_epi<EPI_CHEB_FIRST>(gen, coef_stride, x, batch, e);
    case EPI_CHEB_MID: return launch_epi<EPI_CHEB_MID>(gen, coef_stride, x, batch, e);
    case EPI_CHEB_LAST: return launch_epi<EPI_CHEB_LAST>(gen, coef_stride, x, batch, e);
    case EPI_CHEB_ONLY: return launch_epi<EPI_CHEB_ONLY>(gen, coef_stride, x, batch, e);
    case EPI_DOT: return launch_epi<EPI_DOT>(gen, coef_stride, x, batch, e);
  }
  return qp_fail(gen->ctx, QP_ERR_INTERNAL, "bad epilogue %d", epi);
}

int32_t qp_gen_apply_scaled(qp_gen_t gen, int coef_stride, double2 alpha, const double* alpha_dev, double2 beta, const double2* x,
                            double2* y, int64_t batch) {
  EpiArgs e;
  memset(&e, 0, sizeof(e));
  e.alpha = alpha;
  e.alpha_dev = alpha_dev;
  e.betac = beta;
  e.y = y;
  return qp_launch_fused(gen, EPI_MUL, coef_stride, x, batch, e);
}

int32_t qp_gen_apply(qp_gen_t gen, int coef_stride, double2 alpha, double2 beta, const double2* x,
                     double2* y, int64_t batch) {
  return qp_gen_apply_scaled(gen, coef_stride, alpha, nullptr, beta, x, y, batch);
}

static int32_t check_gen_state(qp_gen_t gen, qp_state_t x, qp_state_t y, const char* what) {
  if (!gen) return qp_fail(nullptr, QP_ERR_INVALID_ARG, "%s: null generator", what);
  qp_ctx_t ctx = gen->ctx;
  QP_REQUIRE(ctx, x && y, "%s: null state", what);
  QP_REQUIRE(ctx, x->ctx == ctx && y->ctx == ctx, "%s: state belongs to another context", what);
  QP_REQUIRE(ctx, x->n == gen->n && y->n == gen->n, "%s: state dimension %lld does not match operator dimension %lld",
             what, (long long)x->n, (long long)gen->n);
  QP_REQUIRE(ctx, x->batch == y->batch, "%s: batch mismatch", what);
  return QP_OK;
}

extern "C" int32_t qp_gen_mul(qp_gen_t gen, const qp_c128* coeffs, qp_c128 alpha, qp_c128 beta,
                              qp_state_t x, qp_state_t y) {
  QP_CHECK(check_gen_state(gen, x, y, "qp_gen_mul"));
  qp_ctx_t ctx = gen->ctx;
  QP_CHECK(qp_ctx_bind(ctx));
  QP_REQUIRE(ctx, x->d != y->d, "qp_gen_mul: x and y must not alias");
  int stride = 0;
  QP_CHECK(qp_gen_set_coeffs(gen, coeffs, 0, x->batch, &stride));
  return qp_gen_apply(gen, stride, make_double2(alpha.re, alpha.im), make_double2(beta.re, beta.im),
                      x->d, y->d, x->batch);
}

extern "C" int32_t qp_gen_dot(qp_gen_t gen, const qp_c128* coeffs, qp_state_t x, qp_state_t y, qp_c128* out) {
  QP_CHECK(check_gen_state(gen, x, y, "qp_gen_dot"));
  qp_ctx_t ctx = gen->ctx;
  QP_CHECK(qp_ctx_bind(ctx));
  QP_REQUIRE(ctx, out != nullptr, "qp_gen_dot: null output");
  int stride = 0;
  QP_CHECK(qp_gen_set_coeffs(gen, coeffs, 0, x->batch, &stride));
  double2* tmp = nullptr;
  QP_CUDA(ctx, cudaMalloc(&tmp, sizeof(double2) * x->n * x->batch));
  int32_t rc = qp_gen_apply(gen, stride, make_double2(1.0, 0.0), make_double2(0.0, 0.0), y->d, tmp, x->batch);
  if (rc == QP_OK) rc = qp_reduce_dot(ctx, x->d, tmp, x->n, x->batch, out);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(tmp);
  return rc;
}

extern "C" int32_t qp_gen_expval(qp_gen_t gen, const qp_c128* coeffs, qp_state_t x, qp_c128* out) {
  QP_CHECK(check_gen_state(gen, x, x, "qp_gen_expval"));
  qp_ctx_t ctx = gen->ctx;
  QP_CHECK(qp_ctx_bind(ctx));
  QP_REQUIRE(ctx, out != nullptr, "qp_gen_expval: null output");
  int stride = 0;
  QP_CHECK(qp_gen_set_coeffs(gen, coeffs, 0, x->batch, &stride));
  const size_t need = (size_t)3 * x->batch;
  QP_CHECK(qp_ctx_reserve_red(ctx, need));
  QP_CUDA(ctx, cudaMemsetAsync(ctx->d_red, 0, sizeof(double) * need, ctx->stream));
  EpiArgs e;
  memset(&e, 0, sizeof(e));
  e.chk = ctx->d_red;
  QP_CHECK(qp_launch_fused(gen, EPI_DOT, stride, x->d, x->batch, e));
  QP_CUDA(ctx, cudaMemcpyAsync(ctx->h_red, ctx->d_red, sizeof(double) * need, cudaMemcpyDeviceToHost, ctx->stream));
  QP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int64_t b = 0; b < x->batch; ++b) out[b] = qp_c128{ctx->h_red[3 * b], ctx->h_red[3 * b + 1]};
  return QP_OK;
}
