// Operators and generators: upload (CSC/CSR Int64 -> device CSR Int32), device-side merge of
// the component operators into one tagged matrix, SELL-32 construction, format selection,
// and the host-side dispatcher of the fused SpMV kernels.
#include <algorithm>
#include <cmath>
#include <cstring>

#include <cub/device/device_scan.cuh>

#include "spmv.cuh"

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
  delete op;
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
  cudaFree((void*)g->d_dense_ops);
  cudaFree(g->d_coef);
  delete g;
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
  QP_REQUIRE(ctx, format >= QP_FORMAT_AUTO && format <= QP_FORMAT_DENSE, "qp_gen_create: bad format %d", format);
  bool any_dense = false, all_dense = true;
  for (int l = 0; l < n_ops; ++l) {
    QP_REQUIRE(ctx, ops[l] != nullptr, "qp_gen_create: operator %d is null", l);
    QP_REQUIRE(ctx, ops[l]->ctx == ctx, "qp_gen_create: operator %d belongs to another context", l);
    QP_REQUIRE(ctx, ops[l]->nrows == ops[l]->ncols, "qp_gen_create: operator %d is not square", l);
    QP_REQUIRE(ctx, ops[l]->nrows == ops[0]->nrows, "qp_gen_create: operator %d has a different size", l);
    any_dense |= ops[l]->dense;
    all_dense &= ops[l]->dense;
  }
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
  if (chosen == QP_FORMAT_AUTO) {
    // thread-per-row needs enough rows to fill the machine; otherwise use sub-warps per row
    const bool sell_ok = n >= (int64_t)ctx->sm_count * 1024 && pad <= 1.25 && sell_entries < (uint64_t(1) << 32);
    chosen = sell_ok ? QP_FORMAT_SELL : QP_FORMAT_CSR;
  }
  if (chosen == QP_FORMAT_SELL && sell_entries >= (uint64_t(1) << 32)) {
    cudaFree(d_len);
    cudaFree(d_slice_entries);
    return bail(qp_fail(ctx, QP_ERR_UNSUPPORTED, "qp_gen_create: SELL storage exceeds 2^32 entries"));
  }
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
  QP_CUDA(ctx, cudaMemcpyAsync(gen->d_coef, h, sizeof(double2) * elems, cudaMemcpyHostToDevice, ctx->stream));
  QP_CUDA(ctx, cudaEventRecord(ctx->ev_stage, ctx->stream));
  if (coef_stride_out) *coef_stride_out = per_traj ? 1 : 0;
  return QP_OK;
}

// ---------------------------------------------------------------------------------------
// kernel dispatch
// ---------------------------------------------------------------------------------------

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
  kern<<<(unsigned)ctas, WARPS * 32, smem, ctx->stream>>>(m, gen->d_coef, gen->n_ops, x, e, spc);
  QP_LAUNCHED(ctx);
  return QP_OK;
}

template <int EPI>
static int32_t launch_epi(qp_gen_t gen, int coef_stride, const double2* x, int64_t batch, const EpiArgs& e) {
  qp_ctx_t ctx = gen->ctx;
  const int64_t n = gen->n;
  cudaStream_t st = ctx->stream;
  if (gen->format == QP_FORMAT_DENSE) {
    if (batch != 1)
      return qp_fail(ctx, QP_ERR_UNSUPPORTED, "dense generators with batch > 1 are not supported yet");
    k_gemv_dense<EPI><<<(unsigned)((n * 32 + 255) / 256), 256, 0, st>>>(gen->d_dense_ops, gen->n_ops, n, gen->d_coef, x, e);
    QP_LAUNCHED(ctx);
    return QP_OK;
  }
  if (batch > 1) {
    MatView m{gen->d_mptr, gen->d_mcolop, gen->d_mval, n};
    int64_t blocks = (n * batch + 255) / 256;
    if (blocks >= (int64_t(1) << 31)) return qp_fail(ctx, QP_ERR_UNSUPPORTED, "n * batch too large for one launch");
    k_spmm_csr<EPI><<<(unsigned)blocks, 256, 0, st>>>(m, gen->d_coef, coef_stride, batch, x, e);
    QP_LAUNCHED(ctx);
    return QP_OK;
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
    k_spmv_sell<EPI><<<(unsigned)blocks, 256, 0, st>>>(m, gen->d_coef, gen->n_ops, x, e);
    QP_LAUNCHED(ctx);
    return QP_OK;
  }
  MatView m{gen->d_mptr, gen->d_mcolop, gen->d_mval, n};
  const int lanes = gen->lanes;
  const unsigned blocks = (unsigned)((n * lanes + 255) / 256);
  switch (lanes) {
    case 1: k_spmv_csr<1, EPI><<<blocks, 256, 0, st>>>(m, gen->d_coef, gen->n_ops, x, e); break;
    case 2: k_spmv_csr<2, EPI><<<blocks, 256, 0, st>>>(m, gen->d_coef, gen->n_ops, x, e); break;
    case 4: k_spmv_csr<4, EPI><<<blocks, 256, 0, st>>>(m, gen->d_coef, gen->n_ops, x, e); break;
    case 8: k_spmv_csr<8, EPI><<<blocks, 256, 0, st>>>(m, gen->d_coef, gen->n_ops, x, e); break;
    case 16: k_spmv_csr<16, EPI><<<blocks, 256, 0, st>>>(m, gen->d_coef, gen->n_ops, x, e); break;
    default: k_spmv_csr<32, EPI><<<blocks, 256, 0, st>>>(m, gen->d_coef, gen->n_ops, x, e); break;
  }
  QP_LAUNCHED(ctx);
  return QP_OK;
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
  }
  return qp_fail(gen->ctx, QP_ERR_INTERNAL, "bad epilogue %d", epi);
}

int32_t qp_gen_apply(qp_gen_t gen, int coef_stride, double2 alpha, double2 beta, const double2* x,
                     double2* y, int64_t batch) {
  EpiArgs e;
  memset(&e, 0, sizeof(e));
  e.alpha = alpha;
  e.betac = beta;
  e.y = y;
  return qp_launch_fused(gen, EPI_MUL, coef_stride, x, batch, e);
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
