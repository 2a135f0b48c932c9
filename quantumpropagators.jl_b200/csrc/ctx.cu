// Context, device-resident states and the level-1 verbs of the state interface
// (reference contract: src/interfaces/state.jl:24-47).
#include <cmath>
#include <cstring>

#include "qprop_internal.h"

// ---------------------------------------------------------------------------------------
// errors / context
// ---------------------------------------------------------------------------------------

static thread_local std::string g_null_ctx_err;

int32_t qp_fail(qp_ctx_t ctx, int32_t code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx)
    ctx->err = buf;
  else
    g_null_ctx_err = buf;
  return code;
}

extern "C" int32_t qp_version(void) { return QP_VERSION; }

extern "C" const char* qp_status_string(int32_t status) {
  switch (status) {
    case QP_OK: return "ok";
    case QP_ERR_INVALID_ARG: return "invalid argument";
    case QP_ERR_CUDA: return "CUDA error";
    case QP_ERR_OOM: return "out of memory";
    case QP_ERR_NOT_CONVERGED: return "not converged";
    case QP_ERR_NORMALIZATION: return "incorrect normalization";
    case QP_ERR_UNSUPPORTED: return "unsupported";
    case QP_ERR_INTERNAL: return "internal error";
    default: return "unknown status";
  }
}

extern "C" const char* qp_last_error(qp_ctx_t ctx) {
  return ctx ? ctx->err.c_str() : g_null_ctx_err.c_str();
}

int32_t qp_ctx_bind(qp_ctx_t ctx) {
  if (!ctx) return qp_fail(nullptr, QP_ERR_INVALID_ARG, "null context");
  QP_CUDA(ctx, cudaSetDevice(ctx->device));
  return QP_OK;
}

extern "C" int32_t qp_ctx_create(int32_t device, qp_ctx_t* out) {
  if (!out) return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_ctx_create: null output pointer");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    return qp_fail(nullptr, QP_ERR_CUDA,
                   "qp_ctx_create: no CUDA device available (%s); libqprop_b200 has no CPU fallback",
                   e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  }
  if (device < 0 || device >= count)
    return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_ctx_create: device %d out of range [0,%d)",
                   device, count);
  qp_ctx_t ctx = new qp_ctx_s();
  ctx->device = device;
  auto fail = [&](cudaError_t err, const char* what) {
    int32_t rc = qp_fail(nullptr, QP_ERR_CUDA, "qp_ctx_create: %s: %s", what, cudaGetErrorString(err));
    delete ctx;
    return rc;
  };
  if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(e, "cudaSetDevice");
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail(e, "cudaGetDeviceProperties");
  ctx->sm_count = prop.multiProcessorCount;
  if (prop.major < 10)
    return fail(cudaErrorInvalidDevice, "device is not sm_100-class (this library ships sm_100a code only)");
  if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess)
    return fail(e, "cudaStreamCreate");
  if ((e = cudaEventCreate(&ctx->ev0)) != cudaSuccess) return fail(e, "cudaEventCreate");
  if ((e = cudaEventCreate(&ctx->ev1)) != cudaSuccess) return fail(e, "cudaEventCreate");
  if ((e = cudaEventCreateWithFlags(&ctx->ev_stage, cudaEventDisableTiming)) != cudaSuccess)
    return fail(e, "cudaEventCreate");
  *out = ctx;
  return QP_OK;
}

void qp_ctx_release(qp_ctx_t ctx) {
  if (--ctx->refs == 0 && ctx->destroy_requested) qp_ctx_destroy(ctx);
}

extern "C" int32_t qp_ctx_destroy(qp_ctx_t ctx) {
  if (!ctx) return QP_OK;
  if (ctx->refs > 0) {  // objects still live on it: freed with the last of them
    ctx->destroy_requested = true;
    return QP_OK;
  }
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->d_red) cudaFree(ctx->d_red);
  if (ctx->d_part) cudaFree(ctx->d_part);
  if (ctx->d_gemv) cudaFree(ctx->d_gemv);
  if (ctx->h_red) cudaFreeHost(ctx->h_red);
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->ev_stage) cudaEventDestroy(ctx->ev_stage);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return QP_OK;
}

extern "C" int32_t qp_sync(qp_ctx_t ctx) {
  QP_CHECK(qp_ctx_bind(ctx));
  QP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return QP_OK;
}

extern "C" int32_t qp_ctx_stream(qp_ctx_t ctx, void** stream) {
  if (!ctx || !stream) return qp_fail(ctx, QP_ERR_INVALID_ARG, "qp_ctx_stream: null argument");
  *stream = (void*)ctx->stream;
  return QP_OK;
}

extern "C" int32_t qp_ctx_launch_count(qp_ctx_t ctx, int64_t* n) {
  if (!ctx || !n) return qp_fail(ctx, QP_ERR_INVALID_ARG, "qp_ctx_launch_count: null argument");
  *n = ctx->launches;
  return QP_OK;
}

int32_t qp_ctx_reserve_red(qp_ctx_t ctx, size_t doubles) {
  if (ctx->red_doubles >= doubles) return QP_OK;
  if (ctx->d_red) cudaFree(ctx->d_red);
  if (ctx->h_red) cudaFreeHost(ctx->h_red);
  ctx->d_red = nullptr;
  ctx->h_red = nullptr;
  ctx->red_doubles = 0;
  QP_CUDA(ctx, cudaMalloc(&ctx->d_red, doubles * sizeof(double)));
  QP_CUDA(ctx, cudaMallocHost(&ctx->h_red, doubles * sizeof(double)));
  ctx->red_doubles = doubles;
  return QP_OK;
}

int32_t qp_ctx_reserve_stage(qp_ctx_t ctx, size_t elems) {
  if (ctx->stage_elems >= elems) return QP_OK;
  if (ctx->h_stage) {
    // the previous staging buffer may still be the source of an in-flight copy
    QP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFreeHost(ctx->h_stage);
  }
  ctx->h_stage = nullptr;
  ctx->stage_elems = 0;
  QP_CUDA(ctx, cudaMallocHost(&ctx->h_stage, elems * sizeof(qp_c128)));
  ctx->stage_elems = elems;
  return QP_OK;
}

int32_t qp_ctx_reserve_part(qp_ctx_t ctx, size_t doubles) {
  if (ctx->part_doubles >= doubles) return QP_OK;
  QP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  cudaFree(ctx->d_part);
  ctx->d_part = nullptr;
  ctx->part_doubles = 0;
  QP_CUDA(ctx, cudaMalloc(&ctx->d_part, sizeof(double) * doubles));
  ctx->part_doubles = doubles;
  return QP_OK;
}

int32_t qp_ctx_reserve_gemv(qp_ctx_t ctx, size_t doubles) {
  if (ctx->gemv_doubles >= doubles) return QP_OK;
  QP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  cudaFree(ctx->d_gemv);
  ctx->d_gemv = nullptr;
  ctx->gemv_doubles = 0;
  QP_CUDA(ctx, cudaMalloc(&ctx->d_gemv, sizeof(double) * doubles));
  ctx->gemv_doubles = doubles;
  return QP_OK;
}

// one block per trajectory: every thread adds its slots (fixed stride), then a fixed-order tree
__global__ void __launch_bounds__(256)
k_part_reduce(const double* __restrict__ part, int64_t n_slots, int64_t batch, double* __restrict__ chk) {
  const int64_t b = blockIdx.x;
  double s[3] = {0.0, 0.0, 0.0};
  for (int64_t i = threadIdx.x; i < n_slots; i += 256)
    for (int k = 0; k < 3; ++k) s[k] += part[(i * batch + b) * 3 + k];
  __shared__ double sh[256][3];
  for (int k = 0; k < 3; ++k) sh[threadIdx.x][k] = s[k];
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o)
      for (int k = 0; k < 3; ++k) sh[threadIdx.x][k] += sh[threadIdx.x + o][k];
    __syncthreads();
  }
  if (threadIdx.x < 3) chk[3 * b + threadIdx.x] = sh[0][threadIdx.x];
}

int32_t qp_part_reduce(qp_ctx_t ctx, const double* part, int64_t n_slots, int64_t batch, double* chk) {
  k_part_reduce<<<(unsigned)batch, 256, 0, ctx->stream>>>(part, n_slots, batch, chk);
  QP_LAUNCHED(ctx);
  return QP_OK;
}

// ---------------------------------------------------------------------------------------
// timers
// ---------------------------------------------------------------------------------------

QpScopedTimer::QpScopedTimer(qp_ctx_t c, const char* l) : ctx(c), label(l), active(c->timers_on) {
  if (!active) return;
  // timers nest ("prop_step!" contains "matrix-vector product"): each scope owns its events
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) {
    active = false;
    return;
  }
  cudaEventRecord(e0, ctx->stream);
}

QpScopedTimer::~QpScopedTimer() {
  if (!active) return;
  cudaEventRecord(e1, ctx->stream);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  qp_timer_rec& r = ctx->timers[label];
  r.ncalls += 1;
  r.seconds += 1e-3 * ms;
}

extern "C" int32_t qp_timer_enable(qp_ctx_t ctx, int32_t on) {
  if (!ctx) return qp_fail(ctx, QP_ERR_INVALID_ARG, "null context");
  ctx->timers_on = on != 0;
  return QP_OK;
}

extern "C" int32_t qp_timer_get(qp_ctx_t ctx, const char* label, int64_t* ncalls, double* seconds) {
  if (!ctx || !label) return qp_fail(ctx, QP_ERR_INVALID_ARG, "qp_timer_get: null argument");
  auto it = ctx->timers.find(label);
  if (ncalls) *ncalls = it == ctx->timers.end() ? 0 : it->second.ncalls;
  if (seconds) *seconds = it == ctx->timers.end() ? 0.0 : it->second.seconds;
  return QP_OK;
}

extern "C" int32_t qp_timer_reset(qp_ctx_t ctx) {
  if (!ctx) return qp_fail(ctx, QP_ERR_INVALID_ARG, "null context");
  ctx->timers.clear();
  return QP_OK;
}

// ---------------------------------------------------------------------------------------
// states
// ---------------------------------------------------------------------------------------

extern "C" int32_t qp_state_create(qp_ctx_t ctx, int64_t n, int64_t batch, qp_state_t* out) {
  QP_CHECK(qp_ctx_bind(ctx));
  QP_REQUIRE(ctx, out != nullptr, "qp_state_create: null output pointer");
  QP_REQUIRE(ctx, n >= 1 && batch >= 1, "qp_state_create: n=%lld, batch=%lld must be >= 1",
             (long long)n, (long long)batch);
  QP_REQUIRE(ctx, n < (int64_t(1) << QP_COL_BITS), "qp_state_create: n=%lld exceeds 2^%d",
             (long long)n, QP_COL_BITS);
  qp_state_t st = new qp_state_s();
  st->ctx = ctx;
  st->n = n;
  st->batch = batch;
  cudaError_t e = cudaMalloc(&st->d, sizeof(double2) * (size_t)n * (size_t)batch);
  if (e != cudaSuccess) {
    cudaGetLastError();
    delete st;
    return qp_fail(ctx, QP_ERR_OOM, "qp_state_create: cudaMalloc of %lld x %lld state failed: %s",
                   (long long)n, (long long)batch, cudaGetErrorString(e));
  }
  *out = st;
  return QP_OK;
}

extern "C" int32_t qp_state_destroy(qp_state_t st) {
  if (!st) return QP_OK;
  cudaSetDevice(st->ctx->device);
  cudaStreamSynchronize(st->ctx->stream);
  cudaFree(st->d);
  delete st;
  return QP_OK;
}

extern "C" int32_t qp_state_info(qp_state_t st, int64_t* n, int64_t* batch) {
  if (!st) return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_state_info: null state");
  if (n) *n = st->n;
  if (batch) *batch = st->batch;
  return QP_OK;
}

extern "C" int32_t qp_state_devptr(qp_state_t st, void** devptr) {
  if (!st || !devptr) return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_state_devptr: null argument");
  *devptr = st->d;
  return QP_OK;
}

static int32_t state_xfer(qp_state_t st, qp_c128* host, int64_t b0, int64_t nb, bool upload, bool sync = true) {
  if (!st) return qp_fail(nullptr, QP_ERR_INVALID_ARG, "null state");
  qp_ctx_t ctx = st->ctx;
  QP_CHECK(qp_ctx_bind(ctx));
  QP_REQUIRE(ctx, host != nullptr, "state transfer: null host pointer");
  QP_REQUIRE(ctx, b0 >= 0 && nb >= 1 && b0 + nb <= st->batch,
             "state transfer: columns [%lld, %lld) outside batch %lld", (long long)b0,
             (long long)(b0 + nb), (long long)st->batch);
  // device row pitch = batch elements, host row pitch = nb elements
  const size_t el = sizeof(double2);
  if (nb == st->batch) {
    // full width: one contiguous transfer at link speed (a pitched copy with 16-byte rows is
    // an order of magnitude slower)
    const size_t bytes = el * (size_t)st->n * (size_t)st->batch;
    if (upload) QP_CUDA(ctx, cudaMemcpyAsync(st->d, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    else QP_CUDA(ctx, cudaMemcpyAsync(host, st->d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  } else if (upload) {
    QP_CUDA(ctx, cudaMemcpy2DAsync(st->d + b0, st->batch * el, host, nb * el, nb * el, st->n,
                                   cudaMemcpyHostToDevice, ctx->stream));
  } else {
    QP_CUDA(ctx, cudaMemcpy2DAsync(host, nb * el, st->d + b0, st->batch * el, nb * el, st->n,
                                   cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (sync) QP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return QP_OK;
}

extern "C" int32_t qp_state_upload_async(qp_state_t st, const qp_c128* host, int64_t b0, int64_t nb) {
  return state_xfer(st, const_cast<qp_c128*>(host), b0, nb, true, false);
}

extern "C" int32_t qp_state_download_async(qp_state_t st, qp_c128* host, int64_t b0, int64_t nb) {
  return state_xfer(st, host, b0, nb, false, false);
}

extern "C" int32_t qp_state_upload(qp_state_t st, const qp_c128* host, int64_t b0, int64_t nb) {
  return state_xfer(st, const_cast<qp_c128*>(host), b0, nb, true);
}

extern "C" int32_t qp_state_download(qp_state_t st, qp_c128* host, int64_t b0, int64_t nb) {
  return state_xfer(st, host, b0, nb, false);
}

// ---------------------------------------------------------------------------------------
// level-1 verbs
// ---------------------------------------------------------------------------------------

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

__global__ void k_fill(double2* __restrict__ x, double2 v, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x)
    x[i] = v;
}

__global__ void k_scal(double2* __restrict__ x, double2 a, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x)
    x[i] = cmul(a, x[i]);
}

__global__ void k_axpy(double2 a, const double2* __restrict__ x, double2* __restrict__ y,
                       int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    double2 t = cmul(a, x[i]);
    double2 yi = y[i];
    y[i] = make_double2(yi.x + t.x, yi.y + t.y);
  }
}

static inline int grid_for(qp_ctx_t ctx, int64_t total, int block) {
  int64_t want = (total + block - 1) / block;
  int64_t cap = (int64_t)ctx->sm_count * 8;
  return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

static int32_t same_shape(qp_state_t a, qp_state_t b, const char* what) {
  if (!a || !b) return qp_fail(a ? (qp_ctx_t)a->ctx : (qp_ctx_t) nullptr, QP_ERR_INVALID_ARG, "%s: null state", what);
  QP_REQUIRE(a->ctx, a->ctx == b->ctx, "%s: states belong to different contexts", what);
  QP_REQUIRE(a->ctx, a->n == b->n && a->batch == b->batch,
             "%s: shape mismatch (%lld x %lld) vs (%lld x %lld)", what, (long long)a->n,
             (long long)a->batch, (long long)b->n, (long long)b->batch);
  return QP_OK;
}

extern "C" int32_t qp_copy(qp_state_t dst, qp_state_t src) {
  QP_CHECK(same_shape(dst, src, "qp_copy"));
  qp_ctx_t ctx = dst->ctx;
  QP_CHECK(qp_ctx_bind(ctx));
  if (dst->d == src->d) return QP_OK;
  QP_CUDA(ctx, cudaMemcpyAsync(dst->d, src->d, sizeof(double2) * dst->n * dst->batch,
                               cudaMemcpyDeviceToDevice, ctx->stream));
  return QP_OK;
}

extern "C" int32_t qp_fill(qp_state_t st, qp_c128 value) {
  if (!st) return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_fill: null state");
  qp_ctx_t ctx = st->ctx;
  QP_CHECK(qp_ctx_bind(ctx));
  int64_t total = st->n * st->batch;
  k_fill<<<grid_for(ctx, total, 256), 256, 0, ctx->stream>>>(st->d, make_double2(value.re, value.im), total);
  QP_LAUNCHED(ctx);
  return QP_OK;
}

extern "C" int32_t qp_scal(qp_state_t st, qp_c128 alpha) {
  if (!st) return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_scal: null state");
  qp_ctx_t ctx = st->ctx;
  QP_CHECK(qp_ctx_bind(ctx));
  int64_t total = st->n * st->batch;
  k_scal<<<grid_for(ctx, total, 256), 256, 0, ctx->stream>>>(st->d, make_double2(alpha.re, alpha.im), total);
  QP_LAUNCHED(ctx);
  return QP_OK;
}

extern "C" int32_t qp_axpy(qp_c128 alpha, qp_state_t x, qp_state_t y) {
  QP_CHECK(same_shape(x, y, "qp_axpy"));
  qp_ctx_t ctx = x->ctx;
  QP_CHECK(qp_ctx_bind(ctx));
  int64_t total = x->n * x->batch;
  k_axpy<<<grid_for(ctx, total, 256), 256, 0, ctx->stream>>>(make_double2(alpha.re, alpha.im), x->d, y->d, total);
  QP_LAUNCHED(ctx);
  return QP_OK;
}

// ---------------------------------------------------------------------------------------
// deterministic reductions (two stages: per-block partials, then a fixed-order final sum)
//
// partial layout: [block][batch][2] doubles.  For batch == 1 threads stride over rows; for
// batch > 1 a 2-D block (32 batch lanes x 8 row lanes) keeps loads coalesced along the batch.
// ---------------------------------------------------------------------------------------

constexpr int RED_BLOCK = 256;

template <bool NORM>
__global__ void k_dot_partial_b1(const double2* __restrict__ x, const double2* __restrict__ y,
                                 int64_t n, double* __restrict__ partial) {
  double sr = 0.0, si = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    double2 a = x[i];
    if (NORM) {
      sr += a.x * a.x + a.y * a.y;
    } else {
      double2 b = y[i];
      sr += a.x * b.x + a.y * b.y;  // conj(a) * b
      si += a.x * b.y - a.y * b.x;
    }
  }
  __shared__ double shr[RED_BLOCK / 32], shi[RED_BLOCK / 32];
  for (int o = 16; o > 0; o >>= 1) {
    sr += __shfl_xor_sync(0xffffffffu, sr, o);
    si += __shfl_xor_sync(0xffffffffu, si, o);
  }
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) {
    shr[w] = sr;
    shi[w] = si;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tr = 0, ti = 0;
    for (int k = 0; k < RED_BLOCK / 32; ++k) {
      tr += shr[k];
      ti += shi[k];
    }
    partial[2 * blockIdx.x + 0] = tr;
    partial[2 * blockIdx.x + 1] = ti;
  }
}

// batch > 1: blockDim = (32, 8); blockIdx.y selects a chunk of 32 batch columns
template <bool NORM>
__global__ void k_dot_partial_bn(const double2* __restrict__ x, const double2* __restrict__ y,
                                 int64_t n, int64_t batch, double* __restrict__ partial) {
  int64_t b = blockIdx.y * 32 + threadIdx.x;
  double sr = 0.0, si = 0.0;
  if (b < batch) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.y + threadIdx.y; r < n;
         r += (int64_t)gridDim.x * blockDim.y) {
      double2 a = x[r * batch + b];
      if (NORM) {
        sr += a.x * a.x + a.y * a.y;
      } else {
        double2 c = y[r * batch + b];
        sr += a.x * c.x + a.y * c.y;
        si += a.x * c.y - a.y * c.x;
      }
    }
  }
  __shared__ double shr[8][33], shi[8][33];
  shr[threadIdx.y][threadIdx.x] = sr;
  shi[threadIdx.y][threadIdx.x] = si;
  __syncthreads();
  if (threadIdx.y == 0 && b < batch) {
    double tr = 0, ti = 0;
    for (int k = 0; k < 8; ++k) {
      tr += shr[k][threadIdx.x];
      ti += shi[k][threadIdx.x];
    }
    partial[2 * (blockIdx.x * batch + b) + 0] = tr;
    partial[2 * (blockIdx.x * batch + b) + 1] = ti;
  }
}

// final: one warp per batch column sums the block partials (fixed order: deterministic)
__global__ void k_dot_final(const double* __restrict__ partial, int nblocks, int64_t batch,
                            double* __restrict__ out) {
  const int64_t b = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= batch) return;
  double tr = 0, ti = 0;
  for (int k = lane; k < nblocks; k += 32) {
    tr += partial[2 * (k * batch + b) + 0];
    ti += partial[2 * (k * batch + b) + 1];
  }
  for (int o = 16; o > 0; o >>= 1) {
    tr += __shfl_xor_sync(0xffffffffu, tr, o);
    ti += __shfl_xor_sync(0xffffffffu, ti, o);
  }
  if (lane == 0) {
    out[2 * b + 0] = tr;
    out[2 * b + 1] = ti;
  }
}

template <bool NORM>
static int32_t reduce_impl(qp_ctx_t ctx, const double2* x, const double2* y, int64_t n,
                           int64_t batch, double* host_out /* 2*batch doubles */) {
  int nblocks;
  if (batch == 1) {
    nblocks = grid_for(ctx, n, RED_BLOCK);
    if (nblocks > ctx->sm_count * 4) nblocks = ctx->sm_count * 4;
  } else {
    int64_t want = (n + 7) / 8;
    int64_t cap = (int64_t)ctx->sm_count * 8 / ((batch + 31) / 32);
    if (cap < 1) cap = 1;
    nblocks = (int)(want < cap ? want : cap);
  }
  size_t need = (size_t)2 * batch * ((size_t)nblocks + 1);
  QP_CHECK(qp_ctx_reserve_red(ctx, need));
  double* partial = ctx->d_red + 2 * batch;  // first 2*batch doubles hold the result
  if (batch == 1) {
    k_dot_partial_b1<NORM><<<nblocks, RED_BLOCK, 0, ctx->stream>>>(x, y, n, partial);
  } else {
    dim3 grid(nblocks, (unsigned)((batch + 31) / 32)), block(32, 8);
    k_dot_partial_bn<NORM><<<grid, block, 0, ctx->stream>>>(x, y, n, batch, partial);
  }
  QP_LAUNCHED(ctx);
  k_dot_final<<<(unsigned)((batch * 32 + 127) / 128), 128, 0, ctx->stream>>>(partial, nblocks, batch, ctx->d_red);
  QP_LAUNCHED(ctx);
  QP_CUDA(ctx, cudaMemcpyAsync(ctx->h_red, ctx->d_red, sizeof(double) * 2 * batch,
                               cudaMemcpyDeviceToHost, ctx->stream));
  QP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(host_out, ctx->h_red, sizeof(double) * 2 * batch);
  return QP_OK;
}

int32_t qp_reduce_dot(qp_ctx_t ctx, const double2* x, const double2* y, int64_t n, int64_t batch,
                      qp_c128* out) {
  return reduce_impl<false>(ctx, x, y, n, batch, reinterpret_cast<double*>(out));
}

int32_t qp_reduce_norm2(qp_ctx_t ctx, const double2* x, int64_t n, int64_t batch, double* out) {
  std::vector<double> tmp(2 * batch);
  QP_CHECK(reduce_impl<true>(ctx, x, nullptr, n, batch, tmp.data()));
  for (int64_t b = 0; b < batch; ++b) out[b] = tmp[2 * b];
  return QP_OK;
}

extern "C" int32_t qp_dot(qp_state_t x, qp_state_t y, qp_c128* out) {
  QP_CHECK(same_shape(x, y, "qp_dot"));
  qp_ctx_t ctx = x->ctx;
  QP_CHECK(qp_ctx_bind(ctx));
  QP_REQUIRE(ctx, out != nullptr, "qp_dot: null output");
  return qp_reduce_dot(ctx, x->d, y->d, x->n, x->batch, out);
}

extern "C" int32_t qp_norm(qp_state_t x, double* out) {
  if (!x) return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_norm: null state");
  qp_ctx_t ctx = x->ctx;
  QP_CHECK(qp_ctx_bind(ctx));
  QP_REQUIRE(ctx, out != nullptr, "qp_norm: null output");
  QP_CHECK(qp_reduce_norm2(ctx, x->d, x->n, x->batch, out));
  for (int64_t b = 0; b < x->batch; ++b) out[b] = sqrt(out[b]);
  return QP_OK;
}
