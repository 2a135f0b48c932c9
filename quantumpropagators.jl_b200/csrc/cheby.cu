// Chebyshev propagation step: replaces ChebyWrk / cheby! (reference src/cheby.jl:87-213).
//
// One prop_step! is n_coeffs-1 launches of the fused kernel (spmv.cuh).  Three buffers
// rotate: the state's own buffer serves as v_0 (no copy, reference :171), w1 receives v_1,
// w2 accumulates psi; v_{k+1} overwrites v_{k-1}; at the end the state handle simply adopts
// the accumulator buffer (pointer swap instead of a copy).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "spmv.cuh"

extern "C" int32_t qp_cheby_create(qp_gen_t gen, qp_state_t like, qp_cheby_t* out) {
  if (!gen) return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_cheby_create: null generator");
  qp_ctx_t ctx = gen->ctx;
  QP_CHECK(qp_ctx_bind(ctx));
  QP_REQUIRE(ctx, out != nullptr && like != nullptr, "qp_cheby_create: null argument");
  *out = nullptr;
  QP_REQUIRE(ctx, like->ctx == ctx, "qp_cheby_create: state belongs to another context");
  QP_REQUIRE(ctx, like->n == gen->n, "qp_cheby_create: state dimension %lld != operator dimension %lld",
             (long long)like->n, (long long)gen->n);
  qp_cheby_t w = new qp_cheby_s();
  w->ctx = ctx;
  w->gen = gen;
  w->n = like->n;
  w->batch = like->batch;
  const size_t bytes = sizeof(double2) * (size_t)w->n * (size_t)w->batch;
  cudaError_t e;
  if ((e = cudaMalloc(&w->w1, bytes)) != cudaSuccess || (e = cudaMalloc(&w->w2, bytes)) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(w->w1);
    delete w;
    return qp_fail(ctx, QP_ERR_OOM, "qp_cheby_create: cudaMalloc of work vectors failed: %s", cudaGetErrorString(e));
  }
  *out = w;
  return QP_OK;
}

extern "C" int32_t qp_cheby_destroy(qp_cheby_t w) {
  if (!w) return QP_OK;
  cudaSetDevice(w->ctx->device);
  cudaStreamSynchronize(w->ctx->stream);
  cudaFree(w->w1);
  cudaFree(w->w2);
  cudaFree(w->d_chk);
  delete w;
  return QP_OK;
}

extern "C" int32_t qp_cheby_set_coeffs(qp_cheby_t w, const double* a, int32_t n_a, double Delta,
                                       double E_min, double dt_abs, double limit) {
  if (!w) return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_cheby_set_coeffs: null workspace");
  qp_ctx_t ctx = w->ctx;
  QP_REQUIRE(ctx, a != nullptr, "qp_cheby_set_coeffs: null coefficient array");
  // "Need at least 2 Chebychev coefficients" src/cheby.jl:165
  QP_REQUIRE(ctx, n_a > 1, "qp_cheby_set_coeffs: need at least 2 Chebychev coefficients (got %d)", n_a);
  QP_REQUIRE(ctx, Delta > 0.0, "qp_cheby_set_coeffs: spectral radius Delta=%g must be positive", Delta);
  QP_REQUIRE(ctx, dt_abs > 0.0, "qp_cheby_set_coeffs: dt=%g must be positive", dt_abs);
  w->a.assign(a, a + n_a);
  w->Delta = Delta;
  w->E_min = E_min;
  w->dt = dt_abs;
  w->limit = limit;  // threshold of the normalization check, src/cheby.jl:164,197
  return QP_OK;
}

extern "C" int32_t qp_cheby_step_bytes(qp_cheby_t w, int64_t* bytes) {
  if (!w || !bytes) return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_cheby_step_bytes: null argument");
  int64_t n_a = (int64_t)w->a.size();
  *bytes = n_a < 2 ? 0 : (n_a - 1) * (w->gen->matrix_bytes + 80 * w->n * w->batch);
  return QP_OK;
}

// The device part of one prop_step!: n_a - 1 fused launches with the operator coefficients
// already at gen->d_coef.  `chk` (or nullptr) receives the normalization-check sums.
static int32_t cheby_step_device(qp_cheby_t w, qp_state_t st, int stride, double dt_signed, double* chk) {
  qp_gen_t gen = w->gen;
  const int n_a = (int)w->a.size();
  const double beta = w->Delta / 2 + w->E_min;                  // :156
  const double2 c1 = make_double2(0.0, dt_signed > 0 ? -2.0 / w->Delta : 2.0 / w->Delta);  // :158-162
  const double2 c2 = make_double2(0.0, 2.0 * c1.y);             // :184
  const double ph = -beta * dt_signed;                          // exp(-i beta dt), :211
  const double2 phase = make_double2(cos(ph), sin(ph));

  double2* v0 = st->d;
  double2* acc = w->w2;
  EpiArgs e;
  memset(&e, 0, sizeof(e));
  {
    static const int hint = getenv("QPROP_VEC_HINT") ? atoi(getenv("QPROP_VEC_HINT")) : 0;
    e.vec_hint = hint;
  }
  e.beta = beta;
  e.phase = phase;
  e.acc = acc;
  if (n_a == 2) {
    e.c = c1;
    e.a0 = w->a[0];
    e.ak = w->a[1];
    QP_CHECK(qp_launch_fused(gen, EPI_CHEB_ONLY, stride, v0, w->batch, e));
  } else {
    e.c = c1;
    e.a0 = w->a[0];
    e.ak = w->a[1];
    e.y = w->w1;
    QP_CHECK(qp_launch_fused(gen, EPI_CHEB_FIRST, stride, v0, w->batch, e));
    double2* cur = w->w1;
    double2* prev = v0;
    e.c = c2;
    for (int k = 2; k < n_a; ++k) {
      e.ak = w->a[k];
      e.y = prev;
      e.chk = chk ? chk + (size_t)3 * w->batch * k : nullptr;
      const bool last = (k == n_a - 1);
      QP_CHECK(qp_launch_fused(gen, last ? EPI_CHEB_LAST : EPI_CHEB_MID, stride, cur, w->batch, e));
      double2* t = cur;
      cur = prev;
      prev = t;
    }
  }
  // the state adopts the accumulator; its old buffer becomes a work vector
  w->w2 = st->d;
  st->d = acc;
  return QP_OK;
}

static int32_t cheby_check_args(qp_cheby_t w, qp_state_t st, double dt_signed, const char* what) {
  qp_ctx_t ctx = w->ctx;
  QP_REQUIRE(ctx, st != nullptr && st->ctx == ctx, "%s: bad state", what);
  QP_REQUIRE(ctx, st->n == w->n && st->batch == w->batch,
             "%s: state shape (%lld x %lld) does not match the workspace (%lld x %lld)", what,
             (long long)st->n, (long long)st->batch, (long long)w->n, (long long)w->batch);
  QP_REQUIRE(ctx, (int)w->a.size() > 1, "%s: coefficients not set (qp_cheby_set_coeffs)", what);
  // @assert abs(dt) ≈ abs(wrk.dt)   src/cheby.jl:157  (isapprox: rtol = sqrt(eps))
  const double x = fabs(dt_signed), y = fabs(w->dt);
  QP_REQUIRE(ctx, fabs(x - y) <= 1.4901161193847656e-08 * fmax(x, y),
             "%s: wrk was initialized for dt=%.17g, not dt=abs(%.17g)", what, w->dt, dt_signed);
  return QP_OK;
}

extern "C" int32_t qp_cheby_step(qp_cheby_t w, qp_state_t st, const qp_c128* op_coeffs,
                                 int32_t coeffs_per_traj, double dt_signed, int32_t check_normalization) {
  if (!w) return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_cheby_step: null workspace");
  qp_ctx_t ctx = w->ctx;
  QP_CHECK(qp_ctx_bind(ctx));
  QP_CHECK(cheby_check_args(w, st, dt_signed, "qp_cheby_step"));
  const int n_a = (int)w->a.size();
  QpScopedTimer timer(ctx, "prop_step!");

  qp_gen_t gen = w->gen;
  int stride = 0;
  QP_CHECK(qp_gen_set_coeffs(gen, op_coeffs, coeffs_per_traj, w->batch, &stride));

  double* chk = nullptr;
  if (check_normalization && n_a > 2) {
    size_t need = (size_t)3 * w->batch * n_a;
    if (w->chk_doubles < need) {
      cudaFree(w->d_chk);
      w->d_chk = nullptr;
      w->chk_doubles = 0;
      QP_CUDA(ctx, cudaMalloc(&w->d_chk, sizeof(double) * need));
      w->chk_doubles = need;
    }
    QP_CUDA(ctx, cudaMemsetAsync(w->d_chk, 0, sizeof(double) * need, ctx->stream));
    chk = w->d_chk;
  }
  QP_CHECK(cheby_step_device(w, st, stride, dt_signed, chk));

  if (chk) {
    std::vector<double> h((size_t)3 * w->batch * n_a);
    QP_CUDA(ctx, cudaMemcpyAsync(h.data(), chk, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, ctx->stream));
    QP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int k = 2; k < n_a; ++k)
      for (int64_t b = 0; b < w->batch; ++b) {
        const double* p = &h[((size_t)k * w->batch + b) * 3];
        const double map_norm = hypot(p[0], p[1]) / (2.0 * p[2]);  // |<v1|v2'>| / (2 |v1|^2)
        if (!(map_norm <= 1.0 + w->limit))
          return qp_fail(ctx, QP_ERR_NORMALIZATION,
                         "Incorrect normalization (E_min=%.17g, Delta=%.17g): map norm %.17g in term %d, trajectory %lld",
                         w->E_min, w->Delta, map_norm, k + 1, (long long)b);
      }
  }
  return QP_OK;
}

// ---------------------------------------------------------------------------------------
// whole-grid propagation with on-device observables
// ---------------------------------------------------------------------------------------

// |x_b|^2 accumulated into out[b*3]
__global__ void k_norm2_acc_b1(const double2* __restrict__ x, int64_t n, double* __restrict__ out) {
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double2 v = x[i];
    s += v.x * v.x + v.y * v.y;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ double sh[32];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) atomicAdd(out, s);
  }
}

__global__ void k_norm2_acc_bn(const double2* __restrict__ x, int64_t n, int64_t batch, double* __restrict__ out) {
  // block (32, 8): x = trajectory within the chunk, y = row
  const int64_t b = (int64_t)blockIdx.y * 32 + threadIdx.x;
  double s = 0.0;
  if (b < batch)
    for (int64_t r = (int64_t)blockIdx.x * 8 + threadIdx.y; r < n; r += (int64_t)gridDim.x * 8) {
      const double2 v = x[r * batch + b];
      s += v.x * v.x + v.y * v.y;
    }
  __shared__ double sh[8][33];
  sh[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && b < batch) {
    for (int j = 1; j < 8; ++j) s += sh[j][threadIdx.x];
    atomicAdd(out + 3 * b, s);
  }
}

extern "C" int32_t qp_cheby_propagate(qp_cheby_t w, qp_state_t st, const qp_c128* coeff_table,
                                      int32_t coeffs_per_traj, int32_t n_steps, double dt_signed,
                                      int32_t n_obs, const qp_gen_t* obs, qp_c128* expvals, double* norms) {
  if (!w) return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_cheby_propagate: null workspace");
  qp_ctx_t ctx = w->ctx;
  QP_CHECK(qp_ctx_bind(ctx));
  QP_CHECK(cheby_check_args(w, st, dt_signed, "qp_cheby_propagate"));
  QP_REQUIRE(ctx, n_steps >= 0, "qp_cheby_propagate: n_steps must be >= 0");
  QP_REQUIRE(ctx, n_obs >= 0 && (n_obs == 0 || (obs != nullptr && expvals != nullptr)),
             "qp_cheby_propagate: observables given without an output array");
  qp_gen_t gen = w->gen;
  QP_REQUIRE(ctx, gen->n_coeffs == 0 || n_steps == 0 || coeff_table != nullptr,
             "qp_cheby_propagate: null coefficient table");
  const int64_t B = w->batch;
  for (int k = 0; k < n_obs; ++k) {
    QP_REQUIRE(ctx, obs[k] != nullptr && obs[k]->ctx == ctx && obs[k]->n == w->n,
               "qp_cheby_propagate: observable %d does not match the state", k);
    QP_REQUIRE(ctx, obs[k]->n_coeffs == 0, "qp_cheby_propagate: observable %d must have no free coefficients", k);
    QP_REQUIRE(ctx, obs[k] != gen, "qp_cheby_propagate: the generator itself cannot be an observable (wrap its operators in a second generator)");
  }
  QpScopedTimer timer(ctx, "propagate");

  // effective per-operator coefficients of every interval, uploaded once:
  // [n_steps][n_ops][width], drift operators = 1 (src/generators.jl:634-636)
  const int64_t width = coeffs_per_traj ? B : 1;
  const size_t per_step = (size_t)gen->n_ops * (size_t)width;
  double2* d_tbl = nullptr;
  double* d_rec = nullptr;
  const int n_slots = n_obs + 1;  // expectation values + |psi|^2
  const size_t rec_doubles = (norms || n_obs > 0) ? (size_t)(n_steps + 1) * n_slots * B * 3 : 0;
  auto cleanup = [&](int32_t rc) {
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_tbl);
    cudaFree(d_rec);
    return rc;
  };
  std::vector<double2> h_tbl(per_step * (size_t)n_steps);
  if (n_steps > 0) {
    for (int s = 0; s < n_steps; ++s)
      for (int l = 0; l < gen->n_ops; ++l)
        for (int64_t b = 0; b < width; ++b) {
          double2& dst = h_tbl[((size_t)s * gen->n_ops + l) * width + b];
          if (l < gen->drift) dst = make_double2(1.0, 0.0);
          else {
            const qp_c128 c = coeff_table[((size_t)s * gen->n_coeffs + (l - gen->drift)) * width + b];
            dst = make_double2(c.re, c.im);
          }
        }
    QP_CUDA(ctx, cudaMalloc(&d_tbl, sizeof(double2) * h_tbl.size()));
    cudaError_t e1 = cudaMemcpy(d_tbl, h_tbl.data(), sizeof(double2) * h_tbl.size(), cudaMemcpyHostToDevice);
    if (e1 != cudaSuccess) return cleanup(qp_fail(ctx, QP_ERR_CUDA, "qp_cheby_propagate: table upload failed: %s", cudaGetErrorString(e1)));
  }
  if (rec_doubles) {
    cudaError_t e1 = cudaMalloc(&d_rec, sizeof(double) * rec_doubles);
    if (e1 == cudaSuccess) e1 = cudaMemsetAsync(d_rec, 0, sizeof(double) * rec_doubles, ctx->stream);
    if (e1 != cudaSuccess) return cleanup(qp_fail(ctx, QP_ERR_OOM, "qp_cheby_propagate: record buffer: %s", cudaGetErrorString(e1)));
  }
  // observables are sums of fixed operators: their coefficient vectors are all ones
  for (int k = 0; k < n_obs; ++k) {
    int32_t rc = qp_gen_set_coeffs(obs[k], nullptr, 0, B, nullptr);
    if (rc != QP_OK) return cleanup(rc);
  }
  auto record = [&](int slot_step) -> int32_t {
    if (!rec_doubles) return QP_OK;
    double* base = d_rec + (size_t)slot_step * n_slots * B * 3;
    for (int k = 0; k < n_obs; ++k) {
      EpiArgs e;
      memset(&e, 0, sizeof(e));
      e.chk = base + (size_t)k * B * 3;
      QP_CHECK(qp_launch_fused(obs[k], EPI_DOT, 0, st->d, B, e));
    }
    if (norms) {
      double* out = base + (size_t)n_obs * B * 3;
      if (B == 1) {
        const int64_t blocks = std::min<int64_t>((w->n + 255) / 256, (int64_t)ctx->sm_count * 4);
        k_norm2_acc_b1<<<(unsigned)blocks, 256, 0, ctx->stream>>>(st->d, w->n, out);
      } else {
        dim3 grid((unsigned)std::min<int64_t>((w->n + 7) / 8, (int64_t)ctx->sm_count * 8), (unsigned)((B + 31) / 32)), block(32, 8);
        k_norm2_acc_bn<<<grid, block, 0, ctx->stream>>>(st->d, w->n, B, out);
      }
      QP_LAUNCHED(ctx);
    }
    return QP_OK;
  };

  double2* saved_coef = gen->d_coef;
  const std::vector<double2> saved_h_coef = gen->h_coef;
  const bool saved_h_valid = gen->h_coef_valid;
  int32_t rc = record(0);
  for (int s = 0; s < n_steps && rc == QP_OK; ++s) {
    gen->d_coef = d_tbl + (size_t)s * per_step;
    gen->h_coef_valid = width == 1;  // host copy: selects the real-table kernel variant per step
    if (width == 1) gen->h_coef.assign(h_tbl.begin() + (size_t)s * per_step, h_tbl.begin() + (size_t)(s + 1) * per_step);
    {
      QpScopedTimer step_timer(ctx, "prop_step!");  // same label and count as the step loop
      rc = cheby_step_device(w, st, coeffs_per_traj ? 1 : 0, dt_signed, nullptr);
    }
    if (rc == QP_OK) rc = record(s + 1);
  }
  gen->d_coef = saved_coef;
  gen->h_coef = saved_h_coef;
  gen->h_coef_valid = saved_h_valid;
  if (rc != QP_OK) return cleanup(rc);

  if (rec_doubles) {
    std::vector<double> h(rec_doubles);
    cudaError_t e1 = cudaMemcpyAsync(h.data(), d_rec, sizeof(double) * rec_doubles, cudaMemcpyDeviceToHost, ctx->stream);
    if (e1 == cudaSuccess) e1 = cudaStreamSynchronize(ctx->stream);
    if (e1 != cudaSuccess) return cleanup(qp_fail(ctx, QP_ERR_CUDA, "qp_cheby_propagate: download failed: %s", cudaGetErrorString(e1)));
    for (int s = 0; s <= n_steps; ++s) {
      const double* base = h.data() + (size_t)s * n_slots * B * 3;
      for (int k = 0; k < n_obs; ++k)
        for (int64_t b = 0; b < B; ++b)
          expvals[((size_t)s * n_obs + k) * B + b] = qp_c128{base[((size_t)k * B + b) * 3], base[((size_t)k * B + b) * 3 + 1]};
      if (norms)
        for (int64_t b = 0; b < B; ++b) norms[(size_t)s * B + b] = sqrt(base[((size_t)n_obs * B + b) * 3]);
    }
  }
  return cleanup(QP_OK);
}
