// Chebyshev propagation step: replaces ChebyWrk / cheby! (reference src/cheby.jl:87-213).
//
// One prop_step! is n_coeffs-1 launches of the fused kernel (spmv.cuh).  Three buffers
// rotate: the state's own buffer serves as v_0 (no copy, reference :171), w1 receives v_1,
// w2 accumulates psi; v_{k+1} overwrites v_{k-1}; at the end the state handle simply adopts
// the accumulator buffer (pointer swap instead of a copy).
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

extern "C" int32_t qp_cheby_step(qp_cheby_t w, qp_state_t st, const qp_c128* op_coeffs,
                                 int32_t coeffs_per_traj, double dt_signed, int32_t check_normalization) {
  if (!w) return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_cheby_step: null workspace");
  qp_ctx_t ctx = w->ctx;
  QP_CHECK(qp_ctx_bind(ctx));
  QP_REQUIRE(ctx, st != nullptr && st->ctx == ctx, "qp_cheby_step: bad state");
  QP_REQUIRE(ctx, st->n == w->n && st->batch == w->batch,
             "qp_cheby_step: state shape (%lld x %lld) does not match the workspace (%lld x %lld)",
             (long long)st->n, (long long)st->batch, (long long)w->n, (long long)w->batch);
  const int n_a = (int)w->a.size();
  QP_REQUIRE(ctx, n_a > 1, "qp_cheby_step: coefficients not set (qp_cheby_set_coeffs)");
  // @assert abs(dt) ≈ abs(wrk.dt)   src/cheby.jl:157  (isapprox: rtol = sqrt(eps))
  {
    const double x = fabs(dt_signed), y = fabs(w->dt);
    QP_REQUIRE(ctx, fabs(x - y) <= 1.4901161193847656e-08 * fmax(x, y),
               "qp_cheby_step: wrk was initialized for dt=%.17g, not dt=abs(%.17g)", w->dt, dt_signed);
  }
  QpScopedTimer timer(ctx, "prop_step!");

  qp_gen_t gen = w->gen;
  int stride = 0;
  QP_CHECK(qp_gen_set_coeffs(gen, op_coeffs, coeffs_per_traj, w->batch, &stride));

  const double beta = w->Delta / 2 + w->E_min;                  // :156
  const double2 c1 = make_double2(0.0, dt_signed > 0 ? -2.0 / w->Delta : 2.0 / w->Delta);  // :158-162
  const double2 c2 = make_double2(0.0, 2.0 * c1.y);             // :184
  const double ph = -beta * dt_signed;                          // exp(-i beta dt), :211
  const double2 phase = make_double2(cos(ph), sin(ph));

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
