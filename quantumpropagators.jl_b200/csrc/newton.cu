// newton!: the restarted Newton-polynomial propagation of reference src/newton.jl:246-385 as ONE
// library call (the optional all-in-one entry point).  The vector work runs on
// the device (qp_arnoldi, qp_krylov_combine); the small dense step -- Ritz values of the
// leading Hessenberg blocks (src/arnoldi.jl:143-170), Leja ordering (src/newton.jl:97-148),
// divided differences (:176-214) and the polynomial in the extended Hessenberg matrix
// (:330-343, 356-367) -- runs on the host in C++ instead of bouncing through the host
// language once per restart.  `func` is exp(-i z) (TDSE), exp(z), or a C callback that is
// evaluated at the Leja points only.  The fine-grained entry points stay the primary ABI for
// hosts that want to keep this step themselves (julia/QPropB200.jl does, reusing the
// reference's own functions).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstring>

#include "qprop_internal.h"

typedef std::complex<double> cplx;

// ---------------------------------------------------------------------------------------
// eigenvalues of a complex upper-Hessenberg matrix: shifted QR iteration with Givens
// rotations, Wilkinson shift, deflation; n <= ~60, row-major copy is destroyed
// ---------------------------------------------------------------------------------------
static bool hessenberg_eigvals(std::vector<cplx>& H, int n, std::vector<cplx>& w) {
  w.assign(n, cplx(0.0, 0.0));
  auto at = [&](int i, int j) -> cplx& { return H[(size_t)i * n + j]; };
  const double eps = 2.220446049250313e-16;
  double anorm = 0.0;
  for (int i = 0; i < n; ++i)
    for (int j = std::max(0, i - 1); j < n; ++j) anorm = std::max(anorm, std::abs(at(i, j)));
  if (anorm == 0.0) return true;
  std::vector<double> cs(n);
  std::vector<cplx> sn(n);
  int hi = n - 1, iter = 0;
  while (hi >= 0) {
    int l = hi;
    while (l > 0) {
      double s = std::abs(at(l - 1, l - 1)) + std::abs(at(l, l));
      if (s == 0.0) s = anorm;
      if (std::abs(at(l, l - 1)) <= eps * s) {
        at(l, l - 1) = 0.0;
        break;
      }
      --l;
    }
    if (l == hi) {  // one eigenvalue has converged
      w[hi] = at(hi, hi);
      --hi;
      iter = 0;
      continue;
    }
    if (++iter > 60 * n) return false;
    // Wilkinson shift: the eigenvalue of the trailing 2x2 block closer to its last entry
    cplx mu;
    if (iter % 11 == 10) {  // exceptional shift against stagnation
      mu = at(hi, hi) + std::abs(at(hi, hi - 1)) + (hi >= 2 ? std::abs(at(hi - 1, hi - 2)) : 0.0);
    } else {
      const cplx a = at(hi - 1, hi - 1), b = at(hi - 1, hi), c = at(hi, hi - 1), d = at(hi, hi);
      const cplx tr2 = 0.5 * (a + d), disc = std::sqrt(tr2 * tr2 - (a * d - b * c));
      const cplx m1 = tr2 + disc, m2 = tr2 - disc;
      mu = std::abs(m1 - d) < std::abs(m2 - d) ? m1 : m2;
    }
    for (int i = l; i <= hi; ++i) at(i, i) -= mu;
    for (int k = l; k < hi; ++k) {  // QR: rotations from the left
      const cplx x = at(k, k), y = at(k + 1, k);
      const double ax = std::abs(x), r = std::hypot(ax, std::abs(y));
      double c;
      cplx s;
      if (r == 0.0) {
        c = 1.0;
        s = 0.0;
      } else if (ax == 0.0) {
        c = 0.0;
        s = std::conj(y) / r;
      } else {
        c = ax / r;
        s = (x / ax) * std::conj(y) / r;
      }
      cs[k] = c;
      sn[k] = s;
      for (int j = k; j <= hi; ++j) {
        const cplx t1 = at(k, j), t2 = at(k + 1, j);
        at(k, j) = c * t1 + s * t2;
        at(k + 1, j) = -std::conj(s) * t1 + c * t2;
      }
    }
    for (int k = l; k < hi; ++k) {  // RQ: the same rotations from the right
      const double c = cs[k];
      const cplx s = sn[k];
      const int imax = std::min(k + 2, hi);
      for (int i = l; i <= imax; ++i) {
        const cplx t1 = at(i, k), t2 = at(i, k + 1);
        at(i, k) = c * t1 + std::conj(s) * t2;
        at(i, k + 1) = -s * t1 + c * t2;
      }
    }
    for (int i = l; i <= hi; ++i) at(i, i) += mu;
  }
  return true;
}

// Ritz values of all leading blocks 1..m of the column-major Hessenberg matrix, concatenated
// (diagonalize_hessenberg_matrix(...; accumulate=true), src/arnoldi.jl:143-170); every block's
// eigenvalues sorted by (real, imag) like Julia's eigvals
static bool ritz_accumulated(const std::vector<cplx>& hess, int ld, int m, std::vector<cplx>& out) {
  out.clear();
  auto H = [&](int i, int j) { return hess[(size_t)j * ld + i]; };
  for (int j = 1; j <= m; ++j) {
    if (j == 1) {
      out.push_back(H(0, 0));
    } else if (j == 2) {  // closed form, :156-163
      const cplx a = H(0, 0), b = H(0, 1), c = H(1, 0), d = H(1, 1);
      const cplx s = std::sqrt(a * a + 4.0 * b * c - 2.0 * a * d + d * d);
      out.push_back(0.5 * (a + d - s));
      out.push_back(0.5 * (a + d + s));
    } else {
      std::vector<cplx> A((size_t)j * j), w;
      for (int r = 0; r < j; ++r)
        for (int c = 0; c < j; ++c) A[(size_t)r * j + c] = (r <= c + 1) ? H(r, c) : cplx(0.0, 0.0);
      if (!hessenberg_eigvals(A, j, w)) return false;
      std::sort(w.begin(), w.end(), [](const cplx& p, const cplx& q) {
        return p.real() != q.real() ? p.real() < q.real() : p.imag() < q.imag();
      });
      out.insert(out.end(), w.begin(), w.end());
    }
  }
  return true;
}

// extend_leja!(leja, n, newpoints, n_use), src/newton.jl:97-148 (newpoints is clobbered)
static void extend_leja(std::vector<cplx>& leja, int& n, std::vector<cplx>& cand, int n_use) {
  if ((int)leja.size() < n + n_use) leja.resize(2 * (size_t)(n + n_use), cplx(0.0, 0.0));
  const int u = (int)cand.size() - 1;
  int start = 0;
  if (n == 0) {  // the candidate of largest magnitude starts the sequence (same pairwise swaps)
    cplx z_last = cand[u];
    for (int i = 0; i < u; ++i)
      if (std::abs(cand[i]) > std::abs(z_last)) {
        cand[u] = cand[i];
        cand[i] = z_last;
        z_last = cand[u];
      }
    leja[0] = cand[u];
    start = 1;
  }
  const double exponent = 1.0 / (n + n_use);
  for (int i_add = start; i_add < n_use; ++i_add) {
    double p_max = 0.0;
    int i_max = 0;
    for (int i = 0; i <= u - i_add; ++i) {
      double p = 1.0;
      for (int j = 0; j < n + i_add; ++j) p *= std::pow(std::abs(cand[i] - leja[j]), exponent);
      if (p > p_max) {
        p_max = p;
        i_max = i;
      }
    }
    leja[n + i_add] = cand[i_max];
    cand[i_max] = cand[u - i_add];
  }
  n += n_use;
}

struct NewtonFunc {
  int id;
  qp_newton_func_t cb;
  void* user;
  cplx operator()(cplx z) const {
    if (id == QP_FUNC_EXPMI) return std::exp(cplx(0.0, -1.0) * z);
    if (id == QP_FUNC_EXP) return std::exp(z);
    qp_c128 in{z.real(), z.imag()}, out{0.0, 0.0};
    cb(&in, &out, user);
    return cplx(out.re, out.im);
  }
};

// extend_newton_coeffs!(a, n_a, leja, func, n_leja, radius), src/newton.jl:176-214
static bool extend_newton_coeffs(std::vector<cplx>& a, int& n_a, const std::vector<cplx>& leja, const NewtonFunc& func,
                                 int n_leja, double radius) {
  if ((int)a.size() < n_leja) a.resize(2 * (size_t)n_leja, cplx(0.0, 0.0));
  int k0 = n_a;
  if (n_a == 0) {
    a[0] = func(leja[0]);
    k0 = 1;
  }
  for (int k = k0; k < n_leja; ++k) {
    cplx d(1.0, 0.0), pn(0.0, 0.0);
    for (int n = 1; n < k; ++n) {
      d = d * (leja[k] - leja[n - 1]) / radius;
      pn += a[n] * d;
    }
    d = d * (leja[k] - leja[k - 1]) / radius;
    if (!(std::abs(d) > 1e-200)) return false;  // "Divided differences too small"
    a[k] = (func(leja[k]) - a[0] - pn) / d;
  }
  n_a = n_leja;
  return true;
}

// host-side sections of newton! under the reference's TimerOutputs labels (src/newton.jl:297-343;
// test/test_timings.jl:8-39 is the contract): wall clock, they run between two device phases
struct HostTimer {
  qp_ctx_t ctx;
  const char* label;
  std::chrono::steady_clock::time_point t0;
  HostTimer(qp_ctx_t c, const char* l) : ctx(c), label(l), t0(std::chrono::steady_clock::now()) {}
  ~HostTimer() {
    if (!ctx->timers_on) return;
    qp_timer_rec& r = ctx->timers[label];
    r.ncalls += 1;
    r.seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
};

// k_scal_pb / per-trajectory scaling lives in krylov.cu
int32_t qp_krylov_scale_pb(qp_krylov_t K, qp_state_t st, const qp_c128* alpha);

// newton! for a bundle of B states sharing one generator: the Arnoldi process and
// the vector updates are batched kernels, every state keeps its own Hessenberg matrix, Leja points
// and Newton coefficients, and the restart loop runs in lock step until the LAST state has converged
// (a converged state contributes zero weights from then on and keeps its restart vector).
static int32_t newton_step_batched(qp_krylov_t K, qp_state_t psi, qp_state_t v, const qp_c128* op_coeffs, double dt,
                                   const NewtonFunc& func, double norm_min, double relerr, int32_t max_restarts,
                                   int32_t* restarts_out) {
  qp_ctx_t ctx = K->ctx;
  const int64_t B = K->batch;
  const int ld = K->m_max + 1;
  struct Traj {
    std::vector<cplx> a, leja, ritz, R, R2, P;
    int n_a = 0, n_leja = 0, m = 0, restarts = 0;
    double radius = 0.0, beta = 0.0;
    bool done = false;
  };
  std::vector<Traj> T((size_t)B);
  std::vector<qp_c128> hess((size_t)ld * ld * B), wP((size_t)(K->m_max + 1) * B), wR((size_t)(K->m_max + 1) * B), sc((size_t)B);
  std::vector<int32_t> m_out((size_t)B);
  std::vector<double> nrm((size_t)B);

  QP_CHECK(qp_copy(v, psi));  // v <- Psi (:268), normalised per state
  QP_CHECK(qp_norm(v, nrm.data()));
  for (int64_t b = 0; b < B; ++b) {
    QP_REQUIRE(ctx, nrm[b] > 0.0, "qp_newton_step: state %lld of the bundle has zero norm", (long long)b);
    T[b].beta = nrm[b];
    T[b].R.resize(ld);
    T[b].R2.resize(ld);
    T[b].P.resize(ld);
    sc[b] = qp_c128{1.0 / nrm[b], 0.0};
  }
  QP_CHECK(qp_krylov_scale_pb(K, v, sc.data()));

  int m = K->m_max, s = 0;
  for (;;) {
    QP_CHECK(qp_arnoldi(K, op_coeffs, v, m, dt, 1, norm_min, hess.data(), ld, m_out.data()));
    bool need_scale = false;
    for (int64_t b = 0; b < B; ++b) sc[b] = qp_c128{1.0, 0.0};
    std::fill(wP.begin(), wP.end(), qp_c128{0.0, 0.0});
    std::fill(wR.begin(), wR.end(), qp_c128{0.0, 0.0});
    for (int64_t b = 0; b < B; ++b) {
      Traj& t = T[b];
      wR[(size_t)0 * B + b] = qp_c128{1.0, 0.0};  // default: the restart vector stays q_0 = v
      if (t.done) continue;
      const cplx* Hb = reinterpret_cast<const cplx*>(hess.data()) + (size_t)b * ld * ld;
      const int mb = m_out[b];
      t.m = mb;
      if (mb == 1 && s == 0) {  // v is an eigenvector: f(H dt) Psi = f(lambda) Psi   (:289-295)
        const cplx f = func(t.beta * Hb[0]);
        sc[b] = qp_c128{f.real(), f.imag()};
        need_scale = true;
        t.done = true;
        continue;
      }
      std::vector<cplx> hb(Hb, Hb + (size_t)ld * ld);
      {
        HostTimer tm(ctx, "diagonalize_hessenberg_matrix");
        if (!ritz_accumulated(hb, ld, mb, t.ritz))
          return qp_fail(ctx, QP_ERR_INTERNAL, "qp_newton_step: QR iteration for the Ritz values did not converge (state %lld)", (long long)b);
      }
      if (t.n_leja == 0) {
        double mx = 0.0;
        for (const cplx& z : t.ritz) mx = std::max(mx, std::abs(z));
        t.radius = 1.2 * mx;
      }
      QP_REQUIRE(ctx, t.radius > 0.0, "qp_newton_step: Leja radius must be positive");
      const int n_s = t.n_leja;
      {
        HostTimer tm(ctx, "get Leja points");
        extend_leja(t.leja, t.n_leja, t.ritz, mb);
      }
      {
        HostTimer tm(ctx, "get Newton coeffs");
        if (!extend_newton_coeffs(t.a, t.n_a, t.leja, func, t.n_leja, t.radius))
          return qp_fail(ctx, QP_ERR_INTERNAL, "qp_newton_step: Divided differences too small");
      }
      auto Hm = [&](int i, int j) { return hb[(size_t)j * ld + i]; };
      auto step_R = [&](cplx shift) {
        for (int i = 0; i <= mb; ++i) {
          cplx acc(0.0, 0.0);
          for (int j = 0; j <= mb; ++j) acc += Hm(i, j) * t.R[j];
          t.R2[i] = (acc - shift * t.R[i]) / t.radius;
        }
        std::swap(t.R, t.R2);
      };
      {
        HostTimer tm(ctx, "evaluate polynomial");
        std::fill(t.R.begin(), t.R.end(), cplx(0.0, 0.0));
        t.R[0] = t.beta;
        for (int i = 0; i <= mb; ++i) t.P[i] = t.a[n_s] * t.R[i];
        for (int k = 1; k < mb; ++k) {
          step_R(t.leja[n_s + k - 1]);
          for (int i = 0; i <= mb; ++i) t.P[i] += t.a[n_s + k] * t.R[i];
        }
      }
      for (int i = 0; i < mb; ++i) wP[(size_t)i * B + b] = qp_c128{t.P[i].real(), t.P[i].imag()};
      step_R(t.leja[n_s + mb - 1]);
      double b2 = 0.0;
      for (int i = 0; i <= mb; ++i) b2 += std::norm(t.R[i]);
      t.beta = std::sqrt(b2);
      if (t.beta > 0.0)
        for (int i = 0; i <= mb; ++i) wR[(size_t)i * B + b] = qp_c128{t.R[i].real() / t.beta, t.R[i].imag() / t.beta};
    }
    if (need_scale) QP_CHECK(qp_krylov_scale_pb(K, psi, sc.data()));
    // Psi (+)= sum_i P_i q_i: the first restart overwrites Psi (fill!(Psi, 0), :346) -- except for states
    // that took the eigenvector shortcut, whose weights are zero and which must keep their (scaled) Psi
    if (s == 0) {
      bool any_shortcut = false;
      for (int64_t b = 0; b < B; ++b) any_shortcut |= T[b].done;
      if (any_shortcut) {  // zero the other states only
        for (int64_t b = 0; b < B; ++b) sc[b] = T[b].done ? qp_c128{1.0, 0.0} : qp_c128{0.0, 0.0};
        QP_CHECK(qp_krylov_scale_pb(K, psi, sc.data()));
        QP_CHECK(qp_krylov_combine(K, wP.data(), 0, m, psi, 1));
      } else {
        QP_CHECK(qp_krylov_combine(K, wP.data(), 0, m, psi, 0));
      }
    } else {
      QP_CHECK(qp_krylov_combine(K, wP.data(), 0, m, psi, 1));
    }
    QP_CHECK(qp_krylov_combine(K, wR.data(), 0, m + 1, v, 0));
    QP_CHECK(qp_norm(psi, nrm.data()));
    bool all_done = true;
    for (int64_t b = 0; b < B; ++b) {
      Traj& t = T[b];
      if (t.done) continue;
      if (t.beta * std::abs(t.a[t.n_a - 1]) / (1.0 + nrm[b]) < relerr) {
        t.done = true;
        t.restarts = s;
      } else {
        all_done = false;
      }
    }
    if (all_done) break;
    ++s;
    if (s > max_restarts)
      return qp_fail(ctx, QP_ERR_NOT_CONVERGED, "newton!: no convergence within max_restarts=%d", max_restarts);
  }
  if (restarts_out) *restarts_out = s;
  K->last_n_a = T[0].n_a;
  K->last_n_leja = T[0].n_leja;
  K->last_radius = T[0].radius;
  K->last_a.assign(reinterpret_cast<const qp_c128*>(T[0].a.data()), reinterpret_cast<const qp_c128*>(T[0].a.data()) + T[0].n_a);
  K->last_leja.assign(reinterpret_cast<const qp_c128*>(T[0].leja.data()), reinterpret_cast<const qp_c128*>(T[0].leja.data()) + T[0].n_leja);
  return QP_OK;
}

extern "C" int32_t qp_newton_step(qp_krylov_t K, qp_state_t psi, qp_state_t v, const qp_c128* op_coeffs, double dt,
                                  int32_t func_id, qp_newton_func_t func_cb, void* user, double norm_min,
                                  double relerr, int32_t max_restarts, int32_t* restarts_out) {
  if (!K) return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_newton_step: null workspace");
  qp_ctx_t ctx = K->ctx;
  QP_CHECK(qp_ctx_bind(ctx));
  QP_REQUIRE(ctx, psi && v && psi->ctx == ctx && v->ctx == ctx, "qp_newton_step: bad state");
  QP_REQUIRE(ctx, psi->n == K->n && v->n == K->n && psi->batch == K->batch && v->batch == K->batch,
             "qp_newton_step: states must have the shape of the workspace (%lld x %lld)", (long long)K->n, (long long)K->batch);
  QP_REQUIRE(ctx, psi->d != v->d, "qp_newton_step: psi and the work vector must not alias");
  QP_REQUIRE(ctx, dt != 0.0, "qp_newton_step: dt must be non-zero");
  QP_REQUIRE(ctx, func_id == QP_FUNC_EXPMI || func_id == QP_FUNC_EXP || (func_id == QP_FUNC_CALLBACK && func_cb != nullptr),
             "qp_newton_step: bad func_id %d", func_id);
  // NewtonWrk: m_max > 2 (src/newton.jl:40-46)
  QP_REQUIRE(ctx, K->m_max > 2, "qp_newton_step: Newton propagation requires m_max > 2 (got %d)", K->m_max);
  const NewtonFunc func{func_id, func_cb, user};
  if (K->batch > 1) return newton_step_batched(K, psi, v, op_coeffs, dt, func, norm_min, relerr, max_restarts, restarts_out);
  const int ld = K->m_max + 1;
  int m = K->m_max;
  std::vector<cplx> hess((size_t)ld * ld), a, leja, ritz, R(ld), R2(ld), P(ld);
  int n_a = 0, n_leja = 0, s = 0;
  double radius = 0.0;

  QP_CHECK(qp_copy(v, psi));  // v <- Psi (:268)
  double beta = 0.0;
  QP_CHECK(qp_norm(v, &beta));
  QP_REQUIRE(ctx, beta > 0.0, "qp_newton_step: the state has zero norm");
  QP_CHECK(qp_scal(v, qp_c128{1.0 / beta, 0.0}));

  for (;;) {
    int32_t m_out = 0;
    QP_CHECK(qp_arnoldi(K, op_coeffs, v, m, dt, 1, norm_min, reinterpret_cast<qp_c128*>(hess.data()), ld, &m_out));
    m = m_out;
    if (m == 1 && s == 0) {  // v is an eigenvector: f(H dt) Psi = f(lambda) Psi   (:289-295)
      const cplx f = func(beta * hess[0]);
      QP_CHECK(qp_scal(psi, qp_c128{f.real(), f.imag()}));
      break;
    }
    {
      HostTimer t(ctx, "diagonalize_hessenberg_matrix");  // :297-299
      if (!ritz_accumulated(hess, ld, m, ritz))
        return qp_fail(ctx, QP_ERR_INTERNAL, "qp_newton_step: QR iteration for the Ritz values did not converge");
    }
    if (s == 0) {  // leja_radius, :67-70
      double mx = 0.0;
      for (const cplx& z : ritz) mx = std::max(mx, std::abs(z));
      radius = 1.2 * mx;
    }
    QP_REQUIRE(ctx, radius > 0.0, "qp_newton_step: Leja radius must be positive");
    const int n_s = n_leja;
    {
      HostTimer t(ctx, "get Leja points");  // :307-311
      extend_leja(leja, n_leja, ritz, m);
    }
    {
      HostTimer t(ctx, "get Newton coeffs");  // :314-317
      if (!extend_newton_coeffs(a, n_a, leja, func, n_leja, radius))
        return qp_fail(ctx, QP_ERR_INTERNAL, "qp_newton_step: Divided differences too small");
    }

    // Newton polynomial in the extended (m+1) x (m+1) Hessenberg block (:330-343)
    auto Hm = [&](int i, int j) { return hess[(size_t)j * ld + i]; };
    auto step_R = [&](cplx shift) {  // R <- (Hm R - shift R) / radius
      for (int i = 0; i <= m; ++i) {
        cplx t(0.0, 0.0);
        for (int j = 0; j <= m; ++j) t += Hm(i, j) * R[j];
        R2[i] = (t - shift * R[i]) / radius;
      }
      std::swap(R, R2);
    };
    {
      HostTimer t(ctx, "evaluate polynomial");  // :330-343
      std::fill(R.begin(), R.end(), cplx(0.0, 0.0));
      R[0] = beta;
      for (int i = 0; i <= m; ++i) P[i] = a[n_s] * R[i];
      for (int k = 1; k < m; ++k) {
        step_R(leja[n_s + k - 1]);
        for (int i = 0; i <= m; ++i) P[i] += a[n_s + k] * R[i];
      }
    }
    // Psi (+)= sum_{i<m} P_i q_i   (:346-352)
    QP_CHECK(qp_krylov_combine(K, reinterpret_cast<const qp_c128*>(P.data()), 0, m, psi, s > 0 ? 1 : 0));
    // restart vector v <- sum_{i<=m} R_i q_i / beta   (:356-367)
    step_R(leja[n_s + m - 1]);
    double b2 = 0.0;
    for (int i = 0; i <= m; ++i) b2 += std::norm(R[i]);
    beta = std::sqrt(b2);
    for (int i = 0; i <= m; ++i) R[i] /= beta;
    QP_CHECK(qp_krylov_combine(K, reinterpret_cast<const qp_c128*>(R.data()), 0, m + 1, v, 0));
    // convergence: relative size of the last Newton term (:370-376)
    double npsi = 0.0;
    QP_CHECK(qp_norm(psi, &npsi));
    if (beta * std::abs(a[n_a - 1]) / (1.0 + npsi) < relerr) break;
    ++s;
    if (s > max_restarts)
      return qp_fail(ctx, QP_ERR_NOT_CONVERGED, "newton!: no convergence within max_restarts=%d", max_restarts);
  }
  if (restarts_out) *restarts_out = s;
  // NewtonWrk bookkeeping of the reference (wrk.n_a, wrk.n_leja, wrk.radius, wrk.a, wrk.leja;
  // src/newton.jl:381-383), readable through qp_newton_last
  K->last_n_a = n_a;
  K->last_n_leja = n_leja;
  K->last_radius = radius;
  K->last_a.assign(reinterpret_cast<const qp_c128*>(a.data()), reinterpret_cast<const qp_c128*>(a.data()) + n_a);
  K->last_leja.assign(reinterpret_cast<const qp_c128*>(leja.data()), reinterpret_cast<const qp_c128*>(leja.data()) + n_leja);
  return QP_OK;
}

extern "C" int32_t qp_newton_last(qp_krylov_t K, int32_t* n_a, int32_t* n_leja, double* radius, qp_c128* a, qp_c128* leja,
                                  int32_t capacity) {
  if (!K) return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_newton_last: null workspace");
  if (n_a) *n_a = K->last_n_a;
  if (n_leja) *n_leja = K->last_n_leja;
  if (radius) *radius = K->last_radius;
  if (a) memcpy(a, K->last_a.data(), sizeof(qp_c128) * std::min<size_t>(K->last_a.size(), (size_t)std::max(capacity, 0)));
  if (leja) memcpy(leja, K->last_leja.data(), sizeof(qp_c128) * std::min<size_t>(K->last_leja.size(), (size_t)std::max(capacity, 0)));
  return QP_OK;
}

// ---------------------------------------------------------------------------------------
// the host-side pieces on their own (no device needed): the same code qp_newton_step runs
// ---------------------------------------------------------------------------------------

extern "C" int32_t qp_diagonalize_hessenberg(const qp_c128* hess, int32_t ld, int32_t m, int32_t accumulate,
                                             qp_c128* out, int32_t* n_out) {
  if (!hess || !out || !n_out || m < 1 || ld < m)
    return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_diagonalize_hessenberg: bad arguments");
  std::vector<cplx> h((size_t)ld * ld), ritz;
  memcpy(static_cast<void*>(h.data()), hess, sizeof(qp_c128) * h.size());
  if (!ritz_accumulated(h, ld, m, ritz))
    return qp_fail(nullptr, QP_ERR_INTERNAL, "qp_diagonalize_hessenberg: QR iteration did not converge");
  const size_t first = accumulate ? 0 : ritz.size() - (size_t)m;  // last block = the m x m matrix itself
  *n_out = (int32_t)(ritz.size() - first);
  memcpy(out, ritz.data() + first, sizeof(qp_c128) * (ritz.size() - first));
  return QP_OK;
}

extern "C" int32_t qp_extend_leja(qp_c128* leja, int32_t capacity, int32_t* n, qp_c128* newpoints, int32_t n_new,
                                  int32_t n_use) {
  if (!leja || !n || !newpoints || n_new < 1 || n_use < 0 || n_use > n_new || *n < 0 || *n + n_use > capacity)
    return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_extend_leja: bad arguments");
  std::vector<cplx> l(reinterpret_cast<cplx*>(leja), reinterpret_cast<cplx*>(leja) + *n);
  std::vector<cplx> cand(reinterpret_cast<cplx*>(newpoints), reinterpret_cast<cplx*>(newpoints) + n_new);
  int nn = *n;
  extend_leja(l, nn, cand, n_use);
  memcpy(leja, l.data(), sizeof(qp_c128) * nn);
  memcpy(newpoints, cand.data(), sizeof(qp_c128) * n_new);
  *n = nn;
  return QP_OK;
}

extern "C" int32_t qp_extend_newton_coeffs(qp_c128* a, int32_t capacity, int32_t* n_a, const qp_c128* leja,
                                           int32_t n_leja, int32_t func_id, qp_newton_func_t func_cb, void* user,
                                           double radius) {
  if (!a || !n_a || !leja || *n_a < 0 || n_leja > capacity || !(radius > 0.0) ||
      !(func_id == QP_FUNC_EXPMI || func_id == QP_FUNC_EXP || (func_id == QP_FUNC_CALLBACK && func_cb)))
    return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_extend_newton_coeffs: bad arguments");
  std::vector<cplx> av(reinterpret_cast<cplx*>(a), reinterpret_cast<cplx*>(a) + *n_a);
  const std::vector<cplx> lv(reinterpret_cast<const cplx*>(leja), reinterpret_cast<const cplx*>(leja) + n_leja);
  int na = *n_a;
  if (!extend_newton_coeffs(av, na, lv, NewtonFunc{func_id, func_cb, user}, n_leja, radius))
    return qp_fail(nullptr, QP_ERR_INTERNAL, "qp_extend_newton_coeffs: Divided differences too small");
  memcpy(a, av.data(), sizeof(qp_c128) * na);
  *n_a = na;
  return QP_OK;
}
