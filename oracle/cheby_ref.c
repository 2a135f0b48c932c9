/*
 * cheby_ref.c -- plain-C restatement of the reference's CPU Chebyshev step, used ONLY as the
 * timed CPU baseline of bench.py (`cpu_baseline`, `--impl reference`) and cross-checked against
 * the NumPy oracle in tests/test_oracle_pins.py.  TEST / MEASUREMENT INFRASTRUCTURE, not product.
 *
 * It reproduces the reference's memory-access pattern, not just its arithmetic:
 *   - cheby!                       reference src/cheby.jl:150-213  (copyto!, lmul!, mul!, 2 axpy!, lmul!, axpy!)
 *   - mul!(C, A::Operator, B, a, b) reference src/generators.jl:634-645: one 5-argument mul! per
 *     component operator, the first with beta, the others accumulating
 *   - 5-argument mul! of a SparseMatrixCSC{ComplexF64,Int64} times a vector: Julia's SparseArrays
 *     scales C by beta, then scatters column by column, single-threaded  [stdlib]
 * `cheby_step_csc` is that faithful single-thread form (Int64 indices, CSC).  `cheby_step_csr_omp`
 * is a stronger baseline the reference does not have: the same step with a row-parallel CSR
 * gather SpMV on all host cores (OpenMP).
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex c128;

/* C <- beta*C + alpha*A*B for a CSC matrix (Julia SparseArrays.mul!, single thread) */
static void csc_mul5(int64_t n, const int64_t* colptr, const int64_t* rowval, const c128* nzval, const c128* B,
                     c128* C, c128 alpha, c128 beta) {
  if (beta != 1.0) {
    if (beta == 0.0)
      memset(C, 0, sizeof(c128) * (size_t)n);
    else
      for (int64_t i = 0; i < n; ++i) C[i] *= beta;
  }
  for (int64_t k = 0; k < n; ++k) {
    const c128 axj = B[k] * alpha;
    for (int64_t j = colptr[k]; j < colptr[k + 1]; ++j) C[rowval[j]] += nzval[j] * axj;
  }
}

/* mul!(C, A::Operator, B, true, false): src/generators.jl:634-645; coeffs for the last n_coeffs ops */
static void operator_mul_csc(int64_t n, int n_ops, const int64_t* const* colptr, const int64_t* const* rowval,
                             const c128* const* nzval, const c128* coeffs, int n_coeffs, const c128* B, c128* C) {
  const int drift = n_ops - n_coeffs;
  for (int l = 0; l < n_ops; ++l) {
    c128 c = 1.0;
    if (l >= drift) c *= coeffs[l - drift];
    csc_mul5(n, colptr[l], rowval[l], nzval[l], B, C, c, l == 0 ? 0.0 : 1.0);
  }
}

/* One cheby! call, reference src/cheby.jl:150-213.  psi is updated in place; v0, v1, v2 are work
 * vectors of length n.  Returns the number of matrix-vector products. */
int cheby_step_csc(int64_t n, int n_ops, const int64_t* const* colptr, const int64_t* const* rowval,
                   const c128* const* nzval, const c128* coeffs, int n_coeffs, c128* psi, c128* v0, c128* v1,
                   c128* v2, const double* a, int n_a, double Delta, double E_min, double dt) {
  const double beta = Delta / 2 + E_min;
  c128 c = (dt > 0 ? -2.0 * I : 2.0 * I) / Delta;
  int n_mv = 0;
  memcpy(v0, psi, sizeof(c128) * (size_t)n);            /* copyto!(v0, psi)   :171 */
  for (int64_t i = 0; i < n; ++i) psi[i] *= a[0];       /* lmul!(a[1], psi)   :172 */
  operator_mul_csc(n, n_ops, colptr, rowval, nzval, coeffs, n_coeffs, v0, v1); /* :175 */
  ++n_mv;
  for (int64_t i = 0; i < n; ++i) v1[i] += -beta * v0[i]; /* axpy!(-beta, v0, v1) :178 */
  for (int64_t i = 0; i < n; ++i) v1[i] *= c;             /* lmul!(c, v1)         :179 */
  for (int64_t i = 0; i < n; ++i) psi[i] += a[1] * v1[i]; /* axpy!(a[2], v1, psi) :182 */
  c *= 2;
  for (int k = 2; k < n_a; ++k) {
    operator_mul_csc(n, n_ops, colptr, rowval, nzval, coeffs, n_coeffs, v1, v2); /* :189 */
    ++n_mv;
    for (int64_t i = 0; i < n; ++i) v2[i] += -beta * v1[i]; /* :192 */
    for (int64_t i = 0; i < n; ++i) v2[i] *= c;             /* :193 */
    for (int64_t i = 0; i < n; ++i) v2[i] += v0[i];         /* :202 */
    for (int64_t i = 0; i < n; ++i) psi[i] += a[k] * v2[i]; /* :205 */
    c128* t = v0;                                           /* :207 */
    v0 = v1;
    v1 = v2;
    v2 = t;
  }
  const c128 phase = cexp(-I * beta * dt);                  /* :211 */
  for (int64_t i = 0; i < n; ++i) psi[i] *= phase;
  return n_mv;
}

/* Stronger (non-reference) baseline: same algorithm, CSR gather SpMV parallel over rows, vector
 * passes fused per term like a hand-optimised CPU code would do. */
int cheby_step_csr_omp(int64_t n, int n_ops, const int64_t* const* rowptr, const int64_t* const* colidx,
                       const c128* const* val, const c128* coeffs, int n_coeffs, c128* psi, c128* v0, c128* v1,
                       c128* v2, const double* a, int n_a, double Delta, double E_min, double dt, int threads) {
  const double beta = Delta / 2 + E_min;
  const c128 c1 = (dt > 0 ? -2.0 * I : 2.0 * I) / Delta;
  const int drift = n_ops - n_coeffs;
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
  memcpy(v0, psi, sizeof(c128) * (size_t)n);
  int n_mv = 0;
  for (int k = 1; k < n_a; ++k) {
    const c128 c = k == 1 ? c1 : 2.0 * c1;
    const c128* x = k == 1 ? v0 : v1;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n; ++r) {
      c128 s = 0.0;
      for (int l = 0; l < n_ops; ++l) {
        c128 t = 0.0;
        for (int64_t j = rowptr[l][r]; j < rowptr[l][r + 1]; ++j) t += val[l][j] * x[colidx[l][j]];
        s += (l >= drift ? coeffs[l - drift] : 1.0) * t;
      }
      if (k == 1) {
        const c128 w = c * (s - beta * x[r]);
        v1[r] = w;
        psi[r] = a[0] * x[r] + a[1] * w;
      } else {
        const c128 w = c * (s - beta * x[r]) + v0[r];
        v2[r] = w;
        psi[r] += a[k] * w;
      }
    }
    ++n_mv;
    if (k > 1) {
      c128* t = v0;
      v0 = v1;
      v1 = v2;
      v2 = t;
    }
  }
  const c128 phase = cexp(-I * beta * dt);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) psi[i] *= phase;
  return n_mv;
}

int cheby_ref_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
