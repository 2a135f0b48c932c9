// Matrix-free left/right super-operators (QP_FORMAT_LR): y = sum_t u_{op(t)} c_t P_t rho Q_t on a
// column-stacked n x n matrix rho (vec index i + n j), fused with the same epilogues as the
// matrix kernels.  This is what `liouvillian(H, c_ops)` means without ever building the n^2 x n^2
// sparse matrix of src/generators.jl:470-508: H rho - rho H + i sum_k (A_k rho A_k^+ - ...) with
// the n x n operators themselves.
//
// Warp = 32 consecutive rows i of JB consecutive columns j of rho (JB 512 B pieces of the state):
//   * a LEFT factor P acts within a column: lane i walks row i of P (CSR) ONCE per JB columns and
//     gathers rho[a, j] for each of them; a warp keeps its rows for a whole sweep over its column
//     range, so the rows of P it needs stay in L1;
//   * a RIGHT factor Q is warp-uniform: the entries (b, Q[b, j]) of column j of Q (stored as row j
//     of Q^T) are broadcast, and the gathers rho[i, b] are 512 B contiguous;
//   * a sandwich P rho Q nests the two.
// The term table sits in shared memory; the epilogue operands of the JB rows are loaded after the
// terms (not held in registers across them).
#pragma once

#include "spmv.cuh"

constexpr int QP_LR_MAX_TERMS = 256;

template <int EPI, int JB>
__global__ void __launch_bounds__(256, 2)
k_spmv_lr(const LRTerm* __restrict__ terms, int n_terms, int n_ops, int64_t nh, const double2* __restrict__ coef,
          const double2* __restrict__ x, EpiArgs e, int cols_per_cta) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  LRTerm* s_terms = reinterpret_cast<LRTerm*>(smem_raw);
  __shared__ double2 s_coef[QP_MAX_OPS];
  if (threadIdx.x < n_ops) s_coef[threadIdx.x] = coef[threadIdx.x];
  for (int t = threadIdx.x; t < n_terms; t += blockDim.x) s_terms[t] = terms[t];
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // row of rho, fixed for the whole sweep
  const bool live = i < nh;
  const int64_t j_begin = (int64_t)blockIdx.y * cols_per_cta;
  const int64_t j_end = min(j_begin + cols_per_cta, nh);
  double dr = 0, di = 0, nn = 0;
  if ((int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31) >= nh) return;  // warp entirely beyond the rows
  for (int64_t j0 = j_begin; j0 < j_end; j0 += JB) {
    double sr[JB], si[JB];
#pragma unroll
    for (int jj = 0; jj < JB; ++jj) sr[jj] = si[jj] = 0.0;
    for (int t = 0; t < n_terms; ++t) {
      const LRTerm& T = s_terms[t];
      const double2 cu = cmul2(s_coef[T.op], T.c);
      double pr[JB], pi[JB];
#pragma unroll
      for (int jj = 0; jj < JB; ++jj) pr[jj] = pi[jj] = 0.0;
      uint32_t l0 = 0, l1 = 0;
      if (T.lptr != nullptr && live) {
        l0 = __ldg(T.lptr + i);
        l1 = __ldg(T.lptr + i + 1);
      }
      if (T.rptr == nullptr) {
        if (T.lptr == nullptr) {  // c * rho
          if (live) {
#pragma unroll
            for (int jj = 0; jj < JB; ++jj)
              if (j0 + jj < j_end) {
                const double2 xo = __ldg(x + i + nh * (j0 + jj));
                pr[jj] = xo.x;
                pi[jj] = xo.y;
              }
          }
        } else {  // P rho: row i of P decoded once for the JB columns, two entries (2 JB gathers) in flight
          for (uint32_t k = l0; k < l1; k += 2) {
            const bool two = k + 1 < l1;
            const double2 v0 = __ldg(T.lval + k);
            const double2 v1 = two ? __ldg(T.lval + k + 1) : make_double2(0.0, 0.0);
            const double2* xa0 = x + __ldg(T.lcol + k) + nh * j0;
            const double2* xa1 = two ? x + __ldg(T.lcol + k + 1) + nh * j0 : xa0;
            double2 xv0[JB], xv1[JB];
#pragma unroll
            for (int jj = 0; jj < JB; ++jj) {
              const bool ok = j0 + jj < j_end;
              xv0[jj] = ok ? __ldg(xa0 + nh * jj) : make_double2(0.0, 0.0);
              xv1[jj] = ok ? __ldg(xa1 + nh * jj) : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int jj = 0; jj < JB; ++jj) {
              pr[jj] = fma(v0.x, xv0[jj].x, fma(-v0.y, xv0[jj].y, fma(v1.x, xv1[jj].x, fma(-v1.y, xv1[jj].y, pr[jj]))));
              pi[jj] = fma(v0.x, xv0[jj].y, fma(v0.y, xv0[jj].x, fma(v1.x, xv1[jj].y, fma(v1.y, xv1[jj].x, pi[jj]))));
            }
          }
        }
      } else {
        // right factor: the entries of the JB columns are walked in lockstep (entry e of every
        // column at once), so JB gathers -- or JB row walks of P -- are in flight together
        uint32_t q0[JB], qn[JB];
        uint32_t qmax = 0;
#pragma unroll
        for (int jj = 0; jj < JB; ++jj) {
          q0[jj] = qn[jj] = 0;
          if (j0 + jj < j_end) {
            q0[jj] = __ldg(T.rptr + j0 + jj);
            qn[jj] = __ldg(T.rptr + j0 + jj + 1) - q0[jj];
          }
          qmax = max(qmax, qn[jj]);
        }
        for (uint32_t eidx = 0; eidx < qmax; ++eidx) {  // warp-uniform trip count
          double2 vb[JB];
          const double2* xb[JB];
          double ar[JB], ai[JB];
#pragma unroll
          for (int jj = 0; jj < JB; ++jj) {
            const bool has = eidx < qn[jj];
            vb[jj] = has ? __ldg(T.rval + q0[jj] + eidx) : make_double2(0.0, 0.0);
            xb[jj] = x + nh * (int64_t)(has ? __ldg(T.rcol + q0[jj] + eidx) : 0u);
            ar[jj] = ai[jj] = 0.0;
          }
          if (T.lptr == nullptr) {  // rho Q
            if (live) {
              double2 xv[JB];
#pragma unroll
              for (int jj = 0; jj < JB; ++jj) xv[jj] = __ldg(xb[jj] + i);
#pragma unroll
              for (int jj = 0; jj < JB; ++jj) {
                ar[jj] = xv[jj].x;
                ai[jj] = xv[jj].y;
              }
            }
          } else {  // P rho Q
            for (uint32_t k = l0; k < l1; ++k) {
              const double2 v = __ldg(T.lval + k);
              const uint32_t a = __ldg(T.lcol + k);
              double2 xv[JB];
#pragma unroll
              for (int jj = 0; jj < JB; ++jj) xv[jj] = __ldg(xb[jj] + a);
#pragma unroll
              for (int jj = 0; jj < JB; ++jj) {
                ar[jj] = fma(v.x, xv[jj].x, fma(-v.y, xv[jj].y, ar[jj]));
                ai[jj] = fma(v.x, xv[jj].y, fma(v.y, xv[jj].x, ai[jj]));
              }
            }
          }
#pragma unroll
          for (int jj = 0; jj < JB; ++jj) {
            pr[jj] = fma(vb[jj].x, ar[jj], fma(-vb[jj].y, ai[jj], pr[jj]));
            pi[jj] = fma(vb[jj].x, ai[jj], fma(vb[jj].y, ar[jj], pi[jj]));
          }
        }
      }
#pragma unroll
      for (int jj = 0; jj < JB; ++jj) {
        sr[jj] = fma(cu.x, pr[jj], fma(-cu.y, pi[jj], sr[jj]));
        si[jj] = fma(cu.x, pi[jj], fma(cu.y, pr[jj], si[jj]));
      }
    }
    if (live) {
#pragma unroll
      for (int jj = 0; jj < JB; ++jj)
        if (j0 + jj < j_end) {
          const int64_t row = i + nh * (j0 + jj);
          double2 xr, yv, av;
          epi_load<EPI>(e, x, row, row, xr, yv, av);
          epi_apply<EPI>(e, row, make_double2(sr[jj], si[jj]), xr, yv, av, dr, di, nn);
        }
    }
  }
  if (epi_has_sums(EPI) && e.chk != nullptr) {
    for (int o = 16; o > 0; o >>= 1) {
      dr += __shfl_xor_sync(0xffffffffu, dr, o);
      di += __shfl_xor_sync(0xffffffffu, di, o);
      nn += __shfl_xor_sync(0xffffffffu, nn, o);
    }
    if (lane == 0) chk_flush_warp(e, dr, di, nn);
  }
}
