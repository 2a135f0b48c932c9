// Matrix-free left/right super-operators (QP_FORMAT_LR): y = sum_t u_{op(t)} c_t P_t rho Q_t on a
// column-stacked n x n matrix rho (vec index i + n j), fused with the same epilogues as the
// matrix kernels.  This is what `liouvillian(H, c_ops)` means without ever building the n^2 x n^2
// sparse matrix of src/generators.jl:470-508: H rho - rho H + i sum_k (A_k rho A_k^+ - ...) with
// the n x n operators themselves.
//
// Warp = 32 consecutive rows i of one column j of rho (a 512 B contiguous piece of the state), so
//   * a LEFT factor P acts within the column: lane i walks row i of P (CSR) and gathers rho[a, j]
//     from the 16 n bytes of column j (L1/L2-resident while the CTA works on the column);
//   * a RIGHT factor Q is warp-uniform: the entries (b, Q[b, j]) of column j of Q (stored as row j
//     of Q^T) are broadcast, and the gathers rho[i, b] are 512 B contiguous;
//   * a sandwich P rho Q nests the two.
// First version (round 1): correctness and memory footprint first; the small matrices are read
// through L1, the term table sits in shared memory.
#pragma once

#include "spmv.cuh"

constexpr int QP_LR_MAX_TERMS = 256;

template <int EPI>
__global__ void __launch_bounds__(256)
k_spmv_lr(const LRTerm* __restrict__ terms, int n_terms, int n_ops, int64_t nh, const double2* __restrict__ coef,
          const double2* __restrict__ x, EpiArgs e) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  LRTerm* s_terms = reinterpret_cast<LRTerm*>(smem_raw);
  __shared__ double2 s_coef[QP_MAX_OPS];
  if (threadIdx.x < n_ops) s_coef[threadIdx.x] = coef[threadIdx.x];
  for (int t = threadIdx.x; t < n_terms; t += blockDim.x) s_terms[t] = terms[t];
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  const int64_t spc = (nh + 31) >> 5;  // slices per column
  const int64_t n_slices = spc * nh;
  double dr = 0, di = 0, nn = 0;
  for (int64_t s = (int64_t)blockIdx.x * wpc + warp; s < n_slices; s += (int64_t)gridDim.x * wpc) {
    const int64_t j = s / spc;
    const int64_t i = ((s - j * spc) << 5) + lane;
    const bool live = i < nh;
    const int64_t row = i + nh * j;
    double2 xr, yv, av;
    xr = yv = av = make_double2(0.0, 0.0);
    if (live) epi_load<EPI>(e, x, row, row, xr, yv, av);
    const double2* xcol = x + nh * j;
    double sr = 0.0, si = 0.0;
    for (int t = 0; t < n_terms; ++t) {
      const LRTerm& T = s_terms[t];
      const double2 cu = cmul2(s_coef[T.op], T.c);
      double pr = 0.0, pi = 0.0;
      uint32_t l0 = 0, l1 = 0;
      if (T.lptr != nullptr && live) {
        l0 = __ldg(T.lptr + i);
        l1 = __ldg(T.lptr + i + 1);
      }
      if (T.rptr == nullptr) {
        if (T.lptr == nullptr) {  // c * rho
          if (live) {
            const double2 xo = EPI == EPI_MUL ? __ldg(x + row) : xr;
            pr = xo.x;
            pi = xo.y;
          }
        } else {  // P rho
          for (uint32_t k = l0; k < l1; ++k) {
            const double2 v = __ldg(T.lval + k);
            const double2 xv = __ldg(xcol + __ldg(T.lcol + k));
            pr += v.x * xv.x - v.y * xv.y;
            pi += v.x * xv.y + v.y * xv.x;
          }
        }
      } else {
        const uint32_t q0 = __ldg(T.rptr + j), q1 = __ldg(T.rptr + j + 1);  // warp-uniform
        for (uint32_t kb = q0; kb < q1; ++kb) {
          const double2 vb = __ldg(T.rval + kb);
          const double2* xb = x + nh * (int64_t)__ldg(T.rcol + kb);
          double ar = 0.0, ai = 0.0;
          if (T.lptr == nullptr) {  // rho Q
            if (live) {
              const double2 xv = __ldg(xb + i);
              ar = xv.x;
              ai = xv.y;
            }
          } else {  // P rho Q
            for (uint32_t k = l0; k < l1; ++k) {
              const double2 v = __ldg(T.lval + k);
              const double2 xv = __ldg(xb + __ldg(T.lcol + k));
              ar += v.x * xv.x - v.y * xv.y;
              ai += v.x * xv.y + v.y * xv.x;
            }
          }
          pr += vb.x * ar - vb.y * ai;
          pi += vb.x * ai + vb.y * ar;
        }
      }
      sr += cu.x * pr - cu.y * pi;
      si += cu.x * pi + cu.y * pr;
    }
    if (live) epi_apply<EPI>(e, row, make_double2(sr, si), xr, yv, av, dr, di, nn);
  }
  if (epi_has_sums(EPI) && e.chk != nullptr) {
    for (int o = 16; o > 0; o >>= 1) {
      dr += __shfl_xor_sync(0xffffffffu, dr, o);
      di += __shfl_xor_sync(0xffffffffu, di, o);
      nn += __shfl_xor_sync(0xffffffffu, nn, o);
    }
    if (lane == 0) chk_flush(e, 0, dr, di, nn);
  }
}
