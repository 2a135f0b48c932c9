// Fused multi-operator SpMV kernels for sm_100a.
//
// One kernel family applies H = sum_l u_l H_l (never materialised: the merged matrix keeps
// every operator's own values, tagged with the operator index in the top 4 bits of the
// column word, and u_l is applied at use) and finishes with a fused epilogue:
//
//   EPI_MUL         y <- beta*y + alpha*Hx                     mul!, src/generators.jl:634-645
//   EPI_CHEB_FIRST  v1 = c(Hx - b x);  y <- v1;  acc <- a0 x + a1 v1      src/cheby.jl:171-182
//   EPI_CHEB_MID    v2 = c(Hx - b x) + y;  y <- v2;  acc += ak v2          src/cheby.jl:186-209
//   EPI_CHEB_LAST   v2 as MID; acc <- phase (acc + ak v2)  (v2 not stored) src/cheby.jl:211
//   EPI_CHEB_ONLY   n_coeffs == 2: acc <- phase (a0 x + a1 c(Hx - b x))
//
// so one Chebyshev term is ONE pass: the reference's mul! + 2 axpy! + lmul! + axpy!
// (src/cheby.jl:189-205) and the per-operator read-modify-write of the result collapse
// into M + 80 N B bytes of traffic.  v_{k+1} overwrites v_{k-1} in place
// (same element, same thread).
//
// Storage formats (chosen per generator at qp_gen_create):
//   merged CSR   -- sub-warp of LANES lanes per row, 128-bit coalesced value loads,
//                   shuffle reduction; good for short grids / long rows.
//   SELL-32      -- sliced ELL, slice height 32 = one warp, entries column-major inside a
//                   slice so every warp-wide load is one 512 B (values) / 128 B (columns)
//                   contiguous segment and neighbouring rows gather neighbouring x entries;
//                   thread per row, no reduction.  The big-N format.
//   batched      -- state layout [N][B], thread per (row, trajectory): the B threads of a
//                   row broadcast-load the matrix entry and gather 16 B-contiguous x values.
#pragma once

#include "qprop_internal.h"

enum { EPI_MUL = 0, EPI_CHEB_FIRST = 1, EPI_CHEB_MID = 2, EPI_CHEB_LAST = 3, EPI_CHEB_ONLY = 4 };

struct EpiArgs {
  double2 alpha, betac;  // EPI_MUL
  double2 c;             // Chebyshev prefactor (c for the first term, 2c afterwards)
  double beta;           // Chebyshev shift  Delta/2 + E_min
  double a0, ak;         // coefficients a_1 (FIRST/ONLY) and a_k
  double2 phase;         // exp(-i beta dt)
  double2* y;            // MUL: y;  FIRST: v1 out;  MID/LAST: v_{k-1} in, v_{k+1} out
  double2* acc;          // psi accumulator
  double* chk;           // normalization-check accumulators [batch][3] or nullptr
};

__device__ __forceinline__ double2 cmul2(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// streaming loads for the matrix arrays: read-only path, do not allocate in L1 (keep L1 for
// the gathered x entries, which are the only data with reuse)
__device__ __forceinline__ double2 ld_stream(const double2* p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t ld_stream(const uint32_t* p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}

template <int EPI>
__device__ __forceinline__ void epilogue(const EpiArgs& e, int64_t idx, int64_t b, double2 hx,
                                         double2 xr, double& chk_dr, double& chk_di, double& chk_n) {
  if (EPI == EPI_MUL) {
    double2 r = cmul2(e.alpha, hx);
    if (e.betac.x != 0.0 || e.betac.y != 0.0) {
      double2 t = cmul2(e.betac, e.y[idx]);
      r.x += t.x;
      r.y += t.y;
    }
    e.y[idx] = r;
    return;
  }
  double2 t = make_double2(hx.x - e.beta * xr.x, hx.y - e.beta * xr.y);
  double2 v = cmul2(e.c, t);  // c (Hx - beta x)
  if (EPI == EPI_CHEB_FIRST) {
    e.y[idx] = v;
    e.acc[idx] = make_double2(e.a0 * xr.x + e.ak * v.x, e.a0 * xr.y + e.ak * v.y);
  } else if (EPI == EPI_CHEB_ONLY) {
    double2 s = make_double2(e.a0 * xr.x + e.ak * v.x, e.a0 * xr.y + e.ak * v.y);
    e.acc[idx] = cmul2(e.phase, s);
  } else {
    if (e.chk != nullptr) {  // <v1|v2'> and |v1|^2, src/cheby.jl:194-200
      chk_dr += xr.x * v.x + xr.y * v.y;
      chk_di += xr.x * v.y - xr.y * v.x;
      chk_n += xr.x * xr.x + xr.y * xr.y;
    }
    double2 p = e.y[idx];
    v.x += p.x;
    v.y += p.y;
    double2 a = e.acc[idx];
    a.x += e.ak * v.x;
    a.y += e.ak * v.y;
    if (EPI == EPI_CHEB_MID) {
      e.y[idx] = v;
      e.acc[idx] = a;
    } else {
      e.acc[idx] = cmul2(e.phase, a);
    }
  }
}

// warp-level flush of the normalization-check partial sums (debug option, atomics)
__device__ __forceinline__ void chk_flush(const EpiArgs& e, int64_t b, double dr, double di, double nn) {
  if (e.chk == nullptr) return;
  atomicAdd(e.chk + 3 * b + 0, dr);
  atomicAdd(e.chk + 3 * b + 1, di);
  atomicAdd(e.chk + 3 * b + 2, nn);
}

// ---------------------------------------------------------------------------------------
// merged CSR, batch == 1: LANES lanes per row
// ---------------------------------------------------------------------------------------
template <int LANES, int EPI>
__global__ void __launch_bounds__(256)
k_spmv_csr(MatView m, const double2* __restrict__ coef, int n_ops, const double2* __restrict__ x,
           EpiArgs e) {
  __shared__ double2 s_coef[QP_MAX_OPS];
  if (threadIdx.x < n_ops) s_coef[threadIdx.x] = coef[threadIdx.x];
  __syncthreads();

  const int64_t gtid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t row = gtid / LANES;
  const int lane = (int)(gtid % LANES);
  double sr = 0.0, si = 0.0;
  const bool active = row < m.n;
  if (active) {
    const uint32_t p0 = m.ptr[row], p1 = m.ptr[row + 1];
#pragma unroll 4
    for (uint32_t k = p0 + lane; k < p1; k += LANES) {
      const uint32_t co = ld_stream(m.colop + k);
      const double2 v = ld_stream(m.val + k);
      const double2 xv = __ldg(x + (co & QP_COL_MASK));
      const double2 u = s_coef[co >> QP_COL_BITS];
      const double tr = v.x * xv.x - v.y * xv.y;
      const double ti = v.x * xv.y + v.y * xv.x;
      sr += u.x * tr - u.y * ti;
      si += u.x * ti + u.y * tr;
    }
  }
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) {
    sr += __shfl_xor_sync(0xffffffffu, sr, o);
    si += __shfl_xor_sync(0xffffffffu, si, o);
  }
  double dr = 0, di = 0, nn = 0;
  if (active && lane == 0) {
    double2 xr = make_double2(0.0, 0.0);
    if (EPI != EPI_MUL) xr = __ldg(x + row);
    epilogue<EPI>(e, row, 0, make_double2(sr, si), xr, dr, di, nn);
  }
  if ((EPI == EPI_CHEB_MID || EPI == EPI_CHEB_LAST) && e.chk != nullptr) {
    for (int o = 16; o > 0; o >>= 1) {
      dr += __shfl_xor_sync(0xffffffffu, dr, o);
      di += __shfl_xor_sync(0xffffffffu, di, o);
      nn += __shfl_xor_sync(0xffffffffu, nn, o);
    }
    if ((threadIdx.x & 31) == 0) chk_flush(e, 0, dr, di, nn);
  }
}

// ---------------------------------------------------------------------------------------
// SELL-32, batch == 1: thread per row, one warp per slice
// ---------------------------------------------------------------------------------------
template <int EPI>
__global__ void __launch_bounds__(256)
k_spmv_sell(MatView m, const double2* __restrict__ coef, int n_ops, const double2* __restrict__ x,
            EpiArgs e) {
  __shared__ double2 s_coef[QP_MAX_OPS];
  if (threadIdx.x < n_ops) s_coef[threadIdx.x] = coef[threadIdx.x];
  __syncthreads();

  const int64_t n_slices = (m.n + QP_SELL_C - 1) / QP_SELL_C;
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
  double dr = 0, di = 0, nn = 0;
  for (int64_t s = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; s < n_slices;
       s += warps_total) {
    const int64_t row = s * QP_SELL_C + lane;
    const uint32_t p0 = m.ptr[s], p1 = m.ptr[s + 1];
    double sr = 0.0, si = 0.0;
#pragma unroll 4
    for (uint32_t k = p0 + lane; k < p1; k += QP_SELL_C) {
      const uint32_t co = ld_stream(m.colop + k);
      const double2 v = ld_stream(m.val + k);
      const double2 xv = __ldg(x + (co & QP_COL_MASK));
      const double2 u = s_coef[co >> QP_COL_BITS];
      const double tr = v.x * xv.x - v.y * xv.y;
      const double ti = v.x * xv.y + v.y * xv.x;
      sr += u.x * tr - u.y * ti;
      si += u.x * ti + u.y * tr;
    }
    if (row < m.n) {
      double2 xr = make_double2(0.0, 0.0);
      if (EPI != EPI_MUL) xr = __ldg(x + row);
      epilogue<EPI>(e, row, 0, make_double2(sr, si), xr, dr, di, nn);
    }
  }
  if ((EPI == EPI_CHEB_MID || EPI == EPI_CHEB_LAST) && e.chk != nullptr) {
    for (int o = 16; o > 0; o >>= 1) {
      dr += __shfl_xor_sync(0xffffffffu, dr, o);
      di += __shfl_xor_sync(0xffffffffu, di, o);
      nn += __shfl_xor_sync(0xffffffffu, nn, o);
    }
    if (lane == 0) chk_flush(e, 0, dr, di, nn);
  }
}

// ---------------------------------------------------------------------------------------
// merged CSR, trajectory-batched: thread per (row, b); state layout [N][B]
// coef layout [n_ops][coef_stride ? B : 1]
// ---------------------------------------------------------------------------------------
template <int EPI>
__global__ void __launch_bounds__(256)
k_spmm_csr(MatView m, const double2* __restrict__ coef, int coef_stride, int64_t batch,
           const double2* __restrict__ x, EpiArgs e) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t total = m.n * batch;
  if (idx >= total) return;
  const int64_t row = idx / batch;
  const int64_t b = idx - row * batch;
  const uint32_t p0 = m.ptr[row], p1 = m.ptr[row + 1];
  double sr = 0.0, si = 0.0;
  uint32_t cur_op = 0xffffffffu;
  double2 u = make_double2(1.0, 0.0);
#pragma unroll 2
  for (uint32_t k = p0; k < p1; ++k) {
    const uint32_t co = __ldg(m.colop + k);
    const double2 v = __ldg(m.val + k);
    const uint32_t op = co >> QP_COL_BITS;
    if (op != cur_op) {
      cur_op = op;
      u = __ldg(coef + (int64_t)op * (coef_stride ? batch : 1) + (coef_stride ? b : 0));
    }
    const double2 xv = __ldg(x + (int64_t)(co & QP_COL_MASK) * batch + b);
    const double tr = v.x * xv.x - v.y * xv.y;
    const double ti = v.x * xv.y + v.y * xv.x;
    sr += u.x * tr - u.y * ti;
    si += u.x * ti + u.y * tr;
  }
  double2 xr = make_double2(0.0, 0.0);
  if (EPI != EPI_MUL) xr = __ldg(x + idx);
  double dr = 0, di = 0, nn = 0;
  epilogue<EPI>(e, idx, b, make_double2(sr, si), xr, dr, di, nn);
  if ((EPI == EPI_CHEB_MID || EPI == EPI_CHEB_LAST) && e.chk != nullptr) chk_flush(e, b, dr, di, nn);
}

// ---------------------------------------------------------------------------------------
// dense row-major operators: one warp per (row) for batch == 1
// ---------------------------------------------------------------------------------------
template <int EPI>
__global__ void __launch_bounds__(256)
k_gemv_dense(const double2* const* __restrict__ ops, int n_ops, int64_t n,
             const double2* __restrict__ coef, const double2* __restrict__ x, EpiArgs e) {
  __shared__ double2 s_coef[QP_MAX_OPS];
  if (threadIdx.x < n_ops) s_coef[threadIdx.x] = coef[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  double sr = 0.0, si = 0.0;
  if (row < n) {
    for (int l = 0; l < n_ops; ++l) {
      const double2* __restrict__ a = ops[l] + row * n;
      const double2 u = s_coef[l];
      double pr = 0.0, pi = 0.0;
#pragma unroll 4
      for (int64_t j = lane; j < n; j += 32) {
        const double2 v = ld_stream(a + j);
        const double2 xv = __ldg(x + j);
        pr += v.x * xv.x - v.y * xv.y;
        pi += v.x * xv.y + v.y * xv.x;
      }
      sr += u.x * pr - u.y * pi;
      si += u.x * pi + u.y * pr;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    sr += __shfl_xor_sync(0xffffffffu, sr, o);
    si += __shfl_xor_sync(0xffffffffu, si, o);
  }
  double dr = 0, di = 0, nn = 0;
  if (row < n && lane == 0) {
    double2 xr = make_double2(0.0, 0.0);
    if (EPI != EPI_MUL) xr = __ldg(x + row);
    epilogue<EPI>(e, row, 0, make_double2(sr, si), xr, dr, di, nn);
  }
  if ((EPI == EPI_CHEB_MID || EPI == EPI_CHEB_LAST) && e.chk != nullptr) {
    for (int o = 16; o > 0; o >>= 1) {
      dr += __shfl_xor_sync(0xffffffffu, dr, o);
      di += __shfl_xor_sync(0xffffffffu, di, o);
      nn += __shfl_xor_sync(0xffffffffu, nn, o);
    }
    if (lane == 0) chk_flush(e, 0, dr, di, nn);
  }
}

// host-side dispatcher (defined in sparse.cu)
int32_t qp_launch_fused(qp_gen_t gen, int epi, int coef_stride, const double2* x, int64_t batch,
                        const EpiArgs& e);
