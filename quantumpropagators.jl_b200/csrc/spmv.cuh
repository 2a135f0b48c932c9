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
//   EPI_DOT         <x|Hx> accumulated per trajectory, nothing stored      dot(x, A, x), src/generators.jl:648-660
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

enum { EPI_MUL = 0, EPI_CHEB_FIRST = 1, EPI_CHEB_MID = 2, EPI_CHEB_LAST = 3, EPI_CHEB_ONLY = 4, EPI_DOT = 5 };

// epilogues that leave per-trajectory sums in e.chk[b*3 + {0,1,2}] (flushed with atomics)
__host__ __device__ constexpr bool epi_has_sums(int epi) {
  return epi == EPI_CHEB_MID || epi == EPI_CHEB_LAST || epi == EPI_DOT;
}

struct EpiArgs {
  double2 alpha, betac;  // EPI_MUL
  const double* alpha_dev;  // EPI_MUL: alpha is multiplied by *alpha_dev (a real scale factor that is still on the
                            // device: the pending normalisation of a Krylov vector, krylov.cu); nullptr = 1
  double2 c;             // Chebyshev prefactor (c for the first term, 2c afterwards)
  double beta;           // Chebyshev shift  Delta/2 + E_min
  double a0, ak;         // coefficients a_1 (FIRST/ONLY) and a_k
  double2 phase;         // exp(-i beta dt)
  double2* y;            // MUL: y;  FIRST: v1 out;  MID/LAST: v_{k-1} in, v_{k+1} out
  double2* acc;          // psi accumulator
  double* chk;           // normalization-check accumulators [batch][3] or nullptr
  double* part;          // deterministic mode: per-warp (single states) / per-tile (tiled batched kernel)
                         // partial sums, added up in a fixed order by k_part_reduce (nullptr: atomics)
  int vec_hint;          // 1: access the Chebyshev vectors with an L2 evict_last policy (TMA kernel)
};

__device__ __forceinline__ double2 cmul2(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// Loads of vectors WRITTEN BY THE PREVIOUS KERNEL inside kernels that use programmatic dependent launch
// (k_spmv_csr, k_spmv_selld, k_spmv_bitflip).  Non-coherent loads (__ldg, const __restrict__, ld.global.nc) are
// invariant loads to the compiler AND to ptxas, and both are free to move them above griddepcontrol.wait -- seen
// in the SASS of k_spmv_bitflip (LDG.E.128.CONSTANT of the first slice's own x above ACQBULK, with __ldg and with
// a volatile ld.global.nc alike) as rare stale reads in back-to-back terms.  A plain ld.global (still allocating
// in L1) is ordered after the wait.
__device__ __forceinline__ double2 ld_x(const double2* p) {
  double2 r;
  asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));  // volatile: keeps its place after the (volatile) wait
  return r;
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

// L2 cache-policy helpers (createpolicy + .L2::cache_hint)
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ double2 ld_hint(const double2* p, uint64_t policy) {
  double2 r;
  asm volatile("ld.global.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(r.x), "=d"(r.y) : "l"(p), "l"(policy));
  return r;
}
__device__ __forceinline__ double2 ld_nc_hint(const double2* p, uint64_t policy) {
  double2 r;
  asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(r.x), "=d"(r.y) : "l"(p), "l"(policy));
  return r;
}
__device__ __forceinline__ void st_hint(double2* p, double2 v, uint64_t policy) {
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(policy)
               : "memory");
}
__device__ __forceinline__ double2 vld(const double2* p, int hint) {
  return hint ? ld_hint(p, policy_evict_last()) : *p;
}
__device__ __forceinline__ double2 vldg(const double2* p, int hint) {
  return hint ? ld_nc_hint(p, policy_evict_last()) : __ldg(p);
}
__device__ __forceinline__ void vst(double2* p, double2 v, int hint) {
  if (hint) st_hint(p, v, policy_evict_last());
  else *p = v;
}

// The epilogue is split in two so that kernels can request its operands (x_r, v_{k-1}[r],
// acc[r]) at the START of a row block and consume them after the SpMV: their DRAM latency
// is then hidden behind the matrix stream instead of being exposed once per block.
// PDL: the kernel is launched with programmatic dependent launch -- x must be read with a coherent load (ld_x);
// everywhere else the read-only path is fine
template <int EPI, int PDL = 0>
__device__ __forceinline__ void epi_load(const EpiArgs& e, const double2* __restrict__ x, int64_t xidx,
                                         int64_t idx, double2& xr, double2& yv, double2& av) {
  xr = yv = av = make_double2(0.0, 0.0);
  if (EPI == EPI_MUL) {
    if (e.betac.x != 0.0 || e.betac.y != 0.0) yv = e.y[idx];  // beta == 0: y is not read (BLAS)
    return;
  }
  xr = PDL ? ld_x(x + xidx) : vldg(x + xidx, e.vec_hint);
  if (EPI == EPI_CHEB_MID || EPI == EPI_CHEB_LAST) {
    yv = vld(e.y + idx, e.vec_hint);
    av = vld(e.acc + idx, e.vec_hint);
  }
}

template <int EPI>
__device__ __forceinline__ void epi_apply(const EpiArgs& e, int64_t idx, double2 hx, double2 xr, double2 yv,
                                          double2 av, double& chk_dr, double& chk_di, double& chk_n) {
  if (EPI == EPI_MUL) {
    double2 r = cmul2(e.alpha, hx);
    if (e.alpha_dev != nullptr) {  // written by the previous kernel of the stream: a coherent load, after the PDL wait
      double sc;
      asm volatile("ld.global.f64 %0, [%1];" : "=d"(sc) : "l"(e.alpha_dev) : "memory");
      r.x *= sc;
      r.y *= sc;
    }
    if (e.betac.x != 0.0 || e.betac.y != 0.0) {
      const double2 t = cmul2(e.betac, yv);
      r.x += t.x;
      r.y += t.y;
    }
    e.y[idx] = r;
    return;
  }
  if (EPI == EPI_DOT) {  // conj(x_r) (Hx)_r
    chk_dr += xr.x * hx.x + xr.y * hx.y;
    chk_di += xr.x * hx.y - xr.y * hx.x;
    return;
  }
  const double2 t = make_double2(hx.x - e.beta * xr.x, hx.y - e.beta * xr.y);
  double2 v = cmul2(e.c, t);  // c (Hx - beta x)
  if (EPI == EPI_CHEB_FIRST) {
    vst(e.y + idx, v, e.vec_hint);
    vst(e.acc + idx, make_double2(e.a0 * xr.x + e.ak * v.x, e.a0 * xr.y + e.ak * v.y), e.vec_hint);
  } else if (EPI == EPI_CHEB_ONLY) {
    const double2 s = make_double2(e.a0 * xr.x + e.ak * v.x, e.a0 * xr.y + e.ak * v.y);
    vst(e.acc + idx, cmul2(e.phase, s), e.vec_hint);
  } else {
    if (e.chk != nullptr) {  // <v1|v2'> and |v1|^2, src/cheby.jl:194-200
      chk_dr += xr.x * v.x + xr.y * v.y;
      chk_di += xr.x * v.y - xr.y * v.x;
      chk_n += xr.x * xr.x + xr.y * xr.y;
    }
    v.x += yv.x;
    v.y += yv.y;
    av.x += e.ak * v.x;
    av.y += e.ak * v.y;
    if (EPI == EPI_CHEB_MID) {
      vst(e.y + idx, v, e.vec_hint);
      vst(e.acc + idx, av, e.vec_hint);
    } else {
      vst(e.acc + idx, cmul2(e.phase, av), e.vec_hint);
    }
  }
}

template <int EPI>
__device__ __forceinline__ void epilogue(const EpiArgs& e, const double2* __restrict__ x, int64_t xidx,
                                         int64_t idx, double2 hx, double& chk_dr, double& chk_di,
                                         double& chk_n) {
  double2 xr, yv, av;
  epi_load<EPI>(e, x, xidx, idx, xr, yv, av);
  epi_apply<EPI>(e, idx, hx, xr, yv, av, chk_dr, chk_di, chk_n);
}

// warp-level flush of the normalization-check partial sums (debug option, atomics)
__device__ __forceinline__ void chk_flush(const EpiArgs& e, int64_t b, double dr, double di, double nn) {
  if (e.chk == nullptr) return;
  atomicAdd(e.chk + 3 * b + 0, dr);
  atomicAdd(e.chk + 3 * b + 1, di);
  atomicAdd(e.chk + 3 * b + 2, nn);
}

// single-state kernels: one partial per warp.  With e.part set the warp stores its sums in its own slot
// (global warp index) and a second kernel adds the slots up in a fixed order -- run-to-run reproducible
// expectation values -- otherwise atomics.
__device__ __forceinline__ void chk_flush_warp(const EpiArgs& e, double dr, double di, double nn) {
  if (e.chk == nullptr) return;
  if (e.part != nullptr) {
    const int64_t gw = (((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    e.part[3 * gw + 0] = dr;
    e.part[3 * gw + 1] = di;
    e.part[3 * gw + 2] = nn;
    return;
  }
  chk_flush(e, 0, dr, di, nn);
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
  // Programmatic dependent launch: at small N a Chebyshev term is a chain of dependent L2 round
  // trips (row pointer -> entries -> gathered x -> epilogue) behind a kernel launch.  Everything
  // that does not depend on the previous term -- the coefficients, the row pointers, the first
  // four entries per lane -- is fetched BEFORE griddepcontrol.wait, i.e. while the previous term
  // is still running (the launch itself overlaps too); only the x gathers and the epilogue
  // operands wait for it.  Without the launch attribute both instructions are no-ops.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  uint32_t p0 = 0, p1 = 0;
  uint32_t pco[4];
  double2 pv[4];
  if (active) {
    p0 = m.ptr[row];
    p1 = m.ptr[row + 1];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t k = p0 + lane + i * LANES;
      pco[i] = 0u;
      pv[i] = make_double2(0.0, 0.0);
      if (k < p1) {
        pco[i] = ld_stream(m.colop + k);
        pv[i] = ld_stream(m.val + k);
      }
    }
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // epilogue operands requested before the gathers are consumed (one round trip less)
  double2 xr, yv, av;
  xr = yv = av = make_double2(0.0, 0.0);
  if (active && lane == 0) epi_load<EPI, 1>(e, x, row, row, xr, yv, av);
  if (active) {
    double2 pxv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) pxv[i] = ld_x(x + (pco[i] & QP_COL_MASK));  // padding: column 0, value 0
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double2 u = s_coef[pco[i] >> QP_COL_BITS];
      const double tr = pv[i].x * pxv[i].x - pv[i].y * pxv[i].y;
      const double ti = pv[i].x * pxv[i].y + pv[i].y * pxv[i].x;
      sr += u.x * tr - u.y * ti;
      si += u.x * ti + u.y * tr;
    }
#pragma unroll 4
    for (uint32_t k = p0 + lane + 4 * LANES; k < p1; k += LANES) {
      const uint32_t co = ld_stream(m.colop + k);
      const double2 v = ld_stream(m.val + k);
      const double2 xv = ld_x(x + (co & QP_COL_MASK));
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
  if (active && lane == 0) epi_apply<EPI>(e, row, make_double2(sr, si), xr, yv, av, dr, di, nn);
  if (epi_has_sums(EPI) && e.chk != nullptr) {
    for (int o = 16; o > 0; o >>= 1) {
      dr += __shfl_xor_sync(0xffffffffu, dr, o);
      di += __shfl_xor_sync(0xffffffffu, di, o);
      nn += __shfl_xor_sync(0xffffffffu, nn, o);
    }
    if ((threadIdx.x & 31) == 0) chk_flush_warp(e, dr, di, nn);
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
    double2 xr, yv, av;
    if (row < m.n) epi_load<EPI>(e, x, row, row, xr, yv, av);  // consumed after the slice
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
    if (row < m.n) epi_apply<EPI>(e, row, make_double2(sr, si), xr, yv, av, dr, di, nn);
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

// ---------------------------------------------------------------------------------------
// SELL-32, batch == 1, TMA-staged: the matrix stream (values + packed columns of a slice
// chunk) is brought into shared memory by 1-D bulk async copies (cp.async.bulk -> SASS UBLKCP)
// completing on per-warp mbarriers, STAGES chunks deep, so the HBM stream stays in flight
// independently of the registers; the threads only read shared memory, gather x (L1/L2) and
// run the fused epilogue.  Each warp runs its own pipeline (lane 0 issues the copies for the
// stage the warp has just drained), so no CTA-wide barrier sits in the loop.
//
// CTA c owns the contiguous slice range [c*spc, (c+1)*spc); warp w takes local slices
// w, w+WARPS, ...  The slice offsets of the range are staged in shared memory once.
// ---------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier, with an L2
// cache-policy hint (the matrix is streamed once per term: evict-first keeps L2 for vectors)
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar,
                                            uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
constexpr int TMA_MAX_LOCAL_SLICES = 2048;  // slice offsets staged per CTA

template <int CH>
struct __align__(128) TmaStage {
  double2 val[CH * QP_SELL_C];
  uint32_t col[CH * QP_SELL_C];
};

template <int EPI, int WARPS, int STAGES, int CH>
__global__ void __launch_bounds__(WARPS * 32, 1)
k_spmv_sell_tma(MatView m, const double2* __restrict__ coef, int n_ops, const double2* __restrict__ x,
                EpiArgs e, int slices_per_cta) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TmaStage<CH>* stages = reinterpret_cast<TmaStage<CH>*>(smem_raw);                       // [WARPS][STAGES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + sizeof(TmaStage<CH>) * WARPS * STAGES);  // [WARPS][STAGES]
  uint32_t* s_ptr = reinterpret_cast<uint32_t*>(bars + WARPS * STAGES);                    // [slices_per_cta + 1]
  __shared__ double2 s_coef[QP_MAX_OPS];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_slices = (m.n + QP_SELL_C - 1) / QP_SELL_C;
  const int64_t s_begin = (int64_t)blockIdx.x * slices_per_cta;
  int64_t s_end = s_begin + slices_per_cta;
  if (s_end > n_slices) s_end = n_slices;
  const int n_loc = s_end > s_begin ? (int)(s_end - s_begin) : 0;
  // the rounded-up slices-per-CTA can leave whole CTAs past the last slice: nothing to do, and
  // their row-pointer reads would lie beyond the array (found by compute-sanitizer memcheck)
  if (s_begin >= n_slices) return;

  if (threadIdx.x < n_ops) s_coef[threadIdx.x] = coef[threadIdx.x];
  for (int i = threadIdx.x; i <= n_loc; i += WARPS * 32) s_ptr[i] = m.ptr[s_begin + i];
  TmaStage<CH>* my_stage = stages + warp * STAGES;
  uint64_t* my_bar = bars + warp * STAGES;
  if (lane == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(my_bar + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (n_loc == 0) return;

  const uint64_t pol_stream = policy_evict_first();
  constexpr uint32_t CHUNK = CH * QP_SELL_C;  // entries per stage

  // producer / consumer cursors (warp-uniform): local slice, entry offset, slice end
  int p_ls = warp, c_ls = warp;
  uint32_t p_off = 0, p_end = 0, c_off = 0, c_end = 0;
  auto load_slice = [&](int& ls, uint32_t& off, uint32_t& end) {
    while (ls < n_loc) {
      off = s_ptr[ls];
      end = s_ptr[ls + 1];
      if (end > off) return;
      ls += WARPS;  // empty slice (all rows empty): nothing to stream, epilogue handled below
    }
  };
  // NOTE: slices whose rows are all empty still need their epilogue (Hx = 0); they are rare
  // (never for a Hamiltonian with a diagonal) and are handled by the consumer loop below via
  // c_ls stepping one slice at a time.
  auto issue = [&](int stage) {
    if (lane == 0) {
      const uint32_t nent = min(CHUNK, p_end - p_off);
      mbar_expect_tx(my_bar + stage, nent * 20u);
      tma_load_1d(my_stage[stage].val, m.val + p_off, nent * 16u, my_bar + stage, pol_stream);
      tma_load_1d(my_stage[stage].col, m.colop + p_off, nent * 4u, my_bar + stage, pol_stream);
    }
  };
  auto p_advance = [&]() {
    p_off += CHUNK;
    if (p_off >= p_end) {
      p_ls += WARPS;
      load_slice(p_ls, p_off, p_end);
    }
  };

  load_slice(p_ls, p_off, p_end);
  // prologue: fill STAGES-1 stages
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (p_ls < n_loc) {
      issue(s);
      p_advance();
    }
  }

  int stage = 0;
  uint32_t parity = 0;
  double dr = 0, di = 0, nn = 0;
  for (; c_ls < n_loc; c_ls += WARPS) {
    c_off = s_ptr[c_ls];
    c_end = s_ptr[c_ls + 1];
    const int64_t row = (s_begin + c_ls) * QP_SELL_C + lane;
    const bool live = row < m.n;
    // epilogue operands are requested now and consumed after the slice: latency fully hidden
    double2 xr, yv, av;
    if (live) epi_load<EPI>(e, x, row, row, xr, yv, av);
    double sr = 0.0, si = 0.0;
    while (c_off < c_end) {
      // refill the stage drained in the previous iteration with the chunk STAGES-1 ahead
      if (p_ls < n_loc) {
        issue((stage + STAGES - 1) % STAGES);
        p_advance();
      }
      mbar_wait(my_bar + stage, parity);
      const uint32_t nj = min(CHUNK, c_end - c_off) / QP_SELL_C;
      const TmaStage<CH>& sb = my_stage[stage];
      if (nj == CH) {
        uint32_t co[CH];
        double2 xv[CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) co[j] = sb.col[j * QP_SELL_C + lane];
#pragma unroll
        for (int j = 0; j < CH; ++j) xv[j] = vldg(x + (co[j] & QP_COL_MASK), e.vec_hint);
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          const double2 v = sb.val[j * QP_SELL_C + lane];
          const double2 u = s_coef[co[j] >> QP_COL_BITS];
          const double tr = v.x * xv[j].x - v.y * xv[j].y;
          const double ti = v.x * xv[j].y + v.y * xv[j].x;
          sr += u.x * tr - u.y * ti;
          si += u.x * ti + u.y * tr;
        }
      } else {
        for (uint32_t j = 0; j < nj; ++j) {
          const uint32_t co = sb.col[j * QP_SELL_C + lane];
          const double2 v = sb.val[j * QP_SELL_C + lane];
          const double2 xv = __ldg(x + (co & QP_COL_MASK));
          const double2 u = s_coef[co >> QP_COL_BITS];
          const double tr = v.x * xv.x - v.y * xv.y;
          const double ti = v.x * xv.y + v.y * xv.x;
          sr += u.x * tr - u.y * ti;
          si += u.x * ti + u.y * tr;
        }
      }
      __syncwarp();  // every lane is done with this stage before lane 0 refills it
      c_off += CHUNK;
      stage = (stage + 1 == STAGES) ? 0 : stage + 1;
      if (stage == 0) parity ^= 1u;
    }
    if (live) epi_apply<EPI>(e, row, make_double2(sr, si), xr, yv, av, dr, di, nn);
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

// ---------------------------------------------------------------------------------------
// SELL-D (dictionary-compressed SELL-32), batch == 1: thread per row, warp per slice.
// The matrix stream shrinks from 20 B to 1-2 B per entry, so the kernel is bounded by the
// five vector streams (80 B/row) and the x gathers (L1/L2), not by the matrix.  The table,
// pre-multiplied by this step's operator coefficients (u_l * value), lives in shared memory:
// per entry one code extract, two shared-memory lookups (offset, value), one gather, 4 DFMA.
// CTA c owns a contiguous slice range (keeps the near gathers of neighbouring slices in L1).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// streaming vector accesses of the SELL-D kernel: v_{k-1}[row], psi[row] are touched exactly once
// per launch, so they must not displace the x lines that neighbouring slices gather from L1
__device__ __forceinline__ double2 ld_noalloc(const double2* p) {
  double2 r;
  asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}

// NG groups of 8 codes taken from the 32-bit words w[]: all 8*NG gathers are issued before the
// first one is consumed
template <int CB, int NC, int REALT>
__device__ __forceinline__ void selld_groups(const uint32_t* w, const double2* s_val, const int32_t* s_delta,
                                             const double2* __restrict__ xbase, double& sr, double& si,
                                             double& sr2, double& si2) {
  uint32_t code[NC];
  double2 xv[NC];
#pragma unroll
  for (int t = 0; t < NC; ++t)
    code[t] = CB == 1 ? (w[t >> 2] >> (8 * (t & 3))) & 0xffu : (w[t >> 1] >> (16 * (t & 1))) & 0xffffu;
#pragma unroll
  for (int t = 0; t < NC; ++t) xv[t] = ld_x(xbase + s_delta[code[t]]);
  if (REALT) {  // every (coefficient x value) of this step is real: 8-byte lookups, 2 DFMA per entry
#pragma unroll
    for (int t = 0; t < NC; ++t) {
      const double v = s_val[code[t]].x;
      if (t & 1) {
        sr2 = fma(v, xv[t].x, sr2);
        si2 = fma(v, xv[t].y, si2);
      } else {
        sr = fma(v, xv[t].x, sr);
        si = fma(v, xv[t].y, si);
      }
    }
    return;
  }
#pragma unroll
  for (int t = 0; t < NC; ++t) {
    // four fused multiply-adds per entry in two accumulator pairs: the plain expression compiles to
    // DMUL + DFMA + DADD per component (6 FP64 instructions), and this loop is sensitive to every
    // instruction (31.8 -> 31.2 us per term on config 2; profiles/r1_selld_split_experiment.md)
    const double2 v = s_val[code[t]];
    sr = fma(v.x, xv[t].x, sr);
    si = fma(v.x, xv[t].y, si);
    sr2 = fma(-v.y, xv[t].y, sr2);
    si2 = fma(v.y, xv[t].x, si2);
  }
}

// one 16-byte word of codes: 8 gathers in flight at a time (16 at once was slower on B200:
// the extra registers cost more latency hiding than the deeper queue buys)
template <int CB, int REALT>
__device__ __forceinline__ void selld_word(const uint4& c, const double2* s_val, const int32_t* s_delta,
                                           const double2* __restrict__ xbase, double& sr, double& si,
                                           double& sr2, double& si2) {
  const uint32_t w[4] = {c.x, c.y, c.z, c.w};
  if (CB == 1) {
    selld_groups<1, 8, REALT>(w, s_val, s_delta, xbase, sr, si, sr2, si2);
    if ((w[2] | w[3]) != 0u) selld_groups<1, 8, REALT>(w + 2, s_val, s_delta, xbase, sr, si, sr2, si2);  // not all padding
  } else {
    selld_groups<2, 8, REALT>(w, s_val, s_delta, xbase, sr, si, sr2, si2);
  }
}

// the LAST word of a row when the longest row of the matrix leaves TAIL (< codes per word) codes
// in it: the padding behind them is never decoded (no run-time test: TAIL is a property of the
// matrix, the last word is where the word loop ends anyway)
template <int CB, int TAIL, int REALT>
__device__ __forceinline__ void selld_word_tail(const uint4& c, const double2* s_val, const int32_t* s_delta,
                                                const double2* __restrict__ xbase, double& sr, double& si,
                                                double& sr2, double& si2) {
  const uint32_t w[4] = {c.x, c.y, c.z, c.w};
  if (CB == 1) {
    if (TAIL <= 8) {
      selld_groups<1, (TAIL <= 8 ? TAIL : 8), REALT>(w, s_val, s_delta, xbase, sr, si, sr2, si2);
    } else {
      selld_groups<1, 8, REALT>(w, s_val, s_delta, xbase, sr, si, sr2, si2);
      if ((w[2] | w[3]) != 0u) selld_groups<1, (TAIL > 8 ? TAIL - 8 : 8), REALT>(w + 2, s_val, s_delta, xbase, sr, si, sr2, si2);
    }
  } else {
    selld_groups<2, (TAIL < 8 ? TAIL : 8), REALT>(w, s_val, s_delta, xbase, sr, si, sr2, si2);
  }
}

template <int EPI>
__device__ __forceinline__ void selld_epi_load(const EpiArgs& e, const double2* __restrict__ x, int64_t row,
                                               double2& xr, double2& yv, double2& av) {
  xr = yv = av = make_double2(0.0, 0.0);
  if (EPI == EPI_MUL) {
    if (e.betac.x != 0.0 || e.betac.y != 0.0) yv = ld_noalloc(e.y + row);
    return;
  }
  xr = ld_x(x + row);  // allocates in L1: neighbouring rows gather it
  if (EPI == EPI_CHEB_MID || EPI == EPI_CHEB_LAST) {
    yv = ld_noalloc(e.y + row);
    av = ld_noalloc(e.acc + row);
  }
}

#ifndef SELLD_THREADS
#define SELLD_THREADS 512
#endif
#ifndef SELLD_LEAN
#define SELLD_LEAN 0
#endif
template <int EPI, int CB, int TAIL, int REALT>
__global__ void __launch_bounds__(SELLD_THREADS, 1)
k_spmv_selld(DictView m, const double2* __restrict__ coef, const double2* __restrict__ x, EpiArgs e,
             int slices_per_cta) {
  // the table, pre-multiplied by this step's operator coefficients, lives in shared memory
  // (passing it as a kernel parameter and reading it through the constant cache was 33 %
  // slower: per-lane indexed LDC serialises on distinct addresses)
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double2* s_val = reinterpret_cast<double2*>(smem_raw);
  int32_t* s_delta = reinterpret_cast<int32_t*>(s_val + m.n_dict);
  // programmatic dependent launch: the launch of the next term and this table set-up (which
  // depends on the matrix and the coefficients only) overlap the tail of the previous term;
  // nothing written by the previous term (x, y, acc) is touched before griddepcontrol.wait
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  for (int j = threadIdx.x; j < m.n_dict; j += blockDim.x) {
    s_val[j] = cmul2(coef[m.dop[j]], ld_stream(m.dval + j));  // static data: may sit above the wait
    s_delta[j] = m.ddelta[j];
  }
  __syncthreads();
  asm volatile("griddepcontrol.wait;" ::: "memory");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int64_t n_slices = (m.n + QP_SELL_C - 1) / QP_SELL_C;
  const int64_t s_begin = (int64_t)blockIdx.x * slices_per_cta;
  int64_t s_end = s_begin + slices_per_cta;
  if (s_end > n_slices) s_end = n_slices;
  double dr = 0, di = 0, nn = 0;

  // software pipeline over the warp's slices: the first code word and the epilogue operands of
  // the NEXT slice are requested before the current slice's gathers are consumed
  uint32_t n_off0 = 0, n_off1 = 0;
  uint4 n_c = make_uint4(0u, 0u, 0u, 0u);
  double2 n_xr, n_yv, n_av;
  n_xr = n_yv = n_av = make_double2(0.0, 0.0);
  auto prefetch = [&](int64_t s) {
    if (m.uniform_words) {
      n_off0 = (uint32_t)s * m.uniform_words;
      n_off1 = n_off0 + m.uniform_words;
    } else {
      n_off0 = m.sptr[s];
      n_off1 = m.sptr[s + 1];
    }
    if (n_off0 < n_off1) n_c = ld_stream(m.codes + n_off0 + lane);
    const int64_t row = s * QP_SELL_C + lane;
    if (row < m.n) selld_epi_load<EPI>(e, x, row, n_xr, n_yv, n_av);
  };
  int64_t s = s_begin + warp;
  if (!SELLD_LEAN && s < s_end) prefetch(s);
  while (s < s_end) {
    if (SELLD_LEAN) prefetch(s);  // nothing carried from slice to slice: fewer registers, more warps
    const int64_t row = s * QP_SELL_C + lane;
    const bool live = row < m.n;
    const double2* xbase = x + (live ? row : m.n - 1);  // dead lanes hold padding codes only
    uint32_t off = n_off0 + lane;
    const uint32_t off1 = n_off1;
    uint4 c = n_c;
    const double2 xr = n_xr, yv = n_yv, av = n_av;
    const bool any = n_off0 < n_off1;
    const int64_t s_next = s + nwarps;
    if (!SELLD_LEAN && s_next < s_end) prefetch(s_next);
    double sr = 0.0, si = 0.0, sr2 = 0.0, si2 = 0.0;
    if (any) {
      for (;;) {
        const uint32_t off_n = off + QP_SELL_C;
        const bool more = off_n < off1;
        uint4 c_n = c;
        if (more) c_n = ld_stream(m.codes + off_n);  // look one word ahead
        if (TAIL > 0 && !more) {
          selld_word_tail<CB, (TAIL > 0 ? TAIL : 8), REALT>(c, s_val, s_delta, xbase, sr, si, sr2, si2);
          break;
        }
        selld_word<CB, REALT>(c, s_val, s_delta, xbase, sr, si, sr2, si2);
        if (!more) break;
        c = c_n;
        off = off_n;
      }
    }
    sr += sr2;
    si += si2;
    if (m.n_diag > 0 && live) {  // explicit diagonals: hx += sum_i u_i d_i[row] x[row]
      const double2 xs = EPI == EPI_MUL ? ld_x(x + row) : xr;
      double2 d = make_double2(0.0, 0.0);
      for (int i = 0; i < m.n_diag; ++i) {
        const double2 t = cmul2(coef[(m.diag_ops >> (4 * i)) & 15ull], ld_stream(m.diag + (int64_t)i * m.n + row));
        d.x += t.x;
        d.y += t.y;
      }
      sr += d.x * xs.x - d.y * xs.y;
      si += d.x * xs.y + d.y * xs.x;
    }
    if (live) epi_apply<EPI>(e, row, make_double2(sr, si), xr, yv, av, dr, di, nn);
    s = s_next;
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
  double dr = 0, di = 0, nn = 0;
  epilogue<EPI>(e, x, idx, idx, make_double2(sr, si), dr, di, nn);
  if (epi_has_sums(EPI) && e.chk != nullptr) chk_flush(e, b, dr, di, nn);
}

// ---------------------------------------------------------------------------------------
// SELL-D, trajectory-batched: warp per (row, T chunks of 32 trajectories); state layout [N][B].
// Every lane of a warp works on the same matrix row, so the code stream, the table lookups and
// the operator changes are warp-uniform and decoded once for T trajectories per lane; the
// gathers are 512 B contiguous.  The per-trajectory coefficient u_l^(b) is applied once per
// operator and row (entries of a row are stored operator by operator).  REALV: every operator
// is purely real or purely imaginary (the usual case: ladder operators, Pauli sums), so an
// entry is one real number (2 DFMA per trajectory instead of 4) and the factor i of an
// imaginary operator is folded into its coefficient.
// The grid runs all row blocks of one trajectory chunk before the next chunk (blockIdx.x = row
// block), so the chunk's slice of x (N x 32T x 16 B) stays in L2 while the rows sweep over it:
// every gather after the first is an L1/L2 hit and HBM traffic stays at the algorithmic 80 B
// per (row, trajectory).  CTA = one slice of 32 rows; warp w takes rows w, w+8, w+16, w+24.
// ---------------------------------------------------------------------------------------
struct DeltaOp {
  int32_t delta;
  int32_t op;
};

template <int EPI, int CB, int REALV, int T, int G, int LATE>
__global__ void __launch_bounds__(256, 2)
k_spmm_selld(DictView m, const double* __restrict__ dvalr, unsigned imag_ops, const double2* __restrict__ coef,
             int coef_stride, int64_t batch, const double2* __restrict__ x, EpiArgs e, int slices_per_cta) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double2* s_val2 = reinterpret_cast<double2*>(smem_raw);  // !REALV
  double* s_val1 = reinterpret_cast<double*>(smem_raw);    // REALV
  DeltaOp* s_dop = reinterpret_cast<DeltaOp*>(smem_raw + (size_t)m.n_dict * (REALV ? 8 : 16));
  for (int j = threadIdx.x; j < m.n_dict; j += blockDim.x) {
    if (REALV) s_val1[j] = dvalr[j];
    else s_val2[j] = m.dval[j];
    s_dop[j] = DeltaOp{m.ddelta[j], (int32_t)m.dop[j]};
  }
  __syncthreads();

  constexpr int CPW = 16 / CB;   // codes per 16-byte word
  static_assert(CPW % G == 0, "group size must divide the codes per word");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_slices = (m.n + QP_SELL_C - 1) / QP_SELL_C;
  int32_t bcol[T];   // this lane's trajectories (clamped; dead ones are computed on a copy)
  bool blive[T];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const int64_t b = ((int64_t)blockIdx.y * T + t) * 32 + lane;
    blive[t] = b < batch;
    bcol[t] = (int32_t)(blive[t] ? b : batch - 1);
  }
  const int64_t ustride = coef_stride ? batch : 1;
  auto ucoef = [&](int op, int t) {  // u_op of trajectory t (times i for an imaginary operator)
    const double2 u = __ldg(coef + (int64_t)op * ustride + (coef_stride ? bcol[t] : 0));
    return (REALV && ((imag_ops >> op) & 1u)) ? make_double2(-u.y, u.x) : u;
  };

  double dr[T], di[T], nn[T];
#pragma unroll
  for (int t = 0; t < T; ++t) dr[t] = di[t] = nn[t] = 0.0;
  // a CTA sweeps `slices_per_cta` consecutive slices: the gathers of nearby rows (offsets up to
  // the tile height) hit lines this CTA fetched a moment ago
  const int64_t slice_end = min((int64_t)(blockIdx.x + 1) * slices_per_cta, n_slices);
  for (int64_t slice = (int64_t)blockIdx.x * slices_per_cta; slice < slice_end; ++slice) {
  uint32_t off0, off1;
  if (m.uniform_words) {
    off0 = (uint32_t)slice * m.uniform_words;
    off1 = off0 + m.uniform_words;
  } else {
    off0 = m.sptr[slice];
    off1 = m.sptr[slice + 1];
  }
  for (int rl = warp; rl < QP_SELL_C; rl += 8) {
    const int64_t row = slice * QP_SELL_C + rl;
    if (row >= m.n) break;
    double2 xr[T], yv[T], av[T];
    const double2* xb0 = x + row * batch;  // x[(row + delta) * batch + b] = xb0[delta * batch + b]
    double tr[T], ti[T], pr[T], pi[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int64_t idx = row * batch + bcol[t];
      // LATE (many trajectories per lane): the epilogue operands are loaded after the row instead
      // of being held in registers across it
      if (!LATE) epi_load<EPI>(e, x, idx, idx, xr[t], yv[t], av[t]);
      tr[t] = ti[t] = pr[t] = pi[t] = 0.0;
    }
    int cur_op = -1;

    // the row's code words form a dependent chain of global loads (word -> table -> gathers):
    // the next word is requested before the current one is decoded
    uint4 c_next = make_uint4(0u, 0u, 0u, 0u);
    if (off0 + rl < off1) c_next = __ldg(m.codes + off0 + rl);  // same word for all lanes: one broadcast load
    for (uint32_t off = off0 + rl; off < off1; off += QP_SELL_C) {
      const uint4 c = c_next;
      if (off + QP_SELL_C < off1) c_next = __ldg(m.codes + off + QP_SELL_C);
      const uint32_t w[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
      for (int h = 0; h < CPW; h += G) {  // G codes x T trajectories = 8 gathers in flight per lane
        uint32_t code[G];
        DeltaOp dop[G];
        double2 xv[G][T];
        bool any = false;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const int tt = h + g;
          code[g] = CB == 1 ? (w[tt >> 2] >> (8 * (tt & 3))) & 0xffu : (w[tt >> 1] >> (16 * (tt & 1))) & 0xffffu;
          any |= code[g] != 0u;
        }
        if (!any) continue;  // all padding (warp-uniform)
#pragma unroll
        for (int g = 0; g < G; ++g) {
          dop[g] = s_dop[code[g]];
          const int64_t o = (int64_t)dop[g].delta * batch;
#pragma unroll
          for (int t = 0; t < T; ++t) xv[g][t] = __ldg(xb0 + o + bcol[t]);
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
          if (code[g] == 0u) continue;  // padding / entry kept in an explicit diagonal
          if (dop[g].op != cur_op) {    // warp-uniform: fold the finished operator
            if (cur_op >= 0) {
#pragma unroll
              for (int t = 0; t < T; ++t) {
                const double2 u = ucoef(cur_op, t);
                tr[t] += u.x * pr[t] - u.y * pi[t];
                ti[t] += u.x * pi[t] + u.y * pr[t];
                pr[t] = pi[t] = 0.0;
              }
            }
            cur_op = dop[g].op;
          }
          if (REALV) {
            const double v = s_val1[code[g]];
#pragma unroll
            for (int t = 0; t < T; ++t) {
              pr[t] += v * xv[g][t].x;
              pi[t] += v * xv[g][t].y;
            }
          } else {
            const double2 v = s_val2[code[g]];
#pragma unroll
            for (int t = 0; t < T; ++t) {
              pr[t] += v.x * xv[g][t].x - v.y * xv[g][t].y;
              pi[t] += v.x * xv[g][t].y + v.y * xv[g][t].x;
            }
          }
        }
      }
    }
    if (cur_op >= 0) {
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const double2 u = ucoef(cur_op, t);
        tr[t] += u.x * pr[t] - u.y * pi[t];
        ti[t] += u.x * pi[t] + u.y * pr[t];
      }
    }
    if (LATE) {
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const int64_t idx = row * batch + bcol[t];
        epi_load<EPI>(e, x, idx, idx, xr[t], yv[t], av[t]);
      }
    }
    if (m.n_diag > 0) {  // explicit diagonals (warp-uniform matrix element, per-trajectory u)
      for (int i = 0; i < m.n_diag; ++i) {
        const int op = (int)((m.diag_ops >> (4 * i)) & 15ull);
        const double2 d = __ldg(m.diag + (int64_t)i * m.n + row);
#pragma unroll
        for (int t = 0; t < T; ++t) {
          const double2 u = __ldg(coef + (int64_t)op * ustride + (coef_stride ? bcol[t] : 0));
          const double2 ud = cmul2(u, d);
          const double2 xs = EPI == EPI_MUL ? __ldg(xb0 + bcol[t]) : xr[t];
          tr[t] += ud.x * xs.x - ud.y * xs.y;
          ti[t] += ud.x * xs.y + ud.y * xs.x;
        }
      }
    }
#pragma unroll
    for (int t = 0; t < T; ++t)
      if (blive[t]) {
        if (EPI == EPI_DOT) {  // sums carried across the CTA's rows
          epi_apply<EPI>(e, row * batch + bcol[t], make_double2(tr[t], ti[t]), xr[t], yv[t], av[t], dr[t], di[t], nn[t]);
        } else {  // normalization check (debug option): flushed per row, nothing held across rows
          double cr = 0, ci = 0, cn = 0;
          epi_apply<EPI>(e, row * batch + bcol[t], make_double2(tr[t], ti[t]), xr[t], yv[t], av[t], cr, ci, cn);
          if (epi_has_sums(EPI) && e.chk != nullptr) chk_flush(e, bcol[t], cr, ci, cn);
        }
      }
  }
  }
  if (EPI == EPI_DOT && e.chk != nullptr) {
#pragma unroll
    for (int t = 0; t < T; ++t)
      if (blive[t]) chk_flush(e, bcol[t], dr[t], di[t], nn[t]);
  }
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
  if (row < n && lane == 0) epilogue<EPI>(e, x, row, row, make_double2(sr, si), dr, di, nn);
  if (epi_has_sums(EPI) && e.chk != nullptr) {
    for (int o = 16; o > 0; o >>= 1) {
      dr += __shfl_xor_sync(0xffffffffu, dr, o);
      di += __shfl_xor_sync(0xffffffffu, di, o);
      nn += __shfl_xor_sync(0xffffffffu, nn, o);
    }
    if (lane == 0) chk_flush_warp(e, dr, di, nn);
  }
}

// The same for large n: a CTA of 8 warps takes one row at a time (grid-stride over the rows), its 256 lanes
// stream the row in 4 KB steps and the partial sums meet in shared memory.  One warp per row (above) keeps
// n concurrent row streams open -- 8192 on config 5, each advancing 512 B at a time -- and DRAM then runs at
// half its rate; here the concurrent streams are the resident CTAs (~1000), each reading 4 KB contiguously.
template <int EPI>
__global__ void __launch_bounds__(256)
k_gemv_dense_cta(const double2* const* __restrict__ ops, int n_ops, int64_t n,
                 const double2* __restrict__ coef, const double2* __restrict__ x, EpiArgs e) {
  __shared__ double2 s_coef[QP_MAX_OPS];
  __shared__ double s_part[8][2];
  if (threadIdx.x < n_ops) s_coef[threadIdx.x] = coef[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double dr = 0, di = 0, nn = 0;
  for (int64_t row = blockIdx.x; row < n; row += gridDim.x) {
    double sr = 0.0, si = 0.0;
    for (int l = 0; l < n_ops; ++l) {
      const double2* __restrict__ a = ops[l] + row * n;
      const double2 u = s_coef[l];
      double pr = 0.0, pi = 0.0;
#pragma unroll 4
      for (int64_t j = threadIdx.x; j < n; j += 256) {
        const double2 v = ld_stream(a + j);
        const double2 xv = __ldg(x + j);
        pr += v.x * xv.x - v.y * xv.y;
        pi += v.x * xv.y + v.y * xv.x;
      }
      sr += u.x * pr - u.y * pi;
      si += u.x * pi + u.y * pr;
    }
    for (int o = 16; o > 0; o >>= 1) {
      sr += __shfl_xor_sync(0xffffffffu, sr, o);
      si += __shfl_xor_sync(0xffffffffu, si, o);
    }
    if (lane == 0) {
      s_part[warp][0] = sr;
      s_part[warp][1] = si;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double tr = 0.0, ti = 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {  // fixed order
        tr += s_part[k][0];
        ti += s_part[k][1];
      }
      epilogue<EPI>(e, x, row, row, make_double2(tr, ti), dr, di, nn);
    }
    __syncthreads();
  }
  if (epi_has_sums(EPI) && e.chk != nullptr && warp == 0) {  // only thread 0 holds sums
    for (int o = 16; o > 0; o >>= 1) {
      dr += __shfl_xor_sync(0xffffffffu, dr, o);
      di += __shfl_xor_sync(0xffffffffu, di, o);
      nn += __shfl_xor_sync(0xffffffffu, nn, o);
    }
    if (lane == 0) chk_flush_warp(e, dr, di, nn);
  }
}

// Flat two-pass GEMV for large dense operators (n a multiple of 128): the matrix is read as ONE contiguous stream
// -- warp w of the grid takes the 2 KB pieces w, w + W, w + 2W, ... of the row-major array, so that at any moment
// the grid reads one contiguous window (like the Krylov streaming kernels, which reach 6.2-6.3 TB/s), instead of
// thousands of row streams 128 KB apart (0.64-0.67 of the roofline: DRAM page conflicts) -- and writes one partial
// dot product per piece; the second kernel adds the n / 128 partials of a row in a fixed order and applies the
// fused epilogue.  part[(l * n + row) * ppr + k].
template <int UNUSED>  // a template only because this header is included by several translation units
__global__ void __launch_bounds__(256)
k_gemv_flat(const double2* const* __restrict__ ops, int n_ops, int64_t n, const double2* __restrict__ x,
            double2* __restrict__ part) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t ppr = n >> 7;  // pieces of 128 elements per row
  const int64_t total = (int64_t)n_ops * n * ppr;
  for (int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < total; p += warps) {
    const int64_t lr = p / ppr, k = p - lr * ppr;  // (operator, row) and piece within the row
    const int l = (int)(lr / n);
    const int64_t row = lr - (int64_t)l * n;
    const double2* __restrict__ a = ops[l] + row * n + (k << 7) + lane;
    const double2* __restrict__ xs = x + (k << 7) + lane;
    double2 v[4], xv[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = ld_stream(a + 32 * q);
#pragma unroll
    for (int q = 0; q < 4; ++q) xv[q] = __ldg(xs + 32 * q);
    double pr = 0.0, pi = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      pr += v[q].x * xv[q].x - v[q].y * xv[q].y;
      pi += v[q].x * xv[q].y + v[q].y * xv[q].x;
    }
    for (int o = 16; o > 0; o >>= 1) {
      pr += __shfl_xor_sync(0xffffffffu, pr, o);
      pi += __shfl_xor_sync(0xffffffffu, pi, o);
    }
    if (lane == 0) part[p] = make_double2(pr, pi);
  }
}

template <int EPI>
__global__ void __launch_bounds__(256)
k_gemv_flat_fin(const double2* __restrict__ part, int n_ops, int64_t n, const double2* __restrict__ coef,
                const double2* __restrict__ x, EpiArgs e) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t ppr = n >> 7;
  double dr = 0, di = 0, nn = 0;
  if (row < n) {
    double sr = 0.0, si = 0.0;
    for (int l = 0; l < n_ops; ++l) {
      const double2* __restrict__ pp = part + ((int64_t)l * n + row) * ppr;
      double pr = 0.0, pi = 0.0;
      for (int64_t k = 0; k < ppr; ++k) {
        const double2 t = pp[k];
        pr += t.x;
        pi += t.y;
      }
      const double2 u = coef[l];
      sr += u.x * pr - u.y * pi;
      si += u.x * pi + u.y * pr;
    }
    epilogue<EPI>(e, x, row, row, make_double2(sr, si), dr, di, nn);
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

// two-pass tiled path for batched states (tile.cu); *handled = false: use the one-pass kernels
// bit-flip (XOR-stencil) form (bitflip.cu): *ok = false if the generator does not have the structure
int32_t qp_bitflip_build(qp_gen_t gen, bool* ok);
int64_t qp_bitflip_stored_bytes(const qp_bitflip_s* b);
int32_t qp_launch_bitflip(qp_gen_t gen, int epi, const double2* x, const EpiArgs& e);
int32_t qp_launch_tile(qp_gen_t gen, int epi, int coef_stride, const double2* x, int64_t batch, const EpiArgs& e,
                       bool* handled);
int32_t qp_tile_info(qp_gen_t gen, int32_t* available, int32_t* S, int32_t* NH, int32_t* n_table, int64_t* entries);

// host-side dispatcher (defined in sparse.cu)
int32_t qp_launch_fused(qp_gen_t gen, int epi, int coef_stride, const double2* x, int64_t batch,
                        const EpiArgs& e);
