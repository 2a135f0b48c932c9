// Dense generators applied to a batch of states on the FP64 tensor cores (DMMA).
//
//   Y[N][B] = sum_l u_l H_l X[N][B]        H_l dense row-major ComplexF64, X batch-fastest
//
// A complex product is a real one with the interleaved (re, im) row of H_l as the K
// dimension:  [Yr Yi] = [Hr Hi] . [[Xr, Xi], [-Xi, Xr]]  -- so the A operand is the operator
// exactly as it lies in memory (K = 2N reals) and the B operand is built on the fly from the
// staged X tile; a thread's accumulator pair is (re, im) of one (row, trajectory), which is
// what the fused Chebyshev epilogue consumes.  tcgen05 has no FP64 kind, so this is
// mma.sync.m8n8k4.f64 (SASS DMMA); the K order inside a group of 8 reals is permuted
// (all real parts, then all imaginary parts) so that both operands come from shared memory
// with 128-bit loads.
//
// CTA tile: (WM * MT * 8) rows x (WN * 8) trajectories; warp tile: MT m8-tiles x 2 n8-tiles
// (8 trajectories); K tile: 16 complex per stage, STAGES-deep cp.async pipeline running
// seamlessly across the operators of the lazy sum.  The coefficient u_l (per trajectory in
// ensemble mode) is applied to the accumulators once per operator, NOT to the fragments: the
// DMMA runs on the FP64 pipe, and any DFMA/DMUL between two DMMAs drains it (measured: 69 % ->
// see profiles/), so the K loop contains only LDS, integer selects and DMMA.
// KS warp groups split every K tile between them (same output tile, partial sums added through
// shared memory at the end): the DMMA pipe needs ~4 warps per scheduler to stay busy, and the
// row count of the operator (8192 / 148 SMs = 56 rows per CTA) leaves no other way to get them.
#pragma once

#include "spmv.cuh"

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem)), "l"(gmem), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

constexpr int DM_KC = 16;       // complex K elements per stage
constexpr int DM_STAGES = 3;
constexpr int DM_A_STRIDE = 40;  // doubles per staged A row (32 + 8: rows 64 B apart mod 128 B)

template <int WM, int WN, int MT, int NQ>
struct DenseTile {
  static constexpr int ROWS = WM * MT * 8;
  static constexpr int TRAJ = WN * NQ * 4;
  static constexpr int X_STRIDE = TRAJ + 1;  // complex per staged X row (rows 16 B apart mod 128 B)
  static constexpr size_t A_BYTES = (size_t)ROWS * DM_A_STRIDE * sizeof(double);
  static constexpr size_t X_BYTES = (size_t)DM_KC * X_STRIDE * sizeof(double2);
  static constexpr size_t STAGE_BYTES = (A_BYTES + X_BYTES + 127) / 128 * 128;
  static constexpr size_t SMEM = STAGE_BYTES * DM_STAGES + sizeof(double2) * QP_MAX_OPS * TRAJ;
};

template <int EPI, int WM, int WN, int MT, int NQ, int KS>
__global__ void __launch_bounds__(WM * WN * KS * 32, 1)
k_gemm_dense(const double2* const* __restrict__ ops, int n_ops, int64_t n, const double2* __restrict__ coef,
             int coef_stride, int64_t batch, const double2* __restrict__ x, EpiArgs e) {
  using T = DenseTile<WM, WN, MT, NQ>;
  constexpr int THREADS = WM * WN * KS * 32;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double2* s_coef = reinterpret_cast<double2*>(smem_raw + T::STAGE_BYTES * DM_STAGES);  // [n_ops][TRAJ]

  const int tid = threadIdx.x, lane = tid & 31;
  const int ks = (tid >> 5) / (WM * WN), warp = (tid >> 5) % (WM * WN);  // K group, warp within the tile
  const int wm = warp / WN, wn = warp % WN;
  const int g = lane >> 2, t = lane & 3;
  const int64_t row0 = (int64_t)blockIdx.x * T::ROWS;
  const int64_t b0 = (int64_t)blockIdx.y * T::TRAJ;

  for (int i = tid; i < n_ops * T::TRAJ; i += THREADS) {
    const int l = i / T::TRAJ, b = i % T::TRAJ;
    const int64_t bb = b0 + b < batch ? b0 + b : batch - 1;
    s_coef[i] = coef[(int64_t)l * (coef_stride ? batch : 1) + (coef_stride ? bb : 0)];
  }

  const int KT = (int)((n + DM_KC - 1) / DM_KC);
  const int total_it = n_ops * KT;

  auto stage_A = [&](int s) { return reinterpret_cast<double*>(smem_raw + T::STAGE_BYTES * s); };
  auto stage_X = [&](int s) { return reinterpret_cast<double2*>(smem_raw + T::STAGE_BYTES * s + T::A_BYTES); };

  auto load_stage = [&](int s, int it) {
    if (it < total_it) {
      const int l = it / KT, kt = it - l * KT;
      const int64_t kc0 = (int64_t)kt * DM_KC;
      const double2* __restrict__ A = ops[l];
      double* As = stage_A(s);
      for (int c = tid; c < T::ROWS * DM_KC; c += THREADS) {
        const int rr = c / DM_KC, kk = c % DM_KC;
        int64_t r = row0 + rr;
        if (r >= n) r = n - 1;  // rows past the end are computed on a copy and never stored
        const int64_t kc = kc0 + kk;
        const bool ok = kc < n;
        cp_async16(As + rr * DM_A_STRIDE + 2 * kk, A + r * n + (ok ? kc : 0), ok ? 16 : 0);
      }
      double2* Xs = stage_X(s);
      for (int c = tid; c < DM_KC * T::TRAJ; c += THREADS) {
        const int kk = c / T::TRAJ, b = c % T::TRAJ;
        const int64_t kc = kc0 + kk, bb = b0 + b;
        const bool ok = kc < n && bb < batch;
        cp_async16(Xs + kk * T::X_STRIDE + b, x + (ok ? kc * batch + bb : 0), ok ? 16 : 0);
      }
    }
    cp_async_commit();
  };

  double acc[MT][NQ][2], tot[MT][NQ][2];  // current operator / sum over finished operators
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int q = 0; q < NQ; ++q) acc[i][q][0] = acc[i][q][1] = tot[i][q][0] = tot[i][q][1] = 0.0;
  // tot += u_l * acc for the operator that just finished (C-side trajectory of n8-tile q: 4q + t)
  auto fold = [&](int l) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const double2 u = s_coef[l * T::TRAJ + wn * NQ * 4 + q * 4 + t];
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        tot[i][q][0] += u.x * acc[i][q][0] - u.y * acc[i][q][1];
        tot[i][q][1] += u.x * acc[i][q][1] + u.y * acc[i][q][0];
        acc[i][q][0] = acc[i][q][1] = 0.0;
      }
    }
  };

#pragma unroll
  for (int s = 0; s < DM_STAGES - 1; ++s) load_stage(s, s);

  const int cpar = g & 1;              // this lane's B column is the (re | im) part of its trajectory
  const int bt = wn * NQ * 4 + (g >> 1);  // B-side trajectory of n8-tile 0 (tile q: +4q)
  int cur_op = 0;

  for (int it = 0; it < total_it; ++it) {
    cp_async_wait<DM_STAGES - 2>();
    __syncthreads();  // stage `it` has landed for everyone; stage it-1 is drained by everyone
    load_stage((it + DM_STAGES - 1) % DM_STAGES, it + DM_STAGES - 1);

    const int l = it / KT;
    if (l != cur_op) {  // operator boundary (n_ops - 1 times per kernel)
      fold(cur_op);
      cur_op = l;
    }
    const double* As = stage_A(it % DM_STAGES) + (wm * MT * 8 + g) * DM_A_STRIDE + 2 * t;
    const double2* Xs = stage_X(it % DM_STAGES) + t * T::X_STRIDE + bt;
#pragma unroll
    for (int jp = ks; jp < DM_KC / 4; jp += KS) {
      double2 a[MT];
#pragma unroll
      for (int i = 0; i < MT; ++i)
        a[i] = *reinterpret_cast<const double2*>(As + i * 8 * DM_A_STRIDE + 8 * jp);
      double be[NQ], bo[NQ];
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const double2 z = Xs[jp * 4 * T::X_STRIDE + q * 4];
        // sign flip on the bit pattern: no FP64-pipe instruction between the DMMAs
        int hi = __double2hiint(z.y);
        asm volatile("xor.b32 %0, %0, 0x80000000;" : "+r"(hi));  // opaque to the optimiser (else it emits DADD)
        const double nzy = __hiloint2double(hi, __double2loint(z.y));
        be[q] = cpar ? z.y : z.x;  // K = real part of H:  [Xr | Xi]
        bo[q] = cpar ? z.x : nzy;  // K = imag part of H:  [-Xi | Xr]
      }
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int q = 0; q < NQ; ++q) dmma884(acc[i][q][0], acc[i][q][1], a[i].x, be[q]);
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int q = 0; q < NQ; ++q) dmma884(acc[i][q][0], acc[i][q][1], a[i].y, bo[q]);
    }
  }
  cp_async_wait<0>();
  fold(cur_op);
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      acc[i][q][0] = tot[i][q][0];
      acc[i][q][1] = tot[i][q][1];
    }
  if (KS > 1) {  // add the K groups' partial sums through shared memory (the stages are drained)
    double2* red = reinterpret_cast<double2*>(smem_raw);
    for (int k = KS - 1; k >= 1; --k) {
      __syncthreads();
      if (ks == k) {
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
          for (int q = 0; q < NQ; ++q)
            red[((warp * MT + i) * NQ + q) * 32 + lane] = make_double2(acc[i][q][0], acc[i][q][1]);
      }
      __syncthreads();
      if (ks == 0) {
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
          for (int q = 0; q < NQ; ++q) {
            const double2 p = red[((warp * MT + i) * NQ + q) * 32 + lane];
            acc[i][q][0] += p.x;
            acc[i][q][1] += p.y;
          }
      }
    }
    if (ks != 0) return;
  }

  // fused epilogue: accumulator pair = (re, im) of (row, trajectory)
  double dr[NQ], di[NQ], nn[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) dr[q] = di[q] = nn[q] = 0.0;
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    const int64_t row = row0 + (wm * MT + i) * 8 + g;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int64_t b = b0 + wn * NQ * 4 + q * 4 + t;
      if (row < n && b < batch) {
        const int64_t idx = row * batch + b;
        epilogue<EPI>(e, x, idx, idx, make_double2(acc[i][q][0], acc[i][q][1]), dr[q], di[q], nn[q]);
      }
    }
  }
  if (epi_has_sums(EPI) && e.chk != nullptr) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      // lanes with equal t hold the same trajectory: reduce over g
      for (int o = 4; o < 32; o <<= 1) {
        dr[q] += __shfl_xor_sync(0xffffffffu, dr[q], o);
        di[q] += __shfl_xor_sync(0xffffffffu, di[q], o);
        nn[q] += __shfl_xor_sync(0xffffffffu, nn[q], o);
      }
      const int64_t b = b0 + wn * NQ * 4 + q * 4 + t;
      if (g == 0 && b < batch) chk_flush(e, b, dr[q], di[q], nn[q]);
    }
  }
}

template <int EPI, int WM, int WN, int MT, int NQ, int KS>
static int32_t launch_gemm_dense(qp_gen_t gen, int coef_stride, const double2* x, int64_t batch, const EpiArgs& e) {
  using T = DenseTile<WM, WN, MT, NQ>;
  qp_ctx_t ctx = gen->ctx;
  static_assert(sizeof(double2) * WM * WN * MT * NQ * 32 <= T::STAGE_BYTES * DM_STAGES, "reduction buffer");
  auto kern = k_gemm_dense<EPI, WM, WN, MT, NQ, KS>;
  if (!ctx->smem_configured.count((const void*)kern)) {
    QP_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM));
    ctx->smem_configured.insert((const void*)kern);
  }
  dim3 grid((unsigned)((gen->n + T::ROWS - 1) / T::ROWS), (unsigned)((batch + T::TRAJ - 1) / T::TRAJ));
  kern<<<grid, WM * WN * KS * 32, T::SMEM, ctx->stream>>>(gen->d_dense_ops, gen->n_ops, gen->n, gen->d_coef, coef_stride,
                                                      batch, x, e);
  QP_LAUNCHED(ctx);
  return QP_OK;
}

// tile shape by batch width: the n extent of the CTA covers the whole batch when it can, so
// every operator element is read from HBM once per application
template <int EPI>
static int32_t launch_dense_batched(qp_gen_t gen, int coef_stride, const double2* x, int64_t batch, const EpiArgs& e) {
  static const int variant = getenv("QPROP_DMMA_VARIANT") ? atoi(getenv("QPROP_DMMA_VARIANT")) : 0;
  if (batch > 32) {
    if (variant == 1) return launch_gemm_dense<EPI, 1, 16, 7, 1, 1>(gen, coef_stride, x, batch, e);  // 56 x 64, 16 warps, 7 acc/warp
    if (variant == 2) return launch_gemm_dense<EPI, 2, 8, 4, 2, 1>(gen, coef_stride, x, batch, e);   // 64 x 64, 16 warps, 8 acc/warp
    if (variant == 3) return launch_gemm_dense<EPI, 1, 8, 7, 2, 2>(gen, coef_stride, x, batch, e);   // 56 x 64, 2 x 8 warps (K split)
    return launch_gemm_dense<EPI, 1, 8, 7, 2, 1>(gen, coef_stride, x, batch, e);                     // 56 x 64, 8 warps, 14 acc/warp
  }
  if (batch > 16) return launch_gemm_dense<EPI, 2, 4, 4, 2, 2>(gen, coef_stride, x, batch, e);   // 64 rows x 32 traj, 16 warps
  if (batch > 8) return launch_gemm_dense<EPI, 4, 2, 2, 2, 2>(gen, coef_stride, x, batch, e);    // 64 rows x 16 traj, 16 warps
  return launch_gemm_dense<EPI, 8, 1, 1, 2, 2>(gen, coef_stride, x, batch, e);                   // 64 rows x 8 traj, 16 warps
}
