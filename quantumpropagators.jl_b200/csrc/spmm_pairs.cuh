// Trajectory-batched SELL-D kernel for generators whose operators share columns.
//
// Control Hamiltonians usually come in quadrature pairs -- H1 = sum (a + a^+), H2 = i sum (a^+ - a)
// on the transmon chain of config 3 -- whose entries sit in the SAME columns: 32 of the 45 entries
// of a row gather only 16 distinct x values.  The rows of the merged matrix are therefore ordered
// "columns hit by two operators first, pair by pair" (k_pair_order_rows), and this kernel looks at
// four codes at a time: two aligned pairs cost two gathers (per trajectory) instead of four, in the
// same registers.  The per-trajectory coefficients forbid pre-multiplying the pair, so every
// operator keeps its own partial sums p_l (folded with u_l^(b) once per row); purely real /
// imaginary operators only (one real per entry), 4 chunks of 32 trajectories per lane.
// The gather count is what bounds the batched kernel (46 GB of 16-byte gathers per
// term through L1 on config 3).
#pragma once

#include "spmv.cuh"

template <int NOPS, int T>
__device__ __forceinline__ void pairs_acc(const int op, const double v, const double2* xv, double (*pr)[T], double (*pi)[T]) {
#pragma unroll
  for (int o = 0; o < NOPS; ++o)
    if (op == o) {  // warp-uniform
#pragma unroll
      for (int t = 0; t < T; ++t) {
        pr[o][t] = fma(v, xv[t].x, pr[o][t]);
        pi[o][t] = fma(v, xv[t].y, pi[o][t]);
      }
    }
}

template <int EPI, int CB, int NOPS>
__global__ void __launch_bounds__(256, 2)
k_spmm_selld_pairs(DictView m, const double* __restrict__ dvalr, unsigned imag_ops, const double2* __restrict__ coef,
                   int coef_stride, int64_t batch, const double2* __restrict__ x, EpiArgs e) {
  constexpr int T = 4;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* s_val1 = reinterpret_cast<double*>(smem_raw);
  DeltaOp* s_dop = reinterpret_cast<DeltaOp*>(smem_raw + (size_t)m.n_dict * 8);
  for (int j = threadIdx.x; j < m.n_dict; j += blockDim.x) {
    s_val1[j] = dvalr[j];
    s_dop[j] = DeltaOp{m.ddelta[j], (int32_t)m.dop[j]};
  }
  __syncthreads();

  constexpr int CPW = 16 / CB;  // codes per 16-byte word
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int32_t bcol[T];
  bool blive[T];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const int64_t b = ((int64_t)blockIdx.y * T + t) * 32 + lane;
    blive[t] = b < batch;
    bcol[t] = (int32_t)(blive[t] ? b : batch - 1);
  }
  const int64_t ustride = coef_stride ? batch : 1;
  const int64_t slice = blockIdx.x;
  uint32_t off0, off1;
  if (m.uniform_words) {
    off0 = (uint32_t)slice * m.uniform_words;
    off1 = off0 + m.uniform_words;
  } else {
    off0 = m.sptr[slice];
    off1 = m.sptr[slice + 1];
  }
  double dr[T], di[T], nn[T];
#pragma unroll
  for (int t = 0; t < T; ++t) dr[t] = di[t] = nn[t] = 0.0;

  for (int rl = warp; rl < QP_SELL_C; rl += 8) {
    const int64_t row = slice * QP_SELL_C + rl;
    if (row >= m.n) break;
    const double2* xb0 = x + row * batch;  // x[(row + delta) * batch + b] = xb0[delta * batch + b]
    double pr[NOPS][T], pi[NOPS][T];
#pragma unroll
    for (int o = 0; o < NOPS; ++o)
#pragma unroll
      for (int t = 0; t < T; ++t) pr[o][t] = pi[o][t] = 0.0;

    uint4 c_next = make_uint4(0u, 0u, 0u, 0u);
    if (off0 + rl < off1) c_next = __ldg(m.codes + off0 + rl);  // same word for all lanes: one broadcast load
    for (uint32_t off = off0 + rl; off < off1; off += QP_SELL_C) {
      const uint4 c = c_next;
      if (off + QP_SELL_C < off1) c_next = __ldg(m.codes + off + QP_SELL_C);
      const uint32_t w[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
      for (int h = 0; h < CPW; h += 4) {  // four codes at a time
        uint32_t code[4];
        bool any = false;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int tt = h + g;
          code[g] = CB == 1 ? (w[tt >> 2] >> (8 * (tt & 3))) & 0xffu : (w[tt >> 1] >> (16 * (tt & 1))) & 0xffffu;
          any |= code[g] != 0u;
        }
        if (!any) continue;  // all padding (warp-uniform)
        DeltaOp dop[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) dop[g] = s_dop[code[g]];
        const bool pair_a = code[0] != 0u && code[1] != 0u && dop[0].delta == dop[1].delta;
        const bool pair_b = code[2] != 0u && code[3] != 0u && dop[2].delta == dop[3].delta;
        double2 xv[2][T];
        if (pair_a && pair_b) {  // two aligned pairs: two gathers feed four entries
          const int64_t oa = (int64_t)dop[0].delta * batch, ob = (int64_t)dop[2].delta * batch;
#pragma unroll
          for (int t = 0; t < T; ++t) {
            xv[0][t] = __ldg(xb0 + oa + bcol[t]);
            xv[1][t] = __ldg(xb0 + ob + bcol[t]);
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) pairs_acc<NOPS, T>(dop[g].op, s_val1[code[g]], xv[g >> 1], pr, pi);
        } else {  // singles (or a pair next to a single): two rounds of two gathers
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            if ((code[2 * half] | code[2 * half + 1]) == 0u) continue;
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              const int64_t o = (int64_t)dop[2 * half + g].delta * batch;
#pragma unroll
              for (int t = 0; t < T; ++t) xv[g][t] = __ldg(xb0 + o + bcol[t]);
            }
#pragma unroll
            for (int g = 0; g < 2; ++g)
              if (code[2 * half + g] != 0u) pairs_acc<NOPS, T>(dop[2 * half + g].op, s_val1[code[2 * half + g]], xv[g], pr, pi);
          }
        }
      }
    }
    // fold the operators with their per-trajectory coefficients (times i for an imaginary operator)
    double tr[T], ti[T];
#pragma unroll
    for (int t = 0; t < T; ++t) tr[t] = ti[t] = 0.0;
#pragma unroll
    for (int o = 0; o < NOPS; ++o)
#pragma unroll
      for (int t = 0; t < T; ++t) {
        double2 u = __ldg(coef + (int64_t)o * ustride + (coef_stride ? bcol[t] : 0));
        if ((imag_ops >> o) & 1u) u = make_double2(-u.y, u.x);
        tr[t] += u.x * pr[o][t] - u.y * pi[o][t];
        ti[t] += u.x * pi[o][t] + u.y * pr[o][t];
      }
    double2 xr[T], yv[T], av[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int64_t idx = row * batch + bcol[t];
      epi_load<EPI>(e, x, idx, idx, xr[t], yv[t], av[t]);
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
        if (EPI == EPI_DOT) {
          epi_apply<EPI>(e, row * batch + bcol[t], make_double2(tr[t], ti[t]), xr[t], yv[t], av[t], dr[t], di[t], nn[t]);
        } else {
          double cr = 0, ci = 0, cn = 0;
          epi_apply<EPI>(e, row * batch + bcol[t], make_double2(tr[t], ti[t]), xr[t], yv[t], av[t], cr, ci, cn);
          if (epi_has_sums(EPI) && e.chk != nullptr) chk_flush(e, bcol[t], cr, ci, cn);
        }
      }
  }
  if (EPI == EPI_DOT && e.chk != nullptr) {
#pragma unroll
    for (int t = 0; t < T; ++t)
      if (blive[t]) chk_flush(e, bcol[t], dr[t], di[t], nn[t]);
  }
}
