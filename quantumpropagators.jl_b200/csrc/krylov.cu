// Arnoldi / Newton vector kernels: replaces arnoldi! (reference src/arnoldi.jl:60-100),
// extend_arnoldi! (:115-129) and the vector updates of newton! (src/newton.jl:346-367).
//
// Orthogonalisation: the reference runs modified Gram-Schmidt, j sequential dot+axpy pairs
// per column, i.e. 2j passes over the new vector.  Here one fused multi-dot pass computes the
// whole Hessenberg column (plus |w|^2), one fused pass subtracts the projections and returns
// the new norm, and a second (DGKS) round is taken only when the norm dropped by more than
// 1/sqrt(2) -- classical Gram-Schmidt with selective re-orthogonalisation, which is at least
// as orthogonal as MGS.  The Hessenberg entries differ from the reference's at rounding
// level; parity is asserted on the propagated state.
#include <cmath>
#include <cstring>

#include "spmv.cuh"

constexpr int KV = 12;          // Krylov vectors handled per fused pass (m_max <= 11: every column in one chunk)
constexpr int KBLOCK = 256;

struct Weights {
  double2 w[KV];
};

// Per-column control block on the device: the whole Arnoldi process runs without a host
// round trip (the host reads all columns once at the end).
struct ColCtl {
  double ww;     // |w|^2 before the projection
  double nrm2;   // |w|^2 after the (last) projection
  double inv;    // 1 / |w|, used by the normalisation kernel
  double again;  // != 0: the DGKS criterion asks for a second projection round
};
// factor of the stored Krylov vector idx (ctl_base == nullptr: vectors are stored normalised)
__device__ __forceinline__ double krylov_scale(const ColCtl* ctl_base, int idx) {
  return (ctl_base == nullptr || idx == 0) ? 1.0 : ctl_base[idx - 1].inv;
}


// Lazy normalisation (single states; QPROP_KRYLOV_EAGER=1 restores the separate pass): a new Krylov vector is
// NOT rescaled after its orthogonalisation -- a read-modify-write pass over the vector, 6 % of a Newton step on
// config 4 -- but stored as it is, and its factor 1 / |w| (ColCtl::inv, still on the device) is applied where
// the vector is used: as a device-side factor of alpha in the next application of the generator, to the dot
// products and projection coefficients of the later columns (below), and by the host to the weights of
// qp_krylov_combine and to the copy of qp_krylov_get.  q[0] and vectors set by qp_krylov_set have factor 1.
static bool krylov_lazy() {
  static const int eager = getenv("QPROP_KRYLOV_EAGER") ? atoi(getenv("QPROP_KRYLOV_EAGER")) : 0;
  return !eager;
}

// partial[block][KV+1][2]: <q_i|w> for i < nv and |w|^2 in slot KV.  `gate` (or nullptr): the
// launch is a no-op unless *gate != 0 (second Gram-Schmidt round, decided on the device).
template <int NV>
__global__ void __launch_bounds__(KBLOCK)
k_multidot(const double2* __restrict__ q0, int64_t stride, const double2* __restrict__ w, int64_t n,
           double* __restrict__ partial, const double* __restrict__ gate) {
  pdl_sync();
  if (gate != nullptr && *gate == 0.0) return;
  double sr[NV], si[NV], ww = 0.0;
#pragma unroll
  for (int i = 0; i < NV; ++i) sr[i] = si[i] = 0.0;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    const double2 wv = w[r];
    ww += wv.x * wv.x + wv.y * wv.y;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const double2 qv = q0[i * stride + r];
      sr[i] += qv.x * wv.x + qv.y * wv.y;  // conj(q) * w
      si[i] += qv.x * wv.y - qv.y * wv.x;
    }
  }
  __shared__ double sh[KBLOCK / 32][2 * NV + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 16; o > 0; o >>= 1) {
    ww += __shfl_xor_sync(0xffffffffu, ww, o);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      sr[i] += __shfl_xor_sync(0xffffffffu, sr[i], o);
      si[i] += __shfl_xor_sync(0xffffffffu, si[i], o);
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      sh[warp][2 * i] = sr[i];
      sh[warp][2 * i + 1] = si[i];
    }
    sh[warp][2 * NV] = ww;
  }
  __syncthreads();
  if (threadIdx.x < 2 * NV + 1) {
    double t = 0.0;
    for (int k = 0; k < KBLOCK / 32; ++k) t += sh[k][threadIdx.x];
    const int slot = threadIdx.x == 2 * NV ? 2 * KV : threadIdx.x;
    partial[(size_t)blockIdx.x * (2 * KV + 2) + slot] = t;
  }
}

// deterministic sum over the blocks' partials, one warp per slot (block = (2 KV + 1) warps):
// h[i] (+)= <q_i|w> ; slot 2 KV -> ctl->ww (first pass of the first round only)
__global__ void __launch_bounds__(32 * (2 * KV + 1))
k_multidot_final(const double* __restrict__ partial, int nblocks, int nv, double2* __restrict__ h, int accumulate,
                 ColCtl* __restrict__ ctl, int write_ww, const double* __restrict__ gate, const ColCtl* ctl_base, int i0) {
  pdl_sync();
  if (gate != nullptr && *gate == 0.0) return;
  const int slot = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool is_ww = slot == 2 * KV;
  if (!is_ww && slot >= 2 * nv) return;
  double t = 0.0;
  for (int k = lane; k < nblocks; k += 32) t += partial[(size_t)k * (2 * KV + 2) + slot];
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if (lane != 0) return;
  if (is_ww) {
    if (write_ww) ctl->ww = t;
  } else {
    t *= krylov_scale(ctl_base, i0 + (slot >> 1));  // <q_i|w> with q_i = scale_i * (stored vector)
    double* dst = reinterpret_cast<double*>(h + (slot >> 1)) + (slot & 1);
    *dst = accumulate ? *dst + t : t;
  }
}

// w -= sum_i h[i] q_i ; partial[block] = |w_new|^2 contribution.  `h` holds this round's
// coefficients (round 2 uses its own correction, kept apart from the accumulated column).
template <int NV>
__global__ void __launch_bounds__(KBLOCK)
k_project_out(const double2* __restrict__ q0, int64_t stride, const double2* __restrict__ h,
              double2* __restrict__ w, int64_t n, double* __restrict__ partial, const double* __restrict__ gate,
              const ColCtl* ctl_base, int i0) {
  pdl_sync();
  if (gate != nullptr && *gate == 0.0) return;
  double2 hv[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const double sc = krylov_scale(ctl_base, i0 + i);
    hv[i] = make_double2(sc * h[i].x, sc * h[i].y);
  }
  double nn = 0.0;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    double2 wv = w[r];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const double2 qv = q0[i * stride + r];
      wv.x -= hv[i].x * qv.x - hv[i].y * qv.y;
      wv.y -= hv[i].x * qv.y + hv[i].y * qv.x;
    }
    w[r] = wv;
    nn += wv.x * wv.x + wv.y * wv.y;
  }
  __shared__ double sh[KBLOCK / 32];
  for (int o = 16; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = nn;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int k = 0; k < KBLOCK / 32; ++k) t += sh[k];
    partial[blockIdx.x] = t;
  }
}

// Round 1 of Gram-Schmidt and the multi-dot of a possible round 2 in ONE pass over q_0..q_{NV-1} and w:
// w' = w - sum_i h[i] q_i is written back, and <q_i|w'> (the coefficients a second round would need) and |w'|^2
// are accumulated from the values still in registers.  Saves the (NV + 1)-vector read pass of the second round's
// multi-dot whenever the DGKS criterion fires (on config 4: in most columns).  dots[block][2 KV + 2] as in
// k_multidot, norms[block] as in k_project_out.
template <int NV>
__global__ void __launch_bounds__(KBLOCK)
k_project_multidot(const double2* __restrict__ q0, int64_t stride, const double2* __restrict__ h, double2* __restrict__ w,
                   int64_t n, double* __restrict__ dots, double* __restrict__ norms, const ColCtl* ctl_base) {
  pdl_sync();
  double2 hv[NV];
  double sr[NV], si[NV], nn = 0.0;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const double sc = krylov_scale(ctl_base, i);
    hv[i] = make_double2(sc * h[i].x, sc * h[i].y);
    sr[i] = si[i] = 0.0;
  }
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    double2 wv = w[r];
    double2 qv[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) qv[i] = q0[i * stride + r];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      wv.x -= hv[i].x * qv[i].x - hv[i].y * qv[i].y;
      wv.y -= hv[i].x * qv[i].y + hv[i].y * qv[i].x;
    }
    w[r] = wv;
    nn += wv.x * wv.x + wv.y * wv.y;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      sr[i] += qv[i].x * wv.x + qv[i].y * wv.y;  // conj(q) * w'
      si[i] += qv[i].x * wv.y - qv[i].y * wv.x;
    }
  }
  __shared__ double sh[KBLOCK / 32][2 * NV + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 16; o > 0; o >>= 1) {
    nn += __shfl_xor_sync(0xffffffffu, nn, o);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      sr[i] += __shfl_xor_sync(0xffffffffu, sr[i], o);
      si[i] += __shfl_xor_sync(0xffffffffu, si[i], o);
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      sh[warp][2 * i] = sr[i];
      sh[warp][2 * i + 1] = si[i];
    }
    sh[warp][2 * NV] = nn;
  }
  __syncthreads();
  if (threadIdx.x < 2 * NV + 1) {
    double t = 0.0;
    for (int k = 0; k < KBLOCK / 32; ++k) t += sh[k][threadIdx.x];
    if (threadIdx.x == 2 * NV) norms[blockIdx.x] = t;
    else dots[(size_t)blockIdx.x * (2 * KV + 2) + threadIdx.x] = t;
  }
}

// one warp: |w|^2 after the projection; round 1 decides on the device whether a second
// (DGKS) round is needed -- when the projection removed more than half of |w|^2 -- and the
// second round adds its correction to the accumulated column
__global__ void k_norm_decide(const double* __restrict__ partial, int nblocks, ColCtl* __restrict__ ctl, int round,
                              double2* __restrict__ h_acc, const double2* __restrict__ h_corr, int nvec, double norm_min,
                              double eta2) {
  pdl_sync();
  if (round == 2 && ctl->again == 0.0) return;
  const int lane = threadIdx.x;
  double t = 0.0;
  for (int k = lane; k < nblocks; k += 32) t += partial[k];
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if (round == 2)
    for (int i = lane; i < nvec; i += 32) {
      h_acc[i].x += h_corr[i].x;
      h_acc[i].y += h_corr[i].y;
    }
  if (lane == 0) {
    ctl->nrm2 = t;
    // Krylov space exhausted (|w| < norm_min, src/arnoldi.jl:92-95): the reference leaves the
    // residue unnormalised (a vector of norm < 1e-15); here it is zeroed, so that neither this
    // vector nor the columns computed after it (discarded by the host) ever hold inf / NaN --
    // the restart vector of newton! combines it with weight R[m+1] (src/newton.jl:360-366)
    const double nrm = sqrt(t);
    ctl->inv = (t > 0.0 && nrm >= norm_min) ? 1.0 / nrm : 0.0;
    // second (DGKS) round when the projection removed more than a fraction 1 - eta^2 of |w|^2, with
    // the textbook eta = 1/sqrt(2).  A laxer eta = 0.1 saves 7-9 % of a Newton step on config 4 (measured:
    // 9.89 -> 9.22 ms at N = 2^22, 40.2 -> 36.5 ms at N = 2^24) but classical Gram-Schmidt then loses
    // orthogonality like eps (|w| / |w'|)^2 per column and the optomech parity pin (Newton == Cheby to 1e-10
    // after 250 steps, test/test_propagate.jl:153-163) degrades to 1.2e-10: parity first.  QPROP_DGKS_ETA
    // overrides.
    if (round == 1) ctl->again = (t < eta2 * ctl->ww && t != 0.0) ? 1.0 : 0.0;
  }
}

__global__ void k_scale_real(double2* __restrict__ x, double s, int64_t n) {
  pdl_sync();
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    double2 v = x[r];
    x[r] = make_double2(s * v.x, s * v.y);
  }
}

// x *= ctl->inv (normalisation with the norm still on the device)
__global__ void k_scale_dev(double2* __restrict__ x, const ColCtl* __restrict__ ctl, int64_t n) {
  pdl_sync();
  const double s = ctl->inv;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    double2 v = x[r];
    x[r] = make_double2(s * v.x, s * v.y);
  }
}

// y (+)= sum_i w_i q_i
template <int NV>
__global__ void __launch_bounds__(KBLOCK)
k_combine(const double2* __restrict__ q0, int64_t stride, Weights wt, double2* __restrict__ y, int64_t n,
          int accumulate) {
  pdl_sync();
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    double2 acc = accumulate ? y[r] : make_double2(0.0, 0.0);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const double2 qv = q0[i * stride + r];
      acc.x += wt.w[i].x * qv.x - wt.w[i].y * qv.y;
      acc.y += wt.w[i].x * qv.y + wt.w[i].y * qv.x;
    }
    y[r] = acc;
  }
}

static int kgrid(qp_ctx_t ctx, int64_t n) {
  int64_t want = (n + KBLOCK - 1) / KBLOCK;
  int64_t cap = (int64_t)ctx->sm_count * 4;
  return (int)std::max<int64_t>(1, std::min(want, cap));
}

#define DISPATCH_NV(nv, CALL)          \
  switch (nv) {                        \
    case 1: { constexpr int NV = 1; CALL; } break; \
    case 2: { constexpr int NV = 2; CALL; } break; \
    case 3: { constexpr int NV = 3; CALL; } break; \
    case 4: { constexpr int NV = 4; CALL; } break; \
    case 5: { constexpr int NV = 5; CALL; } break; \
    case 6: { constexpr int NV = 6; CALL; } break; \
    case 7: { constexpr int NV = 7; CALL; } break; \
    case 8: { constexpr int NV = 8; CALL; } break; \
    case 9: { constexpr int NV = 9; CALL; } break; \
    case 10: { constexpr int NV = 10; CALL; } break; \
    case 11: { constexpr int NV = 11; CALL; } break; \
    default: { constexpr int NV = 12; CALL; } break; \
  }

// ---------------------------------------------------------------------------------------

// ---------------------------------------------------------------------------------------
// State bundles (batch > 1): B states sharing one generator run the Arnoldi process in lock step,
// each with its own Hessenberg matrix (the forward / backward sweeps of GRAPE and
// Krotov propagate many states under one generator).  Orthogonalisation here is the reference's own
// modified Gram-Schmidt -- j dot + axpy pairs per column (src/arnoldi.jl:84-87) -- with every dot
// product, norm and update a batched kernel over the [N][B] layout.
// ---------------------------------------------------------------------------------------

// y[r, b] += alpha[b] * x[r, b]
__global__ void k_axpy_pb(const double2* __restrict__ alpha, const double2* __restrict__ x, double2* __restrict__ y,
                          int64_t n, int64_t batch) {
  const int64_t total = n * batch;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const double2 a = alpha[i % batch];
    const double2 xv = x[i];
    double2 yv = y[i];
    yv.x += a.x * xv.x - a.y * xv.y;
    yv.y += a.x * xv.y + a.y * xv.x;
    y[i] = yv;
  }
}

// x[r, b] *= alpha[b]
__global__ void k_scal_pb(const double2* __restrict__ alpha, double2* __restrict__ x, int64_t n, int64_t batch) {
  const int64_t total = n * batch;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const double2 a = alpha[i % batch];
    const double2 xv = x[i];
    x[i] = make_double2(a.x * xv.x - a.y * xv.y, a.x * xv.y + a.y * xv.x);
  }
}

// y[r, b] (+)= sum_i w[i * batch + b] * q_i[r, b]
__global__ void k_combine_pb(const double2* __restrict__ q0, int64_t stride, const double2* __restrict__ w, int n_w,
                             double2* __restrict__ y, int64_t n, int64_t batch, int accumulate) {
  const int64_t total = n * batch;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i % batch;
    double2 acc = accumulate ? y[i] : make_double2(0.0, 0.0);
    for (int k = 0; k < n_w; ++k) {
      const double2 a = w[(int64_t)k * batch + b];
      const double2 qv = q0[(int64_t)k * stride + i];
      acc.x += a.x * qv.x - a.y * qv.y;
      acc.y += a.x * qv.y + a.y * qv.x;
    }
    y[i] = acc;
  }
}

// per-trajectory numbers -> device (through the context's pinned staging area)
static int32_t upload_pb(qp_krylov_t K, const double2* host, size_t count) {
  qp_ctx_t ctx = K->ctx;
  if (K->pb_elems < count) {
    QP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(K->d_pb);
    K->d_pb = nullptr;
    K->pb_elems = 0;
    QP_CUDA(ctx, cudaMalloc(&K->d_pb, sizeof(double2) * count));
    K->pb_elems = count;
  }
  QP_CHECK(qp_ctx_reserve_stage(ctx, count));
  QP_CUDA(ctx, cudaEventSynchronize(ctx->ev_stage));
  memcpy(ctx->h_stage, host, sizeof(double2) * count);
  QP_CUDA(ctx, cudaMemcpyAsync(K->d_pb, ctx->h_stage, sizeof(double2) * count, cudaMemcpyHostToDevice, ctx->stream));
  QP_CUDA(ctx, cudaEventRecord(ctx->ev_stage, ctx->stream));
  return QP_OK;
}

// st[:, b] *= alpha[b]  (used by the batched newton! in newton.cu)
int32_t qp_krylov_scale_pb(qp_krylov_t K, qp_state_t st, const qp_c128* alpha) {
  qp_ctx_t ctx = K->ctx;
  QP_CHECK(upload_pb(K, reinterpret_cast<const double2*>(alpha), (size_t)K->batch));
  k_scal_pb<<<kgrid(ctx, K->n * K->batch), KBLOCK, 0, ctx->stream>>>(K->d_pb, st->d, K->n, K->batch);
  QP_LAUNCHED(ctx);
  return QP_OK;
}

static int32_t scal_pb(qp_krylov_t K, double2* x, const std::vector<double2>& alpha) {
  qp_ctx_t ctx = K->ctx;
  QP_CHECK(upload_pb(K, alpha.data(), alpha.size()));
  k_scal_pb<<<kgrid(ctx, K->n * K->batch), KBLOCK, 0, ctx->stream>>>(K->d_pb, x, K->n, K->batch);
  QP_LAUNCHED(ctx);
  return QP_OK;
}

// One Arnoldi column for every state of the bundle: q[j+1] = H q[j], MGS against q[0..j].
// hess: [batch][ld * ld] column-major matrices; alive[b] = 0 once trajectory b's Krylov space is exhausted.
static int32_t arnoldi_column_batched(qp_krylov_t K, int stride, int j, double dt, bool normalise, double norm_min,
                                      qp_c128* hess, int ld, std::vector<int>& m_eff, int m) {
  qp_ctx_t ctx = K->ctx;
  const int64_t n = K->n, B = K->batch, vec = n * B;
  double2* w = K->q + (size_t)(j + 1) * vec;
  QP_CHECK(qp_gen_apply(K->gen, stride, make_double2(1.0, 0.0), make_double2(0.0, 0.0), K->q + (size_t)j * vec, w, B));
  std::vector<qp_c128> h((size_t)B);
  std::vector<double2> alpha((size_t)B);
  for (int i = 0; i <= j; ++i) {  // Hess[i, j] = dt <q_i|w>;  w -= (Hess[i, j] / dt) q_i   (:84-87)
    QP_CHECK(qp_reduce_dot(ctx, K->q + (size_t)i * vec, w, n, B, h.data()));
    for (int64_t b = 0; b < B; ++b) {
      hess[(size_t)b * ld * ld + (size_t)j * ld + i] = qp_c128{dt * h[b].re, dt * h[b].im};
      alpha[b] = make_double2(-h[b].re, -h[b].im);
    }
    QP_CHECK(upload_pb(K, alpha.data(), alpha.size()));
    k_axpy_pb<<<kgrid(ctx, vec), KBLOCK, 0, ctx->stream>>>(K->d_pb, K->q + (size_t)i * vec, w, n, B);
    QP_LAUNCHED(ctx);
  }
  if (normalise) {  // :88-97
    std::vector<double> nrm2((size_t)B);
    QP_CHECK(qp_reduce_norm2(ctx, w, n, B, nrm2.data()));
    for (int64_t b = 0; b < B; ++b) {
      const double hn = sqrt(nrm2[b]);
      hess[(size_t)b * ld * ld + (size_t)j * ld + (j + 1)] = qp_c128{dt * hn, 0.0};
      const bool ok = hn >= norm_min && hn > 0.0;
      if (!ok && m_eff[b] == m) m_eff[b] = j + 1;  // dimensionality exhausted for this state
      alpha[b] = make_double2(ok ? 1.0 / hn : 0.0, 0.0);  // an exhausted state keeps a zero vector (no inf / NaN)
    }
    QP_CHECK(scal_pb(K, w, alpha));
  }
  return QP_OK;
}

static int32_t arnoldi_batched(qp_krylov_t K, const qp_c128* op_coeffs, qp_state_t v, int32_t m, double dt,
                               int32_t extended, double norm_min, qp_c128* hess, int32_t ld, int32_t* m_out) {
  qp_ctx_t ctx = K->ctx;
  const int64_t B = K->batch, vec = K->n * B;
  QpScopedTimer timer(ctx, "arnoldi!");
  int stride = 0;
  QP_CHECK(qp_gen_set_coeffs(K->gen, op_coeffs, 0, 1, &stride));
  memset(hess, 0, sizeof(qp_c128) * (size_t)ld * (size_t)ld * (size_t)B);
  QP_CUDA(ctx, cudaMemcpyAsync(K->q, v->d, sizeof(double2) * vec, cudaMemcpyDeviceToDevice, ctx->stream));
  std::vector<int> m_eff((size_t)B, m);
  for (int j = 0; j < m; ++j)
    QP_CHECK(arnoldi_column_batched(K, stride, j, dt, j + 1 < m || extended, norm_min, hess, ld, m_eff, m));
  for (int64_t b = 0; b < B; ++b) m_out[b] = m_eff[b];
  return QP_OK;
}

extern "C" int32_t qp_krylov_create(qp_gen_t gen, qp_state_t like, int32_t m_max, qp_krylov_t* out) {
  if (!gen) return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_krylov_create: null generator");
  qp_ctx_t ctx = gen->ctx;
  QP_CHECK(qp_ctx_bind(ctx));
  QP_REQUIRE(ctx, out != nullptr && like != nullptr, "qp_krylov_create: null argument");
  *out = nullptr;
  QP_REQUIRE(ctx, like->ctx == ctx && like->n == gen->n, "qp_krylov_create: state does not match the generator");
  QP_REQUIRE(ctx, m_max >= 1, "qp_krylov_create: m_max must be >= 1");
  qp_krylov_t K = new qp_krylov_s();
  K->ctx = ctx;
  K->gen = gen;
  K->n = like->n;
  K->batch = like->batch;
  K->m_max = m_max;
  cudaError_t e;
  if ((e = cudaMalloc(&K->q, sizeof(double2) * (size_t)K->n * (size_t)K->batch * (size_t)(m_max + 1))) != cudaSuccess ||
      (e = cudaMalloc(&K->d_h, sizeof(double2) * (size_t)(m_max + 8))) != cudaSuccess ||
      (e = cudaMalloc(&K->d_hall, sizeof(double2) * (size_t)(m_max + 1) * (size_t)(m_max + 2))) != cudaSuccess ||
      (e = cudaMalloc(&K->d_ctl, sizeof(ColCtl) * (size_t)(m_max + 1))) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(K->q);
    cudaFree(K->d_h);
    cudaFree(K->d_hall);
    delete K;
    return qp_fail(ctx, QP_ERR_OOM, "qp_krylov_create: cudaMalloc of %d Krylov vectors failed: %s", m_max + 1,
                   cudaGetErrorString(e));
  }
  *out = K;
  return QP_OK;
}

extern "C" int32_t qp_krylov_destroy(qp_krylov_t K) {
  if (!K) return QP_OK;
  cudaSetDevice(K->ctx->device);
  cudaStreamSynchronize(K->ctx->stream);
  cudaFree(K->q);
  cudaFree(K->d_h);
  cudaFree(K->d_hall);
  cudaFree(K->d_ctl);
  cudaFree(K->d_pb);
  delete K;
  return QP_OK;
}

// Orthogonalise w = q[j+1] against q[0..j], entirely on the stream (no host synchronisation):
// column j of K->d_hall receives <q_i|w> (summed over the rounds), K->d_ctl[j] the norms.
// Classical Gram-Schmidt; the second round always gets launched but is a no-op on the device
// unless the DGKS criterion fired ("twice is enough").
static int32_t orthogonalise_async(qp_krylov_t K, int j, double norm_min) {
  qp_ctx_t ctx = K->ctx;
  const int64_t n = K->n;
  const int nvec = j + 1;
  double2* w = K->q + (size_t)(j + 1) * n;
  const int nblocks = kgrid(ctx, n);
  const size_t need = (size_t)nblocks * (2 * KV + 2);
  QP_CHECK(qp_ctx_reserve_red(ctx, need + (size_t)nblocks));
  double* partial = ctx->d_red;
  double2* h_acc = K->d_hall + (size_t)j * (K->m_max + 2);
  double2* h_corr = K->d_h;  // this round's coefficients
  ColCtl* ctl = K->d_ctl + j;
  const ColCtl* ctl_base = krylov_lazy() ? K->d_ctl : nullptr;
  static const double eta = getenv("QPROP_DGKS_ETA") ? atof(getenv("QPROP_DGKS_ETA")) : 0.70710678118654752;
  const double eta2 = eta * eta;
  static const int unfused = getenv("QPROP_GS_UNFUSED") ? atoi(getenv("QPROP_GS_UNFUSED")) : 0;
  if (nvec <= KV && !unfused) {
    // single chunk: multi-dot | projection fused with the second round's multi-dot | decision |
    // (gated) coefficients of the second round | (gated) second projection | (gated) final norm
    double* norms = partial + need;
    const double* gate = &ctl->again;
    DISPATCH_NV(nvec, (qp_launch_pdl(k_multidot<NV>, dim3(nblocks), dim3(KBLOCK), 0, ctx->stream, K->q, n, w, n, partial, (const double*)nullptr)));
    QP_LAUNCHED(ctx);
    qp_launch_pdl(k_multidot_final, dim3(1), dim3(32 * (2 * KV + 1)), 0, ctx->stream, partial, nblocks, nvec, h_acc, 0, ctl, 1,
                  (const double*)nullptr, ctl_base, 0);
    QP_LAUNCHED(ctx);
    DISPATCH_NV(nvec, (qp_launch_pdl(k_project_multidot<NV>, dim3(nblocks), dim3(KBLOCK), 0, ctx->stream, K->q, n, h_acc, w, n, partial, norms, ctl_base)));
    QP_LAUNCHED(ctx);
    qp_launch_pdl(k_norm_decide, dim3(1), dim3(32), 0, ctx->stream, norms, nblocks, ctl, 1, h_acc, h_corr, nvec, norm_min, eta2);
    QP_LAUNCHED(ctx);
    qp_launch_pdl(k_multidot_final, dim3(1), dim3(32 * (2 * KV + 1)), 0, ctx->stream, partial, nblocks, nvec, h_corr, 0, ctl, 0, gate,
                  ctl_base, 0);
    QP_LAUNCHED(ctx);
    DISPATCH_NV(nvec, (qp_launch_pdl(k_project_out<NV>, dim3(nblocks), dim3(KBLOCK), 0, ctx->stream, K->q, n, h_corr, w, n, norms, gate, ctl_base, 0)));
    QP_LAUNCHED(ctx);
    qp_launch_pdl(k_norm_decide, dim3(1), dim3(32), 0, ctx->stream, norms, nblocks, ctl, 2, h_acc, h_corr, nvec, norm_min, eta2);
    QP_LAUNCHED(ctx);
    return QP_OK;
  }
  for (int round = 1; round <= 2; ++round) {
    const double* gate = round == 2 ? &ctl->again : nullptr;
    double2* h = round == 1 ? h_acc : h_corr;
    for (int i0 = 0; i0 < nvec; i0 += KV) {
      const int nv = std::min(KV, nvec - i0);
      const double2* q0 = K->q + (size_t)i0 * n;
      DISPATCH_NV(nv, (qp_launch_pdl(k_multidot<NV>, dim3(nblocks), dim3(KBLOCK), 0, ctx->stream, q0, n, w, n, partial, gate)));
      QP_LAUNCHED(ctx);
      qp_launch_pdl(k_multidot_final, dim3(1), dim3(32 * (2 * KV + 1)), 0, ctx->stream, partial, nblocks, nv, h + i0, 0, ctl,
                    (round == 1 && i0 == 0) ? 1 : 0, gate, ctl_base, i0);
      QP_LAUNCHED(ctx);
    }
    for (int i0 = 0; i0 < nvec; i0 += KV) {
      const int nv = std::min(KV, nvec - i0);
      const double2* q0 = K->q + (size_t)i0 * n;
      DISPATCH_NV(nv, (qp_launch_pdl(k_project_out<NV>, dim3(nblocks), dim3(KBLOCK), 0, ctx->stream, q0, n, h + i0, w, n, partial, gate, ctl_base, i0)));
      QP_LAUNCHED(ctx);
    }
    qp_launch_pdl(k_norm_decide, dim3(1), dim3(32), 0, ctx->stream, partial, nblocks, ctl, round, h_acc, h_corr, nvec, norm_min, eta2);
    QP_LAUNCHED(ctx);
  }
  return QP_OK;
}

// download columns [j0, j1) of the device-side Hessenberg data (one synchronisation)
static int32_t fetch_columns(qp_krylov_t K, int j0, int j1, std::vector<double2>& h_all, std::vector<ColCtl>& ctl) {
  qp_ctx_t ctx = K->ctx;
  const size_t ldh = (size_t)K->m_max + 2;
  h_all.resize(ldh * (size_t)(j1 - j0));
  ctl.resize((size_t)(j1 - j0));
  QP_CUDA(ctx, cudaMemcpyAsync(h_all.data(), K->d_hall + ldh * j0, sizeof(double2) * h_all.size(), cudaMemcpyDeviceToHost, ctx->stream));
  QP_CUDA(ctx, cudaMemcpyAsync(ctl.data(), K->d_ctl + j0, sizeof(ColCtl) * ctl.size(), cudaMemcpyDeviceToHost, ctx->stream));
  QP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return QP_OK;
}

static int32_t krylov_matvec(qp_krylov_t K, int stride, int j) {
  // q[j+1] = H q[j]   ("matrix-vector product", src/arnoldi.jl:81-83)
  // lazy normalisation: q[j] is stored unnormalised, its factor multiplies alpha on the device
  const double* alpha_dev = (krylov_lazy() && j > 0) ? &K->d_ctl[j - 1].inv : nullptr;
  return qp_gen_apply_scaled(K->gen, stride, make_double2(1.0, 0.0), alpha_dev, make_double2(0.0, 0.0),
                             K->q + (size_t)j * K->n, K->q + (size_t)(j + 1) * K->n, 1);
}

extern "C" int32_t qp_arnoldi(qp_krylov_t K, const qp_c128* op_coeffs, qp_state_t v, int32_t m, double dt,
                              int32_t extended, double norm_min, qp_c128* hess, int32_t ld, int32_t* m_out) {
  if (!K) return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_arnoldi: null workspace");
  qp_ctx_t ctx = K->ctx;
  QP_CHECK(qp_ctx_bind(ctx));
  QP_REQUIRE(ctx, v && v->ctx == ctx && v->n == K->n && v->batch == K->batch, "qp_arnoldi: bad start vector");
  QP_REQUIRE(ctx, hess != nullptr && m_out != nullptr, "qp_arnoldi: null output");
  QP_REQUIRE(ctx, m >= 1, "qp_arnoldi: m must be >= 1");
  // @assert length(q) >= m + 1 ; size(Hess) >= dim_hess   src/arnoldi.jl:76-77
  QP_REQUIRE(ctx, m <= K->m_max, "qp_arnoldi: m=%d exceeds the workspace (m_max=%d)", m, K->m_max);
  const int dim_hess = extended ? m + 1 : m;
  QP_REQUIRE(ctx, ld >= dim_hess, "qp_arnoldi: Hessenberg storage %d x %d too small for dimension %d", ld, ld, dim_hess);
  if (K->batch > 1) return arnoldi_batched(K, op_coeffs, v, m, dt, extended, norm_min, hess, ld, m_out);
  QpScopedTimer timer(ctx, "arnoldi!");
  int stride = 0;
  QP_CHECK(qp_gen_set_coeffs(K->gen, op_coeffs, 0, 1, &stride));
  memset(hess, 0, sizeof(qp_c128) * (size_t)ld * (size_t)ld);  // fill!(Hess, 0)  :78
  const int64_t n = K->n;
  QP_CUDA(ctx, cudaMemcpyAsync(K->q, v->d, sizeof(double2) * n, cudaMemcpyDeviceToDevice, ctx->stream));  // :79
  // all m columns are enqueued without a host round trip; a column whose norm falls below
  // norm_min ends the process in the reference (:92-95) -- here the later columns are still
  // computed (on garbage) and simply discarded when the host reads the norms
  for (int j = 0; j < m; ++j) {
    QP_CHECK(krylov_matvec(K, stride, j));
    QP_CHECK(orthogonalise_async(K, j, norm_min));
    if (!krylov_lazy() && (j + 1 < m || extended)) {  // :88-97
      qp_launch_pdl(k_scale_dev, dim3(kgrid(ctx, n)), dim3(KBLOCK), 0, ctx->stream, K->q + (size_t)(j + 1) * n, K->d_ctl + j, n);
      QP_LAUNCHED(ctx);
    }
  }
  std::vector<double2> h_all;
  std::vector<ColCtl> ctl;
  QP_CHECK(fetch_columns(K, 0, m, h_all, ctl));
  const size_t ldh = (size_t)K->m_max + 2;
  K->h_scale.assign((size_t)K->m_max + 2, 1.0);
  if (krylov_lazy())  // the last vector of a non-extended run stays unnormalised, as in the reference (:88-97)
    for (int j = 0; j < m; ++j)
      if (j + 1 < m || extended) K->h_scale[(size_t)j + 1] = ctl[j].inv;
  int m_eff = m;
  for (int j = 0; j < m; ++j) {
    for (int i = 0; i <= j; ++i) {  // Hess[i,j] = dt <q_i|q_{j+1}>   :85
      hess[(size_t)j * ld + i].re = dt * h_all[ldh * j + i].x;
      hess[(size_t)j * ld + i].im = dt * h_all[ldh * j + i].y;
    }
    if (j + 1 < m || extended) {
      const double hn = sqrt(ctl[j].nrm2);
      hess[(size_t)j * ld + (j + 1)].re = dt * hn;
      hess[(size_t)j * ld + (j + 1)].im = 0.0;
      if (hn < norm_min) {  // dimensionality exhausted
        m_eff = j + 1;
        break;
      }
    }
  }
  *m_out = m_eff;
  return QP_OK;
}

extern "C" int32_t qp_arnoldi_extend(qp_krylov_t K, const qp_c128* op_coeffs, int32_t m, double dt,
                                     double norm_min, qp_c128* hess, int32_t ld, int32_t* extended_out) {
  if (!K) return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_arnoldi_extend: null workspace");
  qp_ctx_t ctx = K->ctx;
  QP_CHECK(qp_ctx_bind(ctx));
  QP_REQUIRE(ctx, hess != nullptr, "qp_arnoldi_extend: null Hessenberg matrix");
  if (K->batch != 1) return qp_fail(ctx, QP_ERR_UNSUPPORTED, "qp_arnoldi_extend: single states only (batch == 1)");
  QP_REQUIRE(ctx, m >= 2 && m <= K->m_max && ld >= m, "qp_arnoldi_extend: bad dimension m=%d (m_max=%d, ld=%d)", m,
             K->m_max, ld);
  const int64_t n = K->n;
  if (extended_out) *extended_out = 0;
  double nrm2 = 0.0;
  QP_CHECK(qp_reduce_norm2(ctx, K->q + (size_t)(m - 1) * n, n, 1, &nrm2));
  const double hn = sqrt(nrm2);
  if (hn < norm_min) return QP_OK;  // src/arnoldi.jl:116-117
  int stride = 0;
  QP_CHECK(qp_gen_set_coeffs(K->gen, op_coeffs, 0, 1, &stride));
  hess[(size_t)(m - 2) * ld + (m - 1)].re = dt * hn;  // Hess[m, m-1]
  hess[(size_t)(m - 2) * ld + (m - 1)].im = 0.0;
  if (krylov_lazy()) {  // the factor 1 / hn is already in d_ctl[m - 2].inv (written when the column was orthogonalised)
    if (K->h_scale.size() < (size_t)K->m_max + 2) K->h_scale.assign((size_t)K->m_max + 2, 1.0);
    K->h_scale[(size_t)m - 1] = 1.0 / hn;
  } else {
    qp_launch_pdl(k_scale_real, dim3(kgrid(ctx, n)), dim3(KBLOCK), 0, ctx->stream, K->q + (size_t)(m - 1) * n, 1.0 / hn, n);
    QP_LAUNCHED(ctx);
  }
  QP_CHECK(krylov_matvec(K, stride, m - 1));
  QP_CHECK(orthogonalise_async(K, m - 1, norm_min));
  std::vector<double2> h_all;
  std::vector<ColCtl> ctl;
  QP_CHECK(fetch_columns(K, m - 1, m, h_all, ctl));
  for (int i = 0; i < m; ++i) {
    hess[(size_t)(m - 1) * ld + i].re = dt * h_all[i].x;
    hess[(size_t)(m - 1) * ld + i].im = dt * h_all[i].y;
  }
  if (extended_out) *extended_out = 1;
  return QP_OK;
}

extern "C" int32_t qp_krylov_combine(qp_krylov_t K, const qp_c128* wts, int32_t first, int32_t n_w,
                                     qp_state_t st, int32_t accumulate) {
  if (!K) return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_krylov_combine: null workspace");
  qp_ctx_t ctx = K->ctx;
  QP_CHECK(qp_ctx_bind(ctx));
  QP_REQUIRE(ctx, st && st->ctx == ctx && st->n == K->n && st->batch == K->batch, "qp_krylov_combine: bad state");
  QP_REQUIRE(ctx, wts != nullptr && n_w >= 1 && first >= 0 && first + n_w <= K->m_max + 1,
             "qp_krylov_combine: vectors [%d, %d) outside the workspace", first, first + n_w);
  const int64_t n = K->n;
  if (K->batch > 1) {  // weights [n_w][batch]
    const int64_t vec = n * K->batch;
    QP_CHECK(upload_pb(K, reinterpret_cast<const double2*>(wts), (size_t)n_w * (size_t)K->batch));
    k_combine_pb<<<kgrid(ctx, vec), KBLOCK, 0, ctx->stream>>>(K->q + (size_t)first * vec, vec, K->d_pb, n_w, st->d, n, K->batch,
                                                               accumulate ? 1 : 0);
    QP_LAUNCHED(ctx);
    return QP_OK;
  }
  for (int i0 = 0; i0 < n_w; i0 += KV) {
    const int nv = std::min(KV, n_w - i0);
    Weights wt;
    memset(&wt, 0, sizeof(wt));
    for (int i = 0; i < nv; ++i) {
      const size_t idx = (size_t)(first + i0 + i);
      const double sc = (krylov_lazy() && idx < K->h_scale.size()) ? K->h_scale[idx] : 1.0;
      wt.w[i] = make_double2(sc * wts[i0 + i].re, sc * wts[i0 + i].im);
    }
    const double2* q0 = K->q + (size_t)(first + i0) * n;
    const int acc = (accumulate || i0 > 0) ? 1 : 0;
    DISPATCH_NV(nv, (qp_launch_pdl(k_combine<NV>, dim3(kgrid(ctx, n)), dim3(KBLOCK), 0, ctx->stream, q0, n, wt, st->d, n, acc)));
    QP_LAUNCHED(ctx);
  }
  return QP_OK;
}

extern "C" int32_t qp_krylov_get(qp_krylov_t K, int32_t index, qp_state_t dst) {
  if (!K) return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_krylov_get: null workspace");
  qp_ctx_t ctx = K->ctx;
  QP_CHECK(qp_ctx_bind(ctx));
  QP_REQUIRE(ctx, dst && dst->ctx == ctx && dst->n == K->n && dst->batch == K->batch, "qp_krylov_get: bad state");
  QP_REQUIRE(ctx, index >= 0 && index <= K->m_max, "qp_krylov_get: index %d out of range", index);
  const size_t vec = (size_t)K->n * (size_t)K->batch;
  QP_CUDA(ctx, cudaMemcpyAsync(dst->d, K->q + (size_t)index * vec, sizeof(double2) * vec, cudaMemcpyDeviceToDevice, ctx->stream));
  if (K->batch == 1 && krylov_lazy() && (size_t)index < K->h_scale.size() && K->h_scale[(size_t)index] != 1.0) {
    qp_launch_pdl(k_scale_real, dim3(kgrid(ctx, K->n)), dim3(KBLOCK), 0, ctx->stream, dst->d, K->h_scale[(size_t)index], K->n);
    QP_LAUNCHED(ctx);
  }
  return QP_OK;
}

extern "C" int32_t qp_krylov_set(qp_krylov_t K, int32_t index, qp_state_t src) {
  if (!K) return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_krylov_set: null workspace");
  qp_ctx_t ctx = K->ctx;
  QP_CHECK(qp_ctx_bind(ctx));
  QP_REQUIRE(ctx, src && src->ctx == ctx && src->n == K->n && src->batch == K->batch, "qp_krylov_set: bad state");
  QP_REQUIRE(ctx, index >= 0 && index <= K->m_max, "qp_krylov_set: index %d out of range", index);
  const size_t vec = (size_t)K->n * (size_t)K->batch;
  QP_CUDA(ctx, cudaMemcpyAsync(K->q + (size_t)index * vec, src->d, sizeof(double2) * vec, cudaMemcpyDeviceToDevice, ctx->stream));
  if (K->batch == 1 && krylov_lazy()) {  // a vector that comes from outside carries no pending factor
    if (K->h_scale.size() < (size_t)K->m_max + 2) K->h_scale.assign((size_t)K->m_max + 2, 1.0);
    K->h_scale[(size_t)index] = 1.0;
    if (index > 0) {
      const double one = 1.0;
      QP_CUDA(ctx, cudaMemcpyAsync(&K->d_ctl[index - 1].inv, &one, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
      QP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
  }
  return QP_OK;
}
