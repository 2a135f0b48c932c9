// Two-pass tiled SpMM for trajectory-batched states (format: tile_format.h).
//
// One launch = one fused term (mul!, or a Chebyshev term with its epilogue) on ALL trajectories,
// executed by a persistent grid of one 512-thread CTA per SM that walks a fixed list of work items
//
//     A(0) | A(1) B(0) | A(2) B(1) | ... | A(C-1) B(C-2) | B(C-1)        C = batch / 32 chunks
//
// A(g): pass A on trajectory chunk g -- for each tile of S consecutive rows x 32 trajectories the
//       tile of X is brought into shared memory once (cp.async, L2 only), every class-A entry and the
//       diagonal are applied from there and the partial product T = sum_l u_l (H_l^A X) is written to
//       a ring of three chunk-sized buffers (which therefore lives in L2, never in HBM);
// B(g): pass B -- tiles of the N/S rows that share a position inside their blocks; class-B entries
//       from shared memory, the few class-O entries gathered from global memory, T added, and the
//       fused epilogue (src/cheby.jl:186-209 / mul!, src/generators.jl:634-645) applied and stored.
//
// B(g) needs all of A(g): items are handed out round-robin in list order and every chunk has a
// monotonic completion counter per pass (release by the producer CTA, acquire by the consumer), so a
// consumer only ever waits for items that were started about two item-rounds earlier.  The grid is
// launched cooperatively (all CTAs resident), which makes the spin-waits deadlock-free.
//
// Inside a tile a warp owns a row and its 32 lanes own the 32 trajectories: the code stream, the
// table look-ups and the class tests are warp-uniform, the loads of X are 512 B wide and
// conflict-free, and a column shared by several operators costs one load.
#include <cuda_runtime.h>

#include <cstring>

#include "spmv.cuh"
#include "tile_format.h"

using qptile::TileEntry;

constexpr int TILE_THREADS = 512;
constexpr int TILE_WARPS = TILE_THREADS / 32;
constexpr int TILE_RING = 3;

struct TileView {
  const TileEntry* tab;
  int n_tab;
  const uint4* codesA;  // [n][WA8] words of eight 16-bit codes
  const uint4* codesB;  // [n][WB8]
  const double* diag;   // [n][3]
  int WA8, WB8;
  int S, NH, n_ops;
  unsigned imag_ops;
  int64_t n;
  double2* tring;                // [TILE_RING][n][32]
  unsigned long long* doneA;     // [chunks] items of pass A finished (monotonic over launches)
  unsigned long long* doneB;     // [chunks]
  unsigned long long epoch;      // 1-based launch number of this view
};

struct qp_tile_s {
  bool ok = false;
  std::string why;
  qptile::TileFormat meta;  // host copy without the big arrays (cleared after upload)
  TileEntry* d_tab = nullptr;
  uint4* d_codesA = nullptr;
  uint4* d_codesB = nullptr;
  double* d_diag = nullptr;
  double2* d_ring = nullptr;
  unsigned long long* d_done = nullptr;  // [2][chunk capacity]
  int64_t chunk_cap = 0;
  int64_t chunks_cur = 0;  // chunk count the counters have been counting with since their last reset
  unsigned long long epoch = 0;
  int n_tab = 0;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double2 ld_cg(const double2* p) {
  double2 r;
  asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_cg(double2* p, double2 v) {
  asm volatile("st.global.cg.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

// spin until *ctr >= target; a dependency that never arrives is a bug -- trap instead of hanging the GPU
__device__ __forceinline__ void wait_counter(const unsigned long long* ctr, unsigned long long target) {
  const long long t0 = clock64();
  while (ld_acquire(ctr) < target) {
    __nanosleep(64);
    if (clock64() - t0 > (1ll << 33)) __trap();  // ~4 s
  }
}

template <int NOPS>
__device__ __forceinline__ void tile_acc(const TileEntry& en, const double2 xv, double (&pr)[NOPS], double (&pi)[NOPS]) {
#pragma unroll
  for (int l = 0; l < NOPS; ++l)
    if ((en.km >> l) & 1u) {  // warp-uniform
      pr[l] = fma(en.v[l], xv.x, pr[l]);
      pi[l] = fma(en.v[l], xv.y, pi[l]);
    }
}

template <int EPI, int NOPS>
__global__ void __launch_bounds__(TILE_THREADS, 1)
k_spmm_tile(TileView tv, const double2* __restrict__ coef, int coef_stride, int64_t batch,
            const double2* __restrict__ x, EpiArgs e) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double2* s_x = reinterpret_cast<double2*>(smem_raw);  // [rows][32]
  const int max_rows = tv.S > tv.NH ? tv.S : tv.NH;
  TileEntry* s_tab = reinterpret_cast<TileEntry*>(smem_raw + (size_t)max_rows * 512);
  for (int j = threadIdx.x; j < tv.n_tab; j += TILE_THREADS) s_tab[j] = tv.tab[j];
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t chunks = batch / 32;
  const int tilesA = tv.NH, tilesB = tv.S;
  const int64_t per_phase = tilesA + tilesB;
  const int64_t total = chunks * per_phase;
  const int64_t ustride = coef_stride ? batch : 1;
  const int64_t ring_elems = tv.n * 32;
  const unsigned long long tgtA = tv.epoch * (unsigned long long)tilesA;
  const unsigned long long tgtB = tv.epoch * (unsigned long long)tilesB;

  for (int64_t item = blockIdx.x; item < total; item += gridDim.x) {
    // decode the item: pass, trajectory chunk, tile
    int pass, tile;
    int64_t g;
    if (item < tilesA) {
      pass = 0; g = 0; tile = (int)item;
    } else {
      const int64_t j = item - tilesA;
      const int64_t p = 1 + j / per_phase;
      const int k = (int)(j % per_phase);
      if (p < chunks && k < tilesA) { pass = 0; g = p; tile = k; }
      else { pass = 1; g = p - 1; tile = p < chunks ? k - tilesA : k; }
    }
    const int rows = pass == 0 ? tv.S : tv.NH;
    const int64_t c0 = g * 32;
    // row of slot s of this tile: pass A: tile * S + s; pass B: s * S + tile
    const int64_t row0 = pass == 0 ? (int64_t)tile * tv.S : tile;
    const int64_t rstep = pass == 0 ? 1 : tv.S;

    // 1. the tile of X -> shared memory (a warp copies one row's 512 B per instruction)
    for (int idx = threadIdx.x; idx < rows * 32; idx += TILE_THREADS) {
      const int s = idx >> 5, j = idx & 31;
      cp_async16(s_x + idx, x + (row0 + s * rstep) * batch + c0 + j);
    }
    // 2. this lane's coefficients (times i for a purely imaginary operator)
    double2 u[NOPS];
#pragma unroll
    for (int l = 0; l < NOPS; ++l) {
      const double2 t = __ldg(coef + (int64_t)l * ustride + (coef_stride ? c0 + lane : 0));
      u[l] = ((tv.imag_ops >> l) & 1u) ? make_double2(-t.y, t.x) : t;
    }
    // 3. dependencies: pass A overwrites the ring slot pass B of chunk g - 3 reads; pass B needs all of A(g)
    if (threadIdx.x == 0) {
      if (pass == 0) {
        if (g >= TILE_RING) wait_counter(tv.doneB + (g - TILE_RING), tgtB);
      } else {
        wait_counter(tv.doneA + g, tgtA);
      }
    }
    cp_async_wait_all();
    __syncthreads();

    double2* tbuf = tv.tring + (g % TILE_RING) * ring_elems;
    const uint4* codes = pass == 0 ? tv.codesA : tv.codesB;
    const int W8 = pass == 0 ? tv.WA8 : tv.WB8;
    double dr = 0.0, di = 0.0, nn = 0.0;

    for (int s = warp; s < rows; s += TILE_WARPS) {
      const int64_t row = row0 + s * rstep;
      const int64_t idx = row * batch + c0 + lane;
      const double2 xown = s_x[s * 32 + lane];
      // epilogue operands requested before the row is decoded
      double2 tv_in = make_double2(0.0, 0.0), yv = tv_in, av = tv_in;
      if (pass == 1) {
        tv_in = ld_cg(tbuf + row * 32 + lane);
        if (EPI == EPI_MUL) {
          if (e.betac.x != 0.0 || e.betac.y != 0.0) yv = ld_noalloc(e.y + idx);
        } else if (EPI == EPI_CHEB_MID || EPI == EPI_CHEB_LAST) {
          yv = ld_noalloc(e.y + idx);
          av = ld_noalloc(e.acc + idx);
        }
      }
      double pr[NOPS], pi[NOPS];
#pragma unroll
      for (int l = 0; l < NOPS; ++l) pr[l] = pi[l] = 0.0;
      const unsigned char* xs_row = reinterpret_cast<const unsigned char*>(s_x) + (size_t)s * 512 + lane * 16;
      const double2* xg_row = x + row * batch + c0 + lane;
      const uint4* cw = codes + row * W8;
      uint4 w_next = make_uint4(0u, 0u, 0u, 0u);
      if (W8 > 0) w_next = __ldg(cw);
      for (int k = 0; k < W8; ++k) {
        const uint4 w = w_next;
        if (k + 1 < W8) w_next = __ldg(cw + k + 1);
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int h = 0; h < 2; ++h) {  // four codes at a time: all loads issued before the first use
          if ((ww[2 * h] | ww[2 * h + 1]) == 0u) continue;  // padding only (warp-uniform)
          TileEntry en[4];
          double2 xv[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t code = (ww[2 * h + (q >> 1)] >> (16 * (q & 1))) & 0xffffu;
            en[q] = s_tab[code];
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (en[q].km & qptile::TILE_KIND_O)
              xv[q] = __ldg(xg_row + (int64_t)en[q].off * batch);
            else
              xv[q] = *reinterpret_cast<const double2*>(xs_row + en[q].off);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) tile_acc<NOPS>(en[q], xv[q], pr, pi);
        }
      }
      if (pass == 0) {  // explicit diagonals
#pragma unroll
        for (int l = 0; l < NOPS; ++l) {
          const double d = __ldg(tv.diag + row * qptile::TILE_MAX_OPS + l);
          pr[l] = fma(d, xown.x, pr[l]);
          pi[l] = fma(d, xown.y, pi[l]);
        }
      }
      double2 hx = tv_in;
#pragma unroll
      for (int l = 0; l < NOPS; ++l) {
        hx.x += u[l].x * pr[l] - u[l].y * pi[l];
        hx.y += u[l].x * pi[l] + u[l].y * pr[l];
      }
      if (pass == 0) st_cg(tbuf + row * 32 + lane, hx);
      else epi_apply<EPI>(e, idx, hx, xown, yv, av, dr, di, nn);
    }
    if (pass == 1 && epi_has_sums(EPI) && e.chk != nullptr) chk_flush(e, c0 + lane, dr, di, nn);

    // every warp is done with the tile (the next item overwrites it) and has issued its stores
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd((pass == 0 ? tv.doneA : tv.doneB) + g, 1ull);
    }
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------

void qp_tile_free(qp_tile_s* t) {
  if (!t) return;
  cudaFree(t->d_tab);
  cudaFree(t->d_codesA);
  cudaFree(t->d_codesB);
  cudaFree(t->d_diag);
  cudaFree(t->d_ring);
  cudaFree(t->d_done);
  delete t;
}

// Builds (once per generator, on first use with a batch that qualifies) the tile format from the
// merged CSR arrays.  Returns QP_OK with gen->tile->ok == false when the generator does not qualify.
static int32_t tile_ensure(qp_gen_t gen) {
  if (gen->tile != nullptr) return QP_OK;
  qp_ctx_t ctx = gen->ctx;
  qp_tile_s* t = new qp_tile_s();
  gen->tile = t;
  if (getenv("QPROP_NO_TILE") && atoi(getenv("QPROP_NO_TILE")) != 0) { t->why = "disabled (QPROP_NO_TILE)"; return QP_OK; }
  if (gen->d_mptr == nullptr || gen->n_ops > qptile::TILE_MAX_OPS) { t->why = "no merged sparse matrix / more than 3 operators"; return QP_OK; }
  if (qptile::choose_split(gen->n) == 0) { t->why = "no two-level split of N"; return QP_OK; }
  const int64_t n = gen->n;
  std::vector<uint32_t> h_ptr((size_t)n + 1);
  QP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  QP_CUDA(ctx, cudaMemcpy(h_ptr.data(), gen->d_mptr, sizeof(uint32_t) * (n + 1), cudaMemcpyDeviceToHost));
  const size_t nnz = h_ptr[(size_t)n];
  std::vector<uint32_t> h_col(nnz ? nnz : 1);
  std::vector<double> h_val(2 * (nnz ? nnz : 1));
  if (nnz) {
    QP_CUDA(ctx, cudaMemcpy(h_col.data(), gen->d_mcolop, sizeof(uint32_t) * nnz, cudaMemcpyDeviceToHost));
    QP_CUDA(ctx, cudaMemcpy(h_val.data(), gen->d_mval, sizeof(double2) * nnz, cudaMemcpyDeviceToHost));
  }
  qptile::TileFormat& f = t->meta;
  if (!qptile::build(f, n, gen->n_ops, h_ptr.data(), h_col.data(), h_val.data())) { t->why = f.why; return QP_OK; }
  const int64_t off_diag = f.n_A + f.n_B + f.n_O;
  if (4 * f.n_O > off_diag) {  // mostly unstructured: the tiles would serve too few of the loads
    t->why = "more than a quarter of the entries straddle the split";
    return QP_OK;
  }
  const int max_rows = std::max(f.S, f.NH);
  if ((size_t)max_rows * 512 + f.table.size() * sizeof(TileEntry) > (size_t)220 * 1024) { t->why = "table too large for shared memory"; return QP_OK; }
  auto up = [&](void** dst, const void* src, size_t bytes) -> cudaError_t {
    cudaError_t e1 = cudaMalloc(dst, bytes ? bytes : 16);
    if (e1 == cudaSuccess && bytes) e1 = cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
    return e1;
  };
  QP_CUDA(ctx, up((void**)&t->d_tab, f.table.data(), f.table.size() * sizeof(TileEntry)));
  QP_CUDA(ctx, up((void**)&t->d_codesA, f.codesA.data(), f.codesA.size() * sizeof(uint16_t)));
  QP_CUDA(ctx, up((void**)&t->d_codesB, f.codesB.data(), f.codesB.size() * sizeof(uint16_t)));
  QP_CUDA(ctx, up((void**)&t->d_diag, f.diag.data(), f.diag.size() * sizeof(double)));
  QP_CUDA(ctx, cudaMalloc(&t->d_ring, sizeof(double2) * (size_t)TILE_RING * (size_t)n * 32));
  t->n_tab = (int)f.table.size();
  // the big host arrays are not needed any more
  std::vector<uint16_t>().swap(f.codesA);
  std::vector<uint16_t>().swap(f.codesB);
  std::vector<double>().swap(f.diag);
  t->ok = true;
  return QP_OK;
}

template <int EPI, int NOPS>
static int32_t tile_launch(qp_gen_t gen, int coef_stride, const double2* x, int64_t batch, const EpiArgs& e) {
  qp_ctx_t ctx = gen->ctx;
  qp_tile_s* t = gen->tile;
  const qptile::TileFormat& f = t->meta;
  const int64_t chunks = batch / 32;
  if (chunks > t->chunk_cap) {  // (re)allocate the completion counters
    QP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(t->d_done);
    t->d_done = nullptr;
    t->chunk_cap = 0;
    t->chunks_cur = 0;
    QP_CUDA(ctx, cudaMalloc(&t->d_done, sizeof(unsigned long long) * 2 * (size_t)chunks));
    t->chunk_cap = chunks;
  }
  if (chunks != t->chunks_cur) {  // another batch size than the counters have been counting with: new epoch
    QP_CUDA(ctx, cudaMemsetAsync(t->d_done, 0, sizeof(unsigned long long) * 2 * (size_t)t->chunk_cap, ctx->stream));
    t->chunks_cur = chunks;
    t->epoch = 0;
  }
  TileView tv;
  tv.tab = t->d_tab;
  tv.n_tab = t->n_tab;
  tv.codesA = t->d_codesA;
  tv.codesB = t->d_codesB;
  tv.diag = t->d_diag;
  tv.WA8 = f.WA / 8;
  tv.WB8 = f.WB / 8;
  tv.S = f.S;
  tv.NH = f.NH;
  tv.n_ops = f.n_ops;
  tv.imag_ops = f.imag_ops;
  tv.n = f.n;
  tv.tring = t->d_ring;
  tv.doneA = t->d_done;
  tv.doneB = t->d_done + t->chunk_cap;
  tv.epoch = ++t->epoch;
  auto kern = k_spmm_tile<EPI, NOPS>;
  const size_t smem = (size_t)std::max(f.S, f.NH) * 512 + (size_t)t->n_tab * sizeof(TileEntry);
  if (!ctx->smem_configured.count((const void*)kern)) {
    QP_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    ctx->smem_configured.insert((const void*)kern);
  }
  const int64_t total = chunks * (int64_t)(f.S + f.NH);
  // all CTAs must be resident (spin-waits between items): one per SM, launched cooperatively
  int per_sm = 0;
  QP_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TILE_THREADS, smem));
  if (per_sm < 1) return qp_fail(ctx, QP_ERR_INTERNAL, "tile kernel does not fit an SM (%zu bytes of shared memory)", smem);
  const int64_t grid = std::min<int64_t>((int64_t)ctx->sm_count * std::min(per_sm, 2), total);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(TILE_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const double2* coef = gen->d_coef;
  QP_CUDA(ctx, cudaLaunchKernelEx(&cfg, kern, tv, coef, coef_stride, batch, x, e));
  QP_LAUNCHED(ctx);
  return QP_OK;
}

template <int EPI>
static int32_t tile_launch_nops(qp_gen_t gen, int coef_stride, const double2* x, int64_t batch, const EpiArgs& e) {
  switch (gen->n_ops) {
    case 1: return tile_launch<EPI, 1>(gen, coef_stride, x, batch, e);
    case 2: return tile_launch<EPI, 2>(gen, coef_stride, x, batch, e);
    default: return tile_launch<EPI, 3>(gen, coef_stride, x, batch, e);
  }
}

// Tries the tiled path; *handled = false when the generator / batch does not qualify (the caller
// then uses the one-pass kernels).
int32_t qp_launch_tile(qp_gen_t gen, int epi, int coef_stride, const double2* x, int64_t batch, const EpiArgs& e,
                       bool* handled) {
  *handled = false;
  if (batch < 32 || batch % 32 != 0 || gen->format == QP_FORMAT_DENSE || gen->format == QP_FORMAT_LR) return QP_OK;
  QP_CHECK(tile_ensure(gen));
  if (!gen->tile->ok) return QP_OK;
  *handled = true;
  switch (epi) {
    case EPI_MUL: return tile_launch_nops<EPI_MUL>(gen, coef_stride, x, batch, e);
    case EPI_CHEB_FIRST: return tile_launch_nops<EPI_CHEB_FIRST>(gen, coef_stride, x, batch, e);
    case EPI_CHEB_MID: return tile_launch_nops<EPI_CHEB_MID>(gen, coef_stride, x, batch, e);
    case EPI_CHEB_LAST: return tile_launch_nops<EPI_CHEB_LAST>(gen, coef_stride, x, batch, e);
    case EPI_CHEB_ONLY: return tile_launch_nops<EPI_CHEB_ONLY>(gen, coef_stride, x, batch, e);
    case EPI_DOT: return tile_launch_nops<EPI_DOT>(gen, coef_stride, x, batch, e);
  }
  return qp_fail(gen->ctx, QP_ERR_INTERNAL, "bad epilogue %d", epi);
}

// what the tiled path of this generator looks like (qp_gen_tile_info)
int32_t qp_tile_info(qp_gen_t gen, int32_t* available, int32_t* S, int32_t* NH, int32_t* n_table, int64_t* entries /*[4]: A, B, O, diag*/) {
  QP_CHECK(qp_ctx_bind(gen->ctx));
  QP_CHECK(tile_ensure(gen));
  const qp_tile_s* t = gen->tile;
  if (available) *available = t->ok ? 1 : 0;
  if (S) *S = t->meta.S;
  if (NH) *NH = t->meta.NH;
  if (n_table) *n_table = t->n_tab;
  if (entries) {
    entries[0] = t->meta.n_A;
    entries[1] = t->meta.n_B;
    entries[2] = t->meta.n_O;
    entries[3] = t->meta.n_diag;
  }
  return QP_OK;
}

extern "C" int32_t qp_gen_tile_info(qp_gen_t gen, int32_t* available, int32_t* split, int32_t* blocks, int32_t* n_table,
                                    int64_t* entries) {
  if (!gen) return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_gen_tile_info: null generator");
  if (gen->format == QP_FORMAT_DENSE || gen->format == QP_FORMAT_LR) {
    if (available) *available = 0;
    return QP_OK;
  }
  return qp_tile_info(gen, available, split, blocks, n_table, entries);
}
