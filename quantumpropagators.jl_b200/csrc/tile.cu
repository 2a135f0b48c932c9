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
  const uint4* codesO;  // [n][WO8] class-O entries (gathered from global memory in pass B)
  const double* diag;   // [n][4] (3 used; 32-byte rows for 16-byte copies)
  int WA8, WB8, WO8;
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
  uint4* d_codesO = nullptr;
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

// shared-memory accesses with 32-bit addresses (no generic-address arithmetic in the entry loop)
__device__ __forceinline__ double2 lds_f64x2(uint32_t addr) {
  double2 r;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(addr));
  return r;
}
__device__ __forceinline__ uint4 lds_u32x4(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}
struct TabLo {  // first 16 bytes of a TileEntry
  double v0;
  int32_t off;
  uint32_t km;
};
__device__ __forceinline__ TabLo lds_tablo(uint32_t addr) {
  TabLo r;
  int lo, hi;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(hi), "=r"(r.off), "=r"(r.km) : "r"(addr));
  r.v0 = __hiloint2double(hi, lo);
  return r;
}

// One 16-byte word = eight 16-bit codes of a row, four at a time: table look-ups (warp-uniform
// broadcasts), then the four loads of X, then the multiply-adds -- unconditionally for every
// operator (an absent operator has v = 0), no per-entry branches.  GLOBAL: class-O entries, X from
// global memory at row + off.
template <int NOPS, bool GLOBAL>
__device__ __forceinline__ void tile_word(const uint4 w, const uint32_t tab_base, const uint32_t xs_row,
                                          const double2* __restrict__ xg_row, const int64_t batch,
                                          double (&pr)[NOPS], double (&pi)[NOPS]) {
  const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if ((ww[2 * h] | ww[2 * h + 1]) == 0u) continue;  // padding only (warp-uniform)
    uint32_t ta[4];
    TabLo lo[4];
    double2 hi[4], xv[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t code = (q & 1) ? (ww[2 * h + (q >> 1)] >> 16) : (ww[2 * h + (q >> 1)] & 0xffffu);
      ta[q] = tab_base + code * 32u;
      lo[q] = lds_tablo(ta[q]);
    }
    if (NOPS > 1) {
#pragma unroll
      for (int q = 0; q < 4; ++q) hi[q] = lds_f64x2(ta[q] + 16u);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (GLOBAL) xv[q] = __ldg(xg_row + (int64_t)lo[q].off * batch);
      else xv[q] = lds_f64x2(xs_row + (uint32_t)lo[q].off);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      pr[0] = fma(lo[q].v0, xv[q].x, pr[0]);
      pi[0] = fma(lo[q].v0, xv[q].y, pi[0]);
      if (NOPS > 1) {
        pr[1] = fma(hi[q].x, xv[q].x, pr[1]);
        pi[1] = fma(hi[q].x, xv[q].y, pi[1]);
      }
      if (NOPS > 2) {
        pr[2] = fma(hi[q].y, xv[q].x, pr[2]);
        pi[2] = fma(hi[q].y, xv[q].y, pi[2]);
      }
    }
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
  // per tile: the rows' code words (pass A: WA8 words; pass B: WB8 + WO8 words) and, in pass A, the diagonals
  uint4* s_codes = reinterpret_cast<uint4*>(smem_raw + (size_t)max_rows * 512 + (((size_t)tv.n_tab * sizeof(TileEntry) + 15) & ~(size_t)15));
  const int cwA = tv.WA8, cwB = tv.WB8 + tv.WO8;
  const int code_words = tv.S * cwA > tv.NH * cwB ? tv.S * cwA : tv.NH * cwB;
  double* s_diag = reinterpret_cast<double*>(s_codes + code_words);  // [S][4] (3 used)
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
    // ... and the code words of its rows (and the diagonals in pass A)
    {
      const int W8 = pass == 0 ? tv.WA8 : tv.WB8;
      const int WT = pass == 0 ? cwA : cwB;
      const uint4* src = pass == 0 ? tv.codesA : tv.codesB;
      for (int idx = threadIdx.x; idx < rows * W8; idx += TILE_THREADS) {
        const int s = idx / W8, k = idx - s * W8;
        cp_async16(s_codes + s * WT + k, src + (row0 + s * rstep) * W8 + k);
      }
      if (pass == 1)
        for (int idx = threadIdx.x; idx < rows * tv.WO8; idx += TILE_THREADS) {
          const int s = idx / tv.WO8, k = idx - s * tv.WO8;
          cp_async16(s_codes + s * WT + W8 + k, tv.codesO + (row0 + s * rstep) * tv.WO8 + k);
        }
      if (pass == 0)
        for (int idx = threadIdx.x; idx < rows * 2; idx += TILE_THREADS) {  // 32 bytes per row: d0 d1 | d2 pad
          const int s = idx >> 1, k = idx & 1;
          cp_async16(s_diag + s * 4 + 2 * k, tv.diag + (row0 + s) * 4 + 2 * k);
        }
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
    const uint32_t tab_base = smem_u32(s_tab);
    const uint32_t xs_base = smem_u32(s_x) + (uint32_t)lane * 16u;
    const int W8 = pass == 0 ? tv.WA8 : tv.WB8;
    double dr = 0.0, di = 0.0, nn = 0.0;

    // Software pipeline over this warp's rows: the epilogue operands of the NEXT row (global / L2
    // loads) are requested before the current row is decoded; everything else a row needs -- X,
    // its code words, the table, the diagonal -- is in shared memory.  All row-dependent global
    // addresses advance by constant strides.
    const int64_t row_stride = (int64_t)TILE_WARPS * rstep;            // rows between two of this warp's rows
    const int64_t row_w = row0 + (int64_t)warp * rstep;
    const double2* xg_row = x + row_w * batch + c0 + lane;             // x[row][c0 + lane]
    const int64_t el_stride = row_stride * batch;
    int64_t idx = row_w * batch + c0 + lane;
    const double2* t_ptr = tbuf + row_w * 32 + lane;
    const int64_t t_stride = row_stride * 32;
    const int WT = pass == 0 ? cwA : cwB;
    const uint32_t codes_base = smem_u32(s_codes);
    const uint32_t diag_base = smem_u32(s_diag);

    double2 n_t = make_double2(0.0, 0.0), n_y = n_t, n_a = n_t;
    auto prefetch = [&](const double2* tp, int64_t ix) {
      if (pass == 1) {
        n_t = ld_cg(tp);
        if (EPI == EPI_MUL) {
          if (e.betac.x != 0.0 || e.betac.y != 0.0) n_y = ld_noalloc(e.y + ix);
        } else if (EPI == EPI_CHEB_MID || EPI == EPI_CHEB_LAST) {
          n_y = ld_noalloc(e.y + ix);
          n_a = ld_noalloc(e.acc + ix);
        }
      }
    };
    if (warp < rows) prefetch(t_ptr, idx);
    for (int s = warp; s < rows; s += TILE_WARPS) {
      const double2 tv_in = n_t, yv = n_y, av = n_a;
      if (s + TILE_WARPS < rows) prefetch(t_ptr + t_stride, idx + el_stride);
      const uint32_t xs_row = xs_base + (uint32_t)s * 512u;
      const uint32_t cw = codes_base + (uint32_t)(s * WT) * 16u;
      const double2 xown = lds_f64x2(xs_row);
      double pr[NOPS], pi[NOPS];
#pragma unroll
      for (int l = 0; l < NOPS; ++l) pr[l] = pi[l] = 0.0;
      for (int k = 0; k < W8; ++k) tile_word<NOPS, false>(lds_u32x4(cw + 16u * k), tab_base, xs_row, xg_row, batch, pr, pi);
      if (pass == 1) {  // class O: the few couplings that straddle the split, from global memory
        for (int k = 0; k < tv.WO8; ++k)
          tile_word<NOPS, true>(lds_u32x4(cw + 16u * (W8 + k)), tab_base, xs_row, xg_row, batch, pr, pi);
      } else {          // explicit diagonals
        const double2 d01 = lds_f64x2(diag_base + (uint32_t)s * 32u);
        pr[0] = fma(d01.x, xown.x, pr[0]);
        pi[0] = fma(d01.x, xown.y, pi[0]);
        if (NOPS > 1) {
          pr[1] = fma(d01.y, xown.x, pr[1]);
          pi[1] = fma(d01.y, xown.y, pi[1]);
        }
        if (NOPS > 2) {
          const double2 d2 = lds_f64x2(diag_base + (uint32_t)s * 32u + 16u);
          pr[2] = fma(d2.x, xown.x, pr[2]);
          pi[2] = fma(d2.x, xown.y, pi[2]);
        }
      }
      double2 hx = tv_in;
#pragma unroll
      for (int l = 0; l < NOPS; ++l) {
        hx.x += u[l].x * pr[l] - u[l].y * pi[l];
        hx.y += u[l].x * pi[l] + u[l].y * pr[l];
      }
      if (pass == 0) st_cg(const_cast<double2*>(t_ptr), hx);
      else epi_apply<EPI>(e, idx, hx, xown, yv, av, dr, di, nn);
      xg_row += el_stride;
      idx += el_stride;
      t_ptr += t_stride;
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

// shared memory of one CTA: X tile + table + the tile's code words + the diagonals (pass A)
static size_t tile_smem_bytes(const qptile::TileFormat& f, int n_tab) {
  const size_t rows = (size_t)std::max(f.S, f.NH);
  const size_t code_words = std::max((size_t)f.S * (f.WA / 8), (size_t)f.NH * (f.WB / 8 + f.WO / 8));
  return rows * 512 + (((size_t)n_tab * sizeof(TileEntry) + 15) & ~(size_t)15) + code_words * 16 + (size_t)f.S * 32;
}

void qp_tile_free(qp_tile_s* t) {
  if (!t) return;
  cudaFree(t->d_tab);
  cudaFree(t->d_codesA);
  cudaFree(t->d_codesB);
  cudaFree(t->d_codesO);
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
  if (tile_smem_bytes(f, (int)f.table.size()) > (size_t)224 * 1024) { t->why = "tile + table + code words exceed shared memory"; return QP_OK; }
  auto up = [&](void** dst, const void* src, size_t bytes) -> cudaError_t {
    cudaError_t e1 = cudaMalloc(dst, bytes ? bytes : 16);
    if (e1 == cudaSuccess && bytes) e1 = cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
    return e1;
  };
  QP_CUDA(ctx, up((void**)&t->d_tab, f.table.data(), f.table.size() * sizeof(TileEntry)));
  QP_CUDA(ctx, up((void**)&t->d_codesA, f.codesA.data(), f.codesA.size() * sizeof(uint16_t)));
  QP_CUDA(ctx, up((void**)&t->d_codesB, f.codesB.data(), f.codesB.size() * sizeof(uint16_t)));
  QP_CUDA(ctx, up((void**)&t->d_codesO, f.codesO.data(), f.codesO.size() * sizeof(uint16_t)));
  {
    std::vector<double> d4((size_t)n * 4, 0.0);
    for (int64_t r = 0; r < n; ++r)
      for (int l = 0; l < qptile::TILE_MAX_OPS; ++l) d4[(size_t)r * 4 + l] = f.diag[(size_t)r * qptile::TILE_MAX_OPS + l];
    QP_CUDA(ctx, up((void**)&t->d_diag, d4.data(), d4.size() * sizeof(double)));
  }
  QP_CUDA(ctx, cudaMalloc(&t->d_ring, sizeof(double2) * (size_t)TILE_RING * (size_t)n * 32));
  t->n_tab = (int)f.table.size();
  // the big host arrays are not needed any more
  std::vector<uint16_t>().swap(f.codesA);
  std::vector<uint16_t>().swap(f.codesB);
  std::vector<uint16_t>().swap(f.codesO);
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
  tv.codesO = t->d_codesO;
  tv.diag = t->d_diag;
  tv.WA8 = f.WA / 8;
  tv.WB8 = f.WB / 8;
  tv.WO8 = f.WO / 8;
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
  const size_t smem = tile_smem_bytes(f, t->n_tab);
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
