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

using qptile::Entry16;
using qptile::Entry32;
using qptile::N_KINDS;

#ifndef TILE_THREADS_N
#define TILE_THREADS_N 512
#endif
#ifndef TILE_LEAN
#define TILE_LEAN 1
#endif
#ifndef TILE_O_EARLY
#define TILE_O_EARLY 1
#endif
constexpr int TILE_THREADS = TILE_THREADS_N;
constexpr int TILE_WARPS = TILE_THREADS / 32;
constexpr int TILE_RING = 3;

struct TileView {
  const Entry16* tab16;
  const Entry32* tab32;
  int n16, n32;
  const uint4* codes[2];       // per pass: [n][WT] words of eight 16-bit codes, the kinds' lists back to back
  int WT[2];                   // words per row
  int w0[2][N_KINDS];          // first word of a kind's list
  int nw[2][N_KINDS];          // words of a kind's list (0: the generator has no such entries)
  const double* diag;          // [n][4] (3 used; 32-byte rows for 16-byte copies)
  int S, NH, n_ops;
  unsigned imag_ops;
  int64_t n;
  double2* tring;                // [TILE_RING][n][32]
  unsigned long long* doneA;     // [chunks] items of pass A finished (monotonic over launches)
  unsigned long long* doneB;     // [chunks]
  unsigned long long epoch;      // 1-based launch number of this view
  long long spin_limit;          // cycles a dependency wait may spin before the kernel traps (0: no limit)
};

// Small 16-byte tables travel as a kernel parameter (constant bank): a warp-uniform look-up is then
// a constant-cache access instead of a 2-wavefront shared-memory broadcast (the shared-memory pipe is
// what bounds this kernel).
constexpr int TILE_CTAB = 448;  // 7 KB of the 32 KB parameter space
struct ConstTab {
  Entry16 e[TILE_CTAB];
};

struct qp_tile_s {
  bool ok = false;
  std::string why;
  qptile::TileFormat meta;  // host copy without the big arrays (cleared after upload)
  Entry16* d_tab16 = nullptr;
  Entry32* d_tab32 = nullptr;
  uint4* d_codes[2] = {nullptr, nullptr};
  double* d_diag = nullptr;
  double2* d_ring = nullptr;
  unsigned long long* d_done = nullptr;  // [2][chunk capacity]
  int64_t chunk_cap = 0;
  int64_t chunks_cur = 0;  // chunk count the counters have been counting with since their last reset
  unsigned long long epoch = 0;
  int n16 = 0, n32 = 0;
  ConstTab* h_ctab = nullptr;  // host copy of the 16-byte table when it fits the kernel-parameter form
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
__device__ __forceinline__ void wait_counter(const unsigned long long* ctr, unsigned long long target, long long limit) {
  const long long t0 = clock64();
  while (ld_acquire(ctr) < target) {
    __nanosleep(64);
    if (limit > 0 && clock64() - t0 > limit) __trap();  // default 2^33 cycles ~ 4 s; QPROP_TILE_SPIN_LIMIT=0 under sanitizers
  }
}

// shared-memory accesses with 32-bit addresses (no generic-address arithmetic in the entry loops)
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
struct Tab16 {  // an Entry16 / the first half of an Entry32 in registers
  double v;
  int32_t off;
  uint32_t flag;
};
__device__ __forceinline__ Tab16 lds_tab16(uint32_t addr) {
  Tab16 r;
  int lo, hi;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(hi), "=r"(r.off), "=r"(r.flag) : "r"(addr));
  r.v = __hiloint2double(hi, lo);
  return r;
}

// The entry loops.  A list = the words [cw, cw + nw) of one kind in the row's code words (shared
// memory); codes are packed at the front, so the first all-zero group of four ends the list.  Per
// group: table look-ups (warp-uniform broadcasts), then the loads of X, then the multiply-adds; no
// per-entry decisions.  MODE 0: kind S_L (one operator, 2 FMA); MODE 1: kind P (operators NOPS-2 and
// NOPS-1, second value = +-v, 4 FMA); MODE 2: kind G (32-byte entries, every operator); MODE 3:
// kind O (as G, X from global memory at row + off).
template <int NOPS, int MODE, int L, int CT>
__device__ __forceinline__ void tile_list(const ConstTab& ctab, uint32_t cw, const int nw, const uint32_t tab_base, const uint32_t xs_row,
                                          const double2* __restrict__ xg_row, const int64_t batch,
                                          double (&pr)[NOPS], double (&pi)[NOPS]) {
  constexpr uint32_t ESZ = MODE >= 2 ? 32u : 16u;
  for (int k = 0; k < nw; ++k, cw += 16u) {
    const uint4 w = lds_u32x4(cw);
    const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if ((ww[2 * h] | ww[2 * h + 1]) == 0u) return;  // end of the list (warp-uniform)
      uint32_t ta[4];
      Tab16 lo[4];
      double2 hi[4], xv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t code = (q & 1) ? (ww[2 * h + (q >> 1)] >> 16) : (ww[2 * h + (q >> 1)] & 0xffffu);
        ta[q] = tab_base + code * ESZ;
        if (CT && MODE < 2) {
          const Entry16 ce = ctab.e[code];
          lo[q].v = ce.v;
          lo[q].off = ce.off;
          lo[q].flag = ce.neg2;
        } else {
          lo[q] = lds_tab16(ta[q]);
        }
      }
      if (MODE >= 2 && NOPS > 1) {
#pragma unroll
        for (int q = 0; q < 4; ++q) hi[q] = lds_f64x2(ta[q] + 16u);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (MODE == 3) xv[q] = __ldg(xg_row + (int64_t)lo[q].off * batch);
        else xv[q] = lds_f64x2(xs_row + (uint32_t)lo[q].off);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (MODE == 0) {
          pr[L] = fma(lo[q].v, xv[q].x, pr[L]);
          pi[L] = fma(lo[q].v, xv[q].y, pi[L]);
        } else if (MODE == 1) {
          constexpr int P1 = NOPS >= 2 ? NOPS - 2 : 0, P2 = NOPS >= 2 ? NOPS - 1 : 0;
          const double v2 = __hiloint2double(__double2hiint(lo[q].v) ^ (int)lo[q].flag, __double2loint(lo[q].v));
          pr[P1] = fma(lo[q].v, xv[q].x, pr[P1]);
          pi[P1] = fma(lo[q].v, xv[q].y, pi[P1]);
          pr[P2] = fma(v2, xv[q].x, pr[P2]);
          pi[P2] = fma(v2, xv[q].y, pi[P2]);
        } else {
          pr[0] = fma(lo[q].v, xv[q].x, pr[0]);
          pi[0] = fma(lo[q].v, xv[q].y, pi[0]);
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
  }
}

// class-O list after its first two codes (rows with more than two straddling couplings: rare)
template <int NOPS, int CT>
__device__ __forceinline__ void tile_list_rest(const ConstTab& ctab, uint4 first, uint32_t cw, const int nw, const uint32_t tab_base,
                                               const uint32_t xs_row, const double2* __restrict__ xg_row, const int64_t batch,
                                               double (&pr)[NOPS], double (&pi)[NOPS]) {
  const uint32_t rest[7] = {first.y & 0xffffu, first.y >> 16, first.z & 0xffffu, first.z >> 16, first.w & 0xffffu, first.w >> 16, 0u};
  for (int q = 0; q < 6 && rest[q] != 0u; ++q) {
    const uint32_t ta = tab_base + rest[q] * 32u;
    const Tab16 lo = lds_tab16(ta);
    const double2 hi = lds_f64x2(ta + 16u);
    const double2 xv = __ldg(xg_row + (int64_t)lo.off * batch);
    pr[0] = fma(lo.v, xv.x, pr[0]);
    pi[0] = fma(lo.v, xv.y, pi[0]);
    if (NOPS > 1) {
      pr[1] = fma(hi.x, xv.x, pr[1]);
      pi[1] = fma(hi.x, xv.y, pi[1]);
    }
    if (NOPS > 2) {
      pr[2] = fma(hi.y, xv.x, pr[2]);
      pi[2] = fma(hi.y, xv.y, pi[2]);
    }
  }
  if (nw > 1 && first.w != 0u) tile_list<NOPS, 3, 0, CT>(ctab, cw + 16u, nw - 1, tab_base, xs_row, xg_row, batch, pr, pi);
}

struct TileRowArgs {
  int rows, warp, lane;
  int64_t row0, rstep, c0, batch;
  double2* tbuf;
  uint32_t tab16_base, tab32_base, xs_base, codes_base, diag_base;
};

// The rows of one tile, PASS known at compile time (list positions and widths are read once per
// tile, not per row).  Software pipeline over this warp's rows: the epilogue operands of the NEXT row
// (global / L2 loads) are requested before the current row is decoded; everything else a row needs
// -- X, its code words, the tables, the diagonal -- is in shared memory.  All row-dependent global
// addresses advance by constant strides.
template <int EPI, int NOPS, int PASS, int CT>
__device__ __forceinline__ void tile_rows(const TileView& tv, const ConstTab& ctab, const TileRowArgs& ra, const double2* __restrict__ x,
                                          const EpiArgs& e, const double2 (&u)[NOPS], double& dr, double& di, double& nn) {
  const int rows = ra.rows, warp = ra.warp;
  const int64_t batch = ra.batch;
  const int WT = tv.WT[PASS];
  int w0[N_KINDS], nw[N_KINDS];
#pragma unroll
  for (int k = 0; k < N_KINDS; ++k) {
    w0[k] = tv.w0[PASS][k] * 16;
    nw[k] = tv.nw[PASS][k];
  }
  const int64_t row_stride = (int64_t)TILE_WARPS * ra.rstep;  // rows between two of this warp's rows
  const int64_t row_w = ra.row0 + (int64_t)warp * ra.rstep;
  const double2* xg_row = x + row_w * batch + ra.c0 + ra.lane;  // x[row][c0 + lane]
  const int64_t el_stride = row_stride * batch;
  int64_t idx = row_w * batch + ra.c0 + ra.lane;
  double2* t_ptr = ra.tbuf + row_w * 32 + ra.lane;
  const int64_t t_stride = row_stride * 32;

  double2 n_t = make_double2(0.0, 0.0), n_y = n_t, n_a = n_t;
  auto prefetch = [&](const double2* tp, int64_t ix) {
    if (PASS == 1) {
      n_t = ld_cg(tp);
      if (EPI == EPI_MUL) {
        if (e.betac.x != 0.0 || e.betac.y != 0.0) n_y = ld_noalloc(e.y + ix);
      } else if (EPI == EPI_CHEB_MID || EPI == EPI_CHEB_LAST) {
        n_y = ld_noalloc(e.y + ix);
        n_a = ld_noalloc(e.acc + ix);
      }
    }
  };
  if (!TILE_LEAN && warp < rows) prefetch(t_ptr, idx);
  for (int s = warp; s < rows; s += TILE_WARPS) {
    if (TILE_LEAN) prefetch(t_ptr, idx);  // nothing carried from row to row: fewer registers, more warps
    const double2 tv_in = n_t, yv = n_y, av = n_a;
    if (!TILE_LEAN && s + TILE_WARPS < rows) prefetch(t_ptr + t_stride, idx + el_stride);
    const uint32_t xs_row = ra.xs_base + (uint32_t)s * 512u;
    const uint32_t cw = ra.codes_base + (uint32_t)(s * WT) * 16u;
    const double2 xown = lds_f64x2(xs_row);
    double pr[NOPS], pi[NOPS];
#pragma unroll
    for (int l = 0; l < NOPS; ++l) pr[l] = pi[l] = 0.0;
    // class O (pass B): the first two entries' global gathers are issued NOW and consumed after the
    // shared-memory lists, so their L2 latency hides behind the rest of the row
    uint4 ow = make_uint4(0u, 0u, 0u, 0u);
    Tab16 o_lo[2];
    double2 o_hi[2], o_x[2];
    if (TILE_O_EARLY && PASS == 1 && nw[5] > 0) {
      ow = lds_u32x4(cw + w0[5]);
      if (ow.x != 0u) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const uint32_t ta = ra.tab32_base + ((q ? (ow.x >> 16) : (ow.x & 0xffffu)) * 32u);
          o_lo[q] = lds_tab16(ta);
          if (NOPS > 1) o_hi[q] = lds_f64x2(ta + 16u);
          o_x[q] = __ldg(xg_row + (int64_t)o_lo[q].off * batch);
        }
      }
    }
    tile_list<NOPS, 0, 0, CT>(ctab, cw + w0[0], nw[0], ra.tab16_base, xs_row, xg_row, batch, pr, pi);
    if (NOPS > 1) tile_list<NOPS, 0, (NOPS > 1 ? 1 : 0), CT>(ctab, cw + w0[1], nw[1], ra.tab16_base, xs_row, xg_row, batch, pr, pi);
    if (NOPS > 2) tile_list<NOPS, 0, (NOPS > 2 ? 2 : 0), CT>(ctab, cw + w0[2], nw[2], ra.tab16_base, xs_row, xg_row, batch, pr, pi);
    if (NOPS > 1) tile_list<NOPS, 1, 0, CT>(ctab, cw + w0[3], nw[3], ra.tab16_base, xs_row, xg_row, batch, pr, pi);
    tile_list<NOPS, 2, 0, CT>(ctab, cw + w0[4], nw[4], ra.tab32_base, xs_row, xg_row, batch, pr, pi);
    if (PASS == 1 && !TILE_O_EARLY) {
      if (nw[5] > 0) tile_list<NOPS, 3, 0, CT>(ctab, cw + w0[5], nw[5], ra.tab32_base, xs_row, xg_row, batch, pr, pi);
    } else if (PASS == 1) {  // class O: the few couplings that straddle the split, from global memory
      if (ow.x != 0u) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          pr[0] = fma(o_lo[q].v, o_x[q].x, pr[0]);
          pi[0] = fma(o_lo[q].v, o_x[q].y, pi[0]);
          if (NOPS > 1) {
            pr[1] = fma(o_hi[q].x, o_x[q].x, pr[1]);
            pi[1] = fma(o_hi[q].x, o_x[q].y, pi[1]);
          }
          if (NOPS > 2) {
            pr[2] = fma(o_hi[q].y, o_x[q].x, pr[2]);
            pi[2] = fma(o_hi[q].y, o_x[q].y, pi[2]);
          }
        }
        if ((ow.y | ow.z | ow.w) != 0u || nw[5] > 1)  // more than two: the rest of the list (first word without its first two codes)
          tile_list_rest<NOPS, CT>(ctab, ow, cw + w0[5], nw[5], ra.tab32_base, xs_row, xg_row, batch, pr, pi);
      }
    } else {          // explicit diagonals
      const double2 d01 = lds_f64x2(ra.diag_base + (uint32_t)s * 32u);
      pr[0] = fma(d01.x, xown.x, pr[0]);
      pi[0] = fma(d01.x, xown.y, pi[0]);
      if (NOPS > 1) {
        pr[1] = fma(d01.y, xown.x, pr[1]);
        pi[1] = fma(d01.y, xown.y, pi[1]);
      }
      if (NOPS > 2) {
        const double2 d2 = lds_f64x2(ra.diag_base + (uint32_t)s * 32u + 16u);
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
    if (PASS == 0) st_cg(t_ptr, hx);
    else epi_apply<EPI>(e, idx, hx, xown, yv, av, dr, di, nn);
    xg_row += el_stride;
    idx += el_stride;
    t_ptr += t_stride;
  }
}

template <int EPI, int NOPS, int CT>
__global__ void __launch_bounds__(TILE_THREADS, 1)
k_spmm_tile(const __grid_constant__ TileView tv, const __grid_constant__ ConstTab ctab, const double2* __restrict__ coef, int coef_stride, int64_t batch,
            const double2* __restrict__ x, EpiArgs e) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // shared memory: X tile | 16-byte table | 32-byte table | the tile's code words | diagonals (pass A)
  double2* s_x = reinterpret_cast<double2*>(smem_raw);  // [rows][32]
  const int max_rows = tv.S > tv.NH ? tv.S : tv.NH;
  Entry16* s_tab16 = reinterpret_cast<Entry16*>(smem_raw + (size_t)max_rows * 512);
  Entry32* s_tab32 = reinterpret_cast<Entry32*>(s_tab16 + tv.n16);
  uint4* s_codes = reinterpret_cast<uint4*>(s_tab32 + tv.n32);
  const int code_words = tv.S * tv.WT[0] > tv.NH * tv.WT[1] ? tv.S * tv.WT[0] : tv.NH * tv.WT[1];
  double* s_diag = reinterpret_cast<double*>(s_codes + code_words);  // [S][4] (3 used)
  double* s_red = s_diag + (size_t)tv.S * 4;                          // [TILE_WARPS][32][3], deterministic sums only
  for (int j = threadIdx.x; j < tv.n16; j += TILE_THREADS) s_tab16[j] = tv.tab16[j];
  for (int j = threadIdx.x; j < tv.n32; j += TILE_THREADS) s_tab32[j] = tv.tab32[j];
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
  const uint32_t tab16_base = smem_u32(s_tab16), tab32_base = smem_u32(s_tab32);
  const uint32_t xs_base = smem_u32(s_x) + (uint32_t)lane * 16u;
  const uint32_t codes_base = smem_u32(s_codes);
  const uint32_t diag_base = smem_u32(s_diag);

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
    const int WT = tv.WT[pass];

    // 1. the tile of X -> shared memory (a warp copies one row's 512 B per instruction) ...
    for (int idx = threadIdx.x; idx < rows * 32; idx += TILE_THREADS) {
      const int s = idx >> 5, j = idx & 31;
      cp_async16(s_x + idx, x + (row0 + s * rstep) * batch + c0 + j);
    }
    // ... and the code words of its rows (and the diagonals in pass A)
    {
      const uint4* src = tv.codes[pass];
      for (int idx = threadIdx.x; idx < rows * WT; idx += TILE_THREADS) {
        const int s = idx / WT, k = idx - s * WT;
        cp_async16(s_codes + idx, src + (row0 + s * rstep) * WT + k);
      }
      if (pass == 0)
        for (int idx = threadIdx.x; idx < rows * 2; idx += TILE_THREADS)  // 32 bytes per row: d0 d1 | d2 pad
          cp_async16(s_diag + idx * 2, tv.diag + row0 * 4 + idx * 2);
    }
    // ... and ask L2 for the tile of the item after this one (first touch of a pass-A tile comes from
    // HBM: the request is in flight while this item is computed)
    {
      const int64_t nxt = item + gridDim.x;
      if (nxt < total) {
        int npass, ntile;
        int64_t ng;
        if (nxt < tilesA) { npass = 0; ng = 0; ntile = (int)nxt; }
        else {
          const int64_t j = nxt - tilesA;
          const int64_t p = 1 + j / per_phase;
          const int k = (int)(j % per_phase);
          if (p < chunks && k < tilesA) { npass = 0; ng = p; ntile = k; }
          else { npass = 1; ng = p - 1; ntile = p < chunks ? k - tilesA : k; }
        }
        const int nrows = npass == 0 ? tv.S : tv.NH;
        const int64_t nrow0 = npass == 0 ? (int64_t)ntile * tv.S : ntile;
        const int64_t nstep = npass == 0 ? 1 : tv.S;
        for (int idx = threadIdx.x; idx < nrows * 4; idx += TILE_THREADS) {  // 4 lines of 128 B per row
          const int s = idx >> 2, j = idx & 3;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(x + (nrow0 + s * nstep) * batch + ng * 32 + j * 8));
        }
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
        if (g >= TILE_RING) wait_counter(tv.doneB + (g - TILE_RING), tgtB, tv.spin_limit);
      } else {
        wait_counter(tv.doneA + g, tgtA, tv.spin_limit);
      }
    }
    cp_async_wait_all();
    __syncthreads();

    double2* tbuf = tv.tring + (g % TILE_RING) * ring_elems;
    double dr = 0.0, di = 0.0, nn = 0.0;
    TileRowArgs ra;
    ra.rows = rows;
    ra.warp = warp;
    ra.lane = lane;
    ra.row0 = row0;
    ra.rstep = rstep;
    ra.c0 = c0;
    ra.batch = batch;
    ra.tbuf = tbuf;
    ra.tab16_base = tab16_base;
    ra.tab32_base = tab32_base;
    ra.xs_base = xs_base;
    ra.codes_base = codes_base;
    ra.diag_base = diag_base;
    if (pass == 0) tile_rows<EPI, NOPS, 0, CT>(tv, ctab, ra, x, e, u, dr, di, nn);
    else tile_rows<EPI, NOPS, 1, CT>(tv, ctab, ra, x, e, u, dr, di, nn);
    const bool det_sums = pass == 1 && epi_has_sums(EPI) && e.chk != nullptr && e.part != nullptr;
    if (pass == 1 && epi_has_sums(EPI) && e.chk != nullptr) {
      if (det_sums) {  // per-warp sums -> shared memory; warp 0 adds them in a fixed order below
        s_red[(warp * 32 + lane) * 3 + 0] = dr;
        s_red[(warp * 32 + lane) * 3 + 1] = di;
        s_red[(warp * 32 + lane) * 3 + 2] = nn;
      } else {
        chk_flush(e, c0 + lane, dr, di, nn);
      }
    }

    // every warp is done with the tile (the next item overwrites it) and has issued its stores
    __syncthreads();
    if (det_sums && warp == 0) {  // slot = B-tile index: e.part[tile][trajectory][3]
      double s0 = 0.0, s1 = 0.0, s2 = 0.0;
      for (int w2 = 0; w2 < TILE_WARPS; ++w2) {
        s0 += s_red[(w2 * 32 + lane) * 3 + 0];
        s1 += s_red[(w2 * 32 + lane) * 3 + 1];
        s2 += s_red[(w2 * 32 + lane) * 3 + 2];
      }
      double* dst = e.part + ((int64_t)tile * batch + c0 + lane) * 3;
      dst[0] = s0;
      dst[1] = s1;
      dst[2] = s2;
    }
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
static size_t tile_smem_bytes(const qptile::TileFormat& f, size_t n16, size_t n32) {
  const size_t rows = (size_t)std::max(f.S, f.NH);
  const size_t code_words = std::max((size_t)f.S * f.WT[0], (size_t)f.NH * f.WT[1]);
  return rows * 512 + n16 * sizeof(Entry16) + n32 * sizeof(Entry32) + code_words * 16 + (size_t)f.S * 32 +
         (size_t)TILE_WARPS * 32 * 3 * sizeof(double);  // + the per-warp sums of the deterministic reductions
}

void qp_tile_free(qp_tile_s* t) {
  if (!t) return;
  cudaFree(t->d_tab16);
  cudaFree(t->d_tab32);
  cudaFree(t->d_codes[0]);
  cudaFree(t->d_codes[1]);
  cudaFree(t->d_diag);
  cudaFree(t->d_ring);
  cudaFree(t->d_done);
  delete t->h_ctab;
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
  const int64_t off_diag = f.n_A() + f.n_B() + f.n_O();
  if (4 * f.n_O() > off_diag) {  // mostly unstructured: the tiles would serve too few of the loads
    t->why = "more than a quarter of the entries straddle the split";
    return QP_OK;
  }
  if (tile_smem_bytes(f, f.tab16.size(), f.tab32.size()) > (size_t)224 * 1024) { t->why = "tile + table + code words exceed shared memory"; return QP_OK; }
  auto up = [&](void** dst, const void* src, size_t bytes) -> cudaError_t {
    cudaError_t e1 = cudaMalloc(dst, bytes ? bytes : 16);
    if (e1 == cudaSuccess && bytes) e1 = cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
    return e1;
  };
  QP_CUDA(ctx, up((void**)&t->d_tab16, f.tab16.data(), f.tab16.size() * sizeof(Entry16)));
  QP_CUDA(ctx, up((void**)&t->d_tab32, f.tab32.data(), f.tab32.size() * sizeof(Entry32)));
  for (int p = 0; p < 2; ++p) QP_CUDA(ctx, up((void**)&t->d_codes[p], f.codes[p].data(), f.codes[p].size() * sizeof(uint16_t)));
  {
    std::vector<double> d4((size_t)n * 4, 0.0);
    for (int64_t r = 0; r < n; ++r)
      for (int l = 0; l < qptile::TILE_MAX_OPS; ++l) d4[(size_t)r * 4 + l] = f.diag[(size_t)r * qptile::TILE_MAX_OPS + l];
    QP_CUDA(ctx, up((void**)&t->d_diag, d4.data(), d4.size() * sizeof(double)));
  }
  QP_CUDA(ctx, cudaMalloc(&t->d_ring, sizeof(double2) * (size_t)TILE_RING * (size_t)n * 32));
  t->n16 = (int)f.tab16.size();
  t->n32 = (int)f.tab32.size();
  static const int no_ctab = getenv("QPROP_TILE_CTAB") ? atoi(getenv("QPROP_TILE_CTAB")) == 0 : 0;
  if (t->n16 <= TILE_CTAB && !no_ctab) {
    t->h_ctab = new ConstTab();
    memset(t->h_ctab, 0, sizeof(ConstTab));
    memcpy(t->h_ctab->e, f.tab16.data(), sizeof(Entry16) * f.tab16.size());
  }
  // the big host arrays are not needed any more
  std::vector<uint16_t>().swap(f.codes[0]);
  std::vector<uint16_t>().swap(f.codes[1]);
  std::vector<double>().swap(f.diag);
  t->ok = true;
  return QP_OK;
}

template <int EPI, int NOPS, int CT>
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
  tv.tab16 = t->d_tab16;
  tv.tab32 = t->d_tab32;
  tv.n16 = t->n16;
  tv.n32 = t->n32;
  for (int p = 0; p < 2; ++p) {
    tv.codes[p] = t->d_codes[p];
    tv.WT[p] = f.WT[p];
    for (int k = 0; k < N_KINDS; ++k) {
      tv.w0[p][k] = f.word0[p][k];
      tv.nw[p][k] = f.W[p][k] / 8;
    }
  }
  tv.diag = t->d_diag;
  tv.S = f.S;
  tv.NH = f.NH;
  tv.n_ops = f.n_ops;
  tv.imag_ops = f.imag_ops;
  tv.n = f.n;
  tv.tring = t->d_ring;
  tv.doneA = t->d_done;
  tv.doneB = t->d_done + t->chunk_cap;
  tv.epoch = ++t->epoch;
  static const long long spin_limit = getenv("QPROP_TILE_SPIN_LIMIT") ? atoll(getenv("QPROP_TILE_SPIN_LIMIT")) : (1ll << 33);
  tv.spin_limit = spin_limit;
  auto kern = k_spmm_tile<EPI, NOPS, CT>;
  static const ConstTab empty_tab = {};
  const ConstTab& ctab = CT ? *t->h_ctab : empty_tab;
  const size_t smem = tile_smem_bytes(f, (size_t)t->n16, (size_t)t->n32);
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
  // deterministic sums (fused expectation value, normalization check): one slot per B tile and
  // trajectory, added up in a fixed order afterwards
  EpiArgs e2 = e;
  e2.part = nullptr;
  static const int atomic = getenv("QPROP_ATOMIC_SUMS") ? atoi(getenv("QPROP_ATOMIC_SUMS")) : 0;
  const bool det = epi_has_sums(EPI) && e.chk != nullptr && !atomic;
  if (det) {
    const size_t doubles = (size_t)f.S * (size_t)batch * 3;
    QP_CHECK(qp_ctx_reserve_part(ctx, doubles));
    e2.part = ctx->d_part;  // every (tile, trajectory) slot is written by exactly one item: no memset needed
  }
  QP_CUDA(ctx, cudaLaunchKernelEx(&cfg, kern, tv, ctab, coef, coef_stride, batch, x, e2));
  QP_LAUNCHED(ctx);
  if (det) QP_CHECK(qp_part_reduce(ctx, e2.part, f.S, batch, e2.chk));
  return QP_OK;
}

template <int EPI>
static int32_t tile_launch_nops(qp_gen_t gen, int coef_stride, const double2* x, int64_t batch, const EpiArgs& e) {
  const bool ct = gen->tile->h_ctab != nullptr;
  switch (gen->n_ops) {
    case 1: return ct ? tile_launch<EPI, 1, 1>(gen, coef_stride, x, batch, e) : tile_launch<EPI, 1, 0>(gen, coef_stride, x, batch, e);
    case 2: return ct ? tile_launch<EPI, 2, 1>(gen, coef_stride, x, batch, e) : tile_launch<EPI, 2, 0>(gen, coef_stride, x, batch, e);
    default: return ct ? tile_launch<EPI, 3, 1>(gen, coef_stride, x, batch, e) : tile_launch<EPI, 3, 0>(gen, coef_stride, x, batch, e);
  }
}

// Tries the tiled path; *handled = false when the generator / batch does not qualify (the caller
// then uses the one-pass kernels).
int32_t qp_launch_tile(qp_gen_t gen, int epi, int coef_stride, const double2* x, int64_t batch, const EpiArgs& e,
                       bool* handled) {
  *handled = false;
  if (batch < 32 || batch % 32 != 0 || gen->format == QP_FORMAT_DENSE || gen->format == QP_FORMAT_LR || gen->d_mptr == nullptr) return QP_OK;
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
  if (n_table) *n_table = t->n16 + t->n32;
  if (entries) {
    entries[0] = t->meta.n_A();
    entries[1] = t->meta.n_B();
    entries[2] = t->meta.n_O();
    entries[3] = t->meta.n_diag;
  }
  return QP_OK;
}

extern "C" int32_t qp_gen_tile_info(qp_gen_t gen, int32_t* available, int32_t* split, int32_t* blocks, int32_t* n_table,
                                    int64_t* entries) {
  if (!gen) return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_gen_tile_info: null generator");
  if (gen->format == QP_FORMAT_DENSE || gen->format == QP_FORMAT_LR || gen->d_mptr == nullptr) {
    if (available) *available = 0;
    return QP_OK;
  }
  return qp_tile_info(gen, available, split, blocks, n_table, entries);
}
