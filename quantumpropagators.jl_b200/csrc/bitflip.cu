// Bit-flip (XOR-stencil) form of a generator, single states (QP_FORMAT_BITFLIP).
//
// Spin-1/2 generators are very often "diagonal operators + sums of bit flips".  The transverse-field Ising
// chain of BASELINE config 2 is H0 = -J sum Z_i Z_{i+1} (diagonal), H1 = sum X_i (row r couples to the 20
// columns r XOR 2^i, every entry the SAME value), H2 = sum Z_i (diagonal); QAOA mixers look the same; the
// Liouvillian of config 4 is a diagonal + the flips of the two commutator halves + one CONDITIONAL flip per
// decay channel (sigma^-_k (x) sigma^-_k: mask 2^k | 2^(k+n), present on the rows whose two bits are 0).  For
// such an operator nothing about an off-diagonal entry depends on the row beyond a sub-cube condition:
//
//     column = row XOR mask_t,   value = v_t  if (row AND cmask_t) == cval_t,  absent otherwise.
//
// So there is no matrix stream at all -- the <= 64 (mask, value, condition, operator) tuples travel as a
// kernel parameter (constant bank), the diagonals are explicit vectors or 16-bit codes -- and a fused term is
//
//     (H x)[r] = sum_l u_l ( d_l[r] x[r] + sum_{t in l, cond_t(r)} v_t x[r XOR mask_t] )
//
// with no code words, no table look-ups and no per-entry decode: masks below 32 are warp shuffles of the
// lane's own x[r], every other term is one coalesced 512 B load per warp.  The structure is DETECTED at
// qp_gen_create from the uploaded sparse matrices (k_bf_scan: per operator the set of distinct row XOR column
// values, and per value the uniform entry and the exact sub-cube of rows that carry it), never assumed;
// anything else keeps the dictionary / SELL / CSR formats.  Same fused epilogues as every other kernel.
//
// Replaces: mul!(C, A::Operator, B, alpha, beta), src/generators.jl:634-645, inside the Chebyshev term
// of src/cheby.jl:186-209 (config 2 of BASELINE.json) and the Arnoldi matvec of src/arnoldi.jl:72 (config 4).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>

#include "spmv.cuh"

constexpr int BF_MAX_TERMS = 64;   // load terms (incl. padding)
constexpr int BF_STATIC = 32;      // load terms at compile-time positions of the constant bank
constexpr int BF_MAX_LOW = 8;      // shuffle terms
constexpr int BF_MAX_DIAG = 4;
constexpr int BF_MAX_JOINT = 2048;     // entries of the folded diagonal table (shared memory)
constexpr int BF_DIAG_DISTINCT = 256;  // distinct values per coded diagonal
constexpr int BF_HASH_CAP = 1024;
constexpr int BF_SET_CAP = 128;        // hash slots for the distinct masks of one operator
#ifndef BF_SHUF_EARLY
#define BF_SHUF_EARLY 0  // 1: shuffle terms in the shadow of the first batch of gathers -- measured 2400 instead of 2510 prop_step!/s
#endif
constexpr int SHUF_EARLY = BF_SHUF_EARLY;

struct BitflipView {
  int64_t n;
  // terms served by loads (any mask; [0, n_high)) and by warp shuffles (mask < 32; [0, n_low))
  int n_high, n_low;
  uint32_t mask[BF_MAX_TERMS], cmask[BF_MAX_TERMS], cval[BF_MAX_TERMS];
  uint32_t lmask[BF_MAX_LOW], lcmask[BF_MAX_LOW], lcval[BF_MAX_LOW];
  // values and operator indices (device arrays; [0, BF_MAX_TERMS) load terms, then the shuffle terms): only the
  // launches that take their coefficients from device memory read them
  const double2* tval;
  const int* top;
  int n_diag;
  const double* diag_r[BF_MAX_DIAG];   // real diagonal (or nullptr)
  const double2* diag_c[BF_MAX_DIAG];  // complex diagonal (or nullptr)
  int diag_op[BF_MAX_DIAG];
  // coded diagonals (all diagonals real with few distinct values each): one 16-bit code per row whose mixed-radix
  // digits index the sorted tables of distinct values; the kernel prologue folds this launch's coefficients into
  // one shared-memory table, so the diagonals cost 2 B and one look-up per row
  const uint16_t* dcode;   // [n] or nullptr
  const double* dtab;      // the tables, one after the other
  int n_joint;
  int dcount[BF_MAX_DIAG], dstride[BF_MAX_DIAG], doff[BF_MAX_DIAG];
  // CK >= 1 launches: the products coefficient x value (and the diagonals' coefficients), computed on the host
  // from the host copy of the coefficients -- they reach the DFMAs straight from the constant bank
  double cre[BF_MAX_TERMS], cim[BF_MAX_TERMS];
  double lcre[BF_MAX_LOW], lcim[BF_MAX_LOW];
  double2 cd[BF_MAX_DIAG];
};

__device__ __forceinline__ double ld_stream_f64(const double* p) {
  double r;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
  return r;
}

struct qp_bitflip_s {
  BitflipView view;
  std::vector<void*> owned;  // device arrays
  std::vector<double2> h_val;  // host copies of the term values / operators, same layout as tval / top
  std::vector<int> h_op;
  bool values_real = false;  // every term value is real
  bool diags_real = false;   // every diagonal is real
  bool cond = false;         // some term is conditional
};

void qp_bitflip_free(qp_bitflip_s* b) {
  if (!b) return;
  for (void* p : b->owned) cudaFree(p);
  delete b;
}

// ---------------------------------------------------------------------------------------
// detection
// ---------------------------------------------------------------------------------------

// Per distinct mask = row XOR column of one operator: how many rows carry it, which row bits are constant
// among them (ones = AND of the rows, zeros = AND of their complements), the value of the first entry seen
// and whether every other entry has the same one.
struct MaskSet {
  unsigned long long key[BF_SET_CAP];    // mask | 1 << 32; 0 = empty
  unsigned int ones[BF_SET_CAP], zeros[BF_SET_CAP];
  unsigned long long count[BF_SET_CAP];
  unsigned long long vre[BF_SET_CAP], vim[BF_SET_CAP];  // bit patterns; BF_UNSET = none yet
  int bad;       // != 0: a mask with two different values
  int overflow;  // != 0: more distinct masks than slots
};
constexpr unsigned long long BF_UNSET = 0x7ff8dead0badbeefull;

__device__ __forceinline__ int bf_set_find(MaskSet* s, unsigned long long key) {
  unsigned h = (unsigned)((key * 0x9E3779B97F4A7C15ull) >> 57) % BF_SET_CAP;
  for (int probes = 0; probes < BF_SET_CAP; ++probes) {
    unsigned long long cur = s->key[h];
    if (cur == key) return (int)h;
    if (cur == 0ull) {
      cur = atomicCAS(&s->key[h], 0ull, key);
      if (cur == 0ull || cur == key) return (int)h;
    }
    h = (h + 1) % BF_SET_CAP;
  }
  return -1;
}

__device__ __forceinline__ void bf_set_value(MaskSet* s, int h, unsigned long long bre, unsigned long long bim) {
  const unsigned long long o1 = atomicCAS(&s->vre[h], BF_UNSET, bre);
  const unsigned long long o2 = atomicCAS(&s->vim[h], BF_UNSET, bim);
  if ((o1 != BF_UNSET && o1 != bre) || (o2 != BF_UNSET && o2 != bim)) s->bad = 1;
}

__device__ __forceinline__ void bf_set_clear(MaskSet* s) {
  for (int t = threadIdx.x; t < BF_SET_CAP; t += blockDim.x) {
    s->key[t] = 0ull;
    s->ones[t] = s->zeros[t] = 0xffffffffu;
    s->count[t] = 0ull;
    s->vre[t] = s->vim[t] = BF_UNSET;
  }
  if (threadIdx.x == 0) s->bad = s->overflow = 0;
}

__global__ void k_bf_set_init(MaskSet* g) { bf_set_clear(g); }

// one block per chunk of rows: statistics in shared memory first, merged into the global set at the end
__global__ void __launch_bounds__(256)
k_bf_scan(const uint32_t* __restrict__ ptr, const uint32_t* __restrict__ col, const double2* __restrict__ val, int64_t n,
          int rows_per_block, MaskSet* g) {
  __shared__ MaskSet s;
  bf_set_clear(&s);
  __syncthreads();
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  for (int64_t r = r0 + threadIdx.x; r < r0 + rows_per_block && r < n; r += blockDim.x) {
    const uint32_t p0 = ptr[r], p1 = ptr[r + 1];
    for (uint32_t p = p0; p < p1; ++p) {
      const uint32_t m = col[p] ^ (uint32_t)r;
      const int h = bf_set_find(&s, (unsigned long long)m | (1ull << 32));
      if (h < 0) {
        s.overflow = 1;
        continue;
      }
      atomicAnd(&s.ones[h], (unsigned int)r);
      atomicAnd(&s.zeros[h], ~(unsigned int)r);
      atomicAdd(&s.count[h], 1ull);
      if (m != 0u) {  // the diagonal may take any values
        const double2 v = val[p];
        bf_set_value(&s, h, (unsigned long long)__double_as_longlong(v.x + 0.0), (unsigned long long)__double_as_longlong(v.y + 0.0));
      }
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < BF_SET_CAP; t += blockDim.x) {
    if (s.key[t] == 0ull) continue;
    const int h = bf_set_find(g, s.key[t]);
    if (h < 0) {
      g->overflow = 1;
      continue;
    }
    atomicAnd(&g->ones[h], s.ones[t]);
    atomicAnd(&g->zeros[h], s.zeros[t]);
    atomicAdd(&g->count[h], s.count[t]);
    if (s.vre[t] != BF_UNSET) bf_set_value(g, h, s.vre[t], s.vim[t]);
  }
  if (threadIdx.x == 0) {
    if (s.bad) g->bad = 1;
    if (s.overflow) g->overflow = 1;
  }
}

// diag[r] = the diagonal entry of row r (0 if absent); flags[1] = 1 if any imaginary part is non-zero
__global__ void k_bf_extract_diag(const uint32_t* __restrict__ ptr, const uint32_t* __restrict__ col,
                                  const double2* __restrict__ val, int64_t n, double2* __restrict__ dc,
                                  double* __restrict__ dr, int* flags) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    double2 v = make_double2(0.0, 0.0);
    for (uint32_t p = ptr[r]; p < ptr[r + 1]; ++p)
      if (col[p] == (uint32_t)r) v = val[p];
    dc[r] = v;
    dr[r] = v.x;
    if (v.y != 0.0) flags[1] = 1;
  }
}

// distinct values of a real diagonal: open addressing on the bit patterns (-0.0 counts as 0.0)
__global__ void k_bf_distinct(const double* __restrict__ d, int64_t n, unsigned long long* keys, int* flags) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    const double val = d[r] + 0.0;
    const unsigned long long key = (unsigned long long)__double_as_longlong(val);
    unsigned h = (unsigned)((key * 0x9E3779B97F4A7C15ull) >> 54) % BF_HASH_CAP;
    int probes = 0;
    for (; probes < BF_HASH_CAP; ++probes) {
      unsigned long long cur = keys[h];
      if (cur == key) break;
      if (cur == ~0ull) {
        cur = atomicCAS(&keys[h], ~0ull, key);
        if (cur == ~0ull || cur == key) break;
      }
      h = (h + 1) % BF_HASH_CAP;
    }
    if (probes == BF_HASH_CAP) flags[0] = 1;
  }
}

struct DiagTabs {
  int n_diag;
  const double* d[BF_MAX_DIAG];
  int count[BF_MAX_DIAG], stride[BF_MAX_DIAG], off[BF_MAX_DIAG];
};

// code[r] = sum_i digit_i(r) * stride_i, digit_i = position of d_i[r] in the sorted table i
__global__ void k_bf_encode(DiagTabs t, const double* __restrict__ tab, int64_t n, uint16_t* __restrict__ code, int* flags) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    int c = 0;
    for (int i = 0; i < t.n_diag; ++i) {
      const double val = t.d[i][r] + 0.0;
      int lo = 0, hi = t.count[i] - 1;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (tab[t.off[i] + mid] < val) lo = mid + 1;
        else hi = mid;
      }
      if (tab[t.off[i] + lo] != val) flags[0] = 1;
      c += lo * t.stride[i];
    }
    code[r] = (uint16_t)c;
  }
}

struct BfTerm {
  uint32_t mask, cmask, cval;
  double2 val;
  int op;
};
struct BfCollect {
  std::vector<BfTerm> terms;
  bool values_real = true, diags_real = true;
};

// statistics of one CSR matrix (n rows) -> S
static int32_t bf_scan_matrix(qp_ctx_t ctx, const uint32_t* ptr, const uint32_t* col, const double2* val, int64_t n, MaskSet* d_set,
                              MaskSet& S) {
  constexpr int ROWS_PER_BLOCK = 2048;
  k_bf_set_init<<<1, 128, 0, ctx->stream>>>(d_set);
  k_bf_scan<<<(unsigned)((n + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK), 256, 0, ctx->stream>>>(ptr, col, val, n, ROWS_PER_BLOCK, d_set);
  ctx->launches += 2;
  cudaMemcpyAsync(&S, d_set, sizeof(MaskSet), cudaMemcpyDeviceToHost, ctx->stream);
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return qp_fail(ctx, QP_ERR_CUDA, "bit-flip form: operator scan failed");
  return QP_OK;
}

// the flips of a scanned matrix (op = 0); false if the matrix is not a diagonal plus (conditional) uniform flips
static bool bf_terms_from_set(const MaskSet& S, int64_t n, std::vector<BfTerm>& out, bool& has_diag) {
  const bool pow2 = (n & (n - 1)) == 0;
  has_diag = false;
  if (S.bad || S.overflow) return false;
  for (int t = 0; t < BF_SET_CAP; ++t) {
    if (S.key[t] == 0ull) continue;
    const uint32_t m = (uint32_t)(S.key[t] & 0xffffffffull);
    if (m == 0u) {
      has_diag = true;
      continue;
    }
    // rows carrying the entry: all of them, or exactly the sub-cube (row AND cmask) == cval
    uint32_t cmask = 0u, cval = 0u;
    if ((int64_t)S.count[t] != n) {
      if (!pow2) return false;
      cmask = (S.ones[t] | S.zeros[t]) & (uint32_t)(n - 1);
      cval = S.ones[t] & (uint32_t)(n - 1);
      if ((int64_t)S.count[t] != (n >> __builtin_popcount(cmask))) return false;
    }
    BfTerm T;
    T.mask = m;
    T.cmask = cmask;
    T.cval = cval;
    memcpy(&T.val.x, &S.vre[t], 8);
    memcpy(&T.val.y, &S.vim[t], 8);
    T.op = 0;
    out.push_back(T);
  }
  // the slot order of the hash set depends on the race of the insertions: sort
  std::sort(out.begin(), out.end(), [](const BfTerm& a, const BfTerm& b) {
    if (a.mask != b.mask) return a.mask < b.mask;
    if (a.cmask != b.cmask) return a.cmask < b.cmask;
    return a.cval < b.cval;
  });
  return true;
}

// sparse operators: every operator a diagonal plus (conditional) uniform flips
static int32_t bf_collect_sparse(qp_gen_t g, qp_bitflip_s* B, MaskSet* d_set, int* d_flags, BfCollect& C, bool* ok) {
  *ok = false;
  qp_ctx_t ctx = g->ctx;
  const int64_t n = g->n;
  BitflipView& v = B->view;
  const unsigned blocks = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 16);
  std::vector<MaskSet> h_set(1);
  for (int l = 0; l < g->n_ops; ++l) {
    qp_op_t op = g->ops[l];
    if (op->dense || op->leftright) return QP_OK;
    if (op->nnz == 0) continue;  // an all-zero operator contributes nothing
    QP_CHECK(bf_scan_matrix(ctx, op->d_ptr, op->d_col, op->d_val, n, d_set, h_set[0]));
    std::vector<BfTerm> flips;
    bool has_diag = false;
    if (!bf_terms_from_set(h_set[0], n, flips, has_diag)) return QP_OK;
    for (BfTerm& T : flips) {
      T.op = l;
      if (T.val.y != 0.0) C.values_real = false;
      C.terms.push_back(T);
    }
    if ((int)C.terms.size() > BF_MAX_TERMS + BF_MAX_LOW - 4) return QP_OK;
    if (has_diag) {
      if (v.n_diag >= BF_MAX_DIAG) return QP_OK;
      double2* dc = nullptr;
      double* dr = nullptr;
      if (cudaMalloc(&dc, sizeof(double2) * n) != cudaSuccess || cudaMalloc(&dr, sizeof(double) * n) != cudaSuccess) {
        cudaFree(dc);
        return qp_fail(ctx, QP_ERR_OOM, "bit-flip form: cudaMalloc of a diagonal failed");
      }
      int h_flags[2] = {0, 0};
      cudaMemsetAsync(d_flags, 0, 2 * sizeof(int), ctx->stream);
      k_bf_extract_diag<<<blocks, 256, 0, ctx->stream>>>(op->d_ptr, op->d_col, op->d_val, n, dc, dr, d_flags);
      ctx->launches++;
      cudaMemcpyAsync(h_flags, d_flags, sizeof(h_flags), cudaMemcpyDeviceToHost, ctx->stream);
      if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        cudaFree(dc);
        cudaFree(dr);
        return qp_fail(ctx, QP_ERR_CUDA, "bit-flip form: diagonal extraction failed");
      }
      if (h_flags[1] == 0) {
        cudaFree(dc);
        B->owned.push_back(dr);
        v.diag_r[v.n_diag] = dr;
      } else {
        cudaFree(dr);
        B->owned.push_back(dc);
        v.diag_c[v.n_diag] = dc;
        C.diags_real = false;
      }
      v.diag_op[v.n_diag++] = l;
    }
  }
  *ok = true;
  return QP_OK;
}

// D[i + n j] += c * (dl ? dl[i] : 1) * (dr ? dr[j] : 1)
__global__ void k_bf_outer_add(double2* __restrict__ D, const double2* __restrict__ dl, const double2* __restrict__ dr, double2 c,
                               int64_t n) {
  const int64_t total = n * n;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < total; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = r / n, i = r - j * n;
    double2 t = c;
    if (dl != nullptr) t = cmul2(t, dl[i]);
    if (dr != nullptr) t = cmul2(t, dr[j]);
    double2 d = D[r];
    d.x += t.x;
    d.y += t.y;
    D[r] = d;
  }
}

// Matrix-free left/right operators (qp_op_create_leftright): a term c P rho Q acts on the column-stacked index
// i + n j as (Q^T (x) P).  If every factor is itself a diagonal plus (conditional) uniform flips on n rows, the
// product is one on n^2 rows as long as no flip is multiplied by a NON-constant diagonal: flip x flip combines
// masks and conditions, flip x (constant diagonal, e.g. the identity) scales the value, diagonal x diagonal is an
// outer product added to the operator's diagonal vector.  This is the Liouvillian of config 4 WITHOUT ever building
// the n^2 x n^2 matrices: H rho and rho H have the identity on the other side, sigma^- rho sigma^+ is flip x flip,
// the anticommutator terms are diagonal.
static int32_t bf_collect_lr(qp_gen_t g, qp_bitflip_s* B, MaskSet* d_set, int* d_flags, BfCollect& C, bool* ok) {
  *ok = false;
  qp_ctx_t ctx = g->ctx;
  const int64_t nh = g->lr_n, N = g->n;
  if (nh <= 0 || (nh & (nh - 1)) != 0 || nh * nh != N) return QP_OK;
  int nb = 0;
  while ((int64_t(1) << nb) < nh) ++nb;
  BitflipView& v = B->view;
  struct Factor {
    bool ok = false, has_diag = false, diag_const = false;
    double2 diag_c = {0.0, 0.0};
    double2* d_diag = nullptr;  // device copy [nh] (nullptr: all ones)
    std::vector<BfTerm> flips;
  };
  std::vector<std::pair<const void*, Factor>> cache;
  std::vector<void*> temp;  // device arrays freed at the end
  auto cleanup = [&](int32_t rc) {
    for (void* p : temp) cudaFree(p);
    return rc;
  };
  std::vector<MaskSet> h_set(1);
  int32_t err = QP_OK;
  auto analyse = [&](const uint32_t* ptr, const uint32_t* col, const double2* val) -> const Factor* {
    for (auto& kv : cache)
      if (kv.first == (const void*)ptr) return &kv.second;
    Factor F;
    if (ptr == nullptr) {  // identity
      F.ok = F.has_diag = F.diag_const = true;
      F.diag_c = make_double2(1.0, 0.0);
    } else {
      err = bf_scan_matrix(ctx, ptr, col, val, nh, d_set, h_set[0]);
      if (err != QP_OK) return nullptr;
      F.ok = bf_terms_from_set(h_set[0], nh, F.flips, F.has_diag);
      if (F.ok && F.has_diag) {
        double2* dc = nullptr;
        double* dr = nullptr;
        if (cudaMalloc(&dc, sizeof(double2) * nh) != cudaSuccess || cudaMalloc(&dr, sizeof(double) * nh) != cudaSuccess) {
          cudaFree(dc);
          err = qp_fail(ctx, QP_ERR_OOM, "bit-flip form: cudaMalloc failed");
          return nullptr;
        }
        temp.push_back(dc);
        temp.push_back(dr);
        cudaMemsetAsync(d_flags, 0, 2 * sizeof(int), ctx->stream);
        k_bf_extract_diag<<<(unsigned)((nh + 255) / 256), 256, 0, ctx->stream>>>(ptr, col, val, nh, dc, dr, d_flags);
        ctx->launches++;
        std::vector<double2> h((size_t)nh);
        cudaMemcpyAsync(h.data(), dc, sizeof(double2) * nh, cudaMemcpyDeviceToHost, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
          err = qp_fail(ctx, QP_ERR_CUDA, "bit-flip form: diagonal extraction failed");
          return nullptr;
        }
        F.d_diag = dc;
        F.diag_const = true;
        for (int64_t i = 1; i < nh; ++i) F.diag_const = F.diag_const && h[i].x == h[0].x && h[i].y == h[0].y;
        F.diag_c = h[0];
      }
    }
    cache.emplace_back((const void*)ptr, F);
    return &cache.back().second;
  };
  const unsigned oblocks = (unsigned)std::min<int64_t>((N + 255) / 256, (int64_t)ctx->sm_count * 16);
  for (int l = 0; l < g->n_ops; ++l) {
    qp_op_t op = g->ops[l];
    if (!op->leftright) return cleanup(QP_OK);
    struct Acc { uint32_t mask, cmask, cval; double2 val; };
    std::vector<Acc> acc;
    auto add = [&](uint32_t m, uint32_t cm, uint32_t cv, double2 val) {
      for (Acc& a : acc)
        if (a.mask == m && a.cmask == cm && a.cval == cv) {
          a.val.x += val.x;
          a.val.y += val.y;
          return;
        }
      acc.push_back(Acc{m, cm, cv, val});
    };
    double2* D = nullptr;
    for (const LRTermHost& T : op->lr_terms) {
      const Factor* L = T.left ? analyse(T.left->d_ptr, T.left->d_col, T.left->d_val) : analyse(nullptr, nullptr, nullptr);
      if (L == nullptr) return cleanup(err);
      const Factor Lf = *L;  // the cache may reallocate
      const Factor* R = analyse(T.d_rptr, T.d_rcol, T.d_rval);
      if (R == nullptr) return cleanup(err);
      const Factor Rf = *R;
      if (!Lf.ok || !Rf.ok) return cleanup(QP_OK);
      if (T.left && (T.left->dense || T.left->leftright)) return cleanup(QP_OK);
      const double2 c = T.c;
      auto mul = [](double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); };
      if (Lf.has_diag && Rf.has_diag) {
        if (D == nullptr) {
          if (v.n_diag >= BF_MAX_DIAG) return cleanup(QP_OK);
          if (cudaMalloc(&D, sizeof(double2) * N) != cudaSuccess) return cleanup(qp_fail(ctx, QP_ERR_OOM, "bit-flip form: cudaMalloc of a diagonal failed"));
          B->owned.push_back(D);
          cudaMemsetAsync(D, 0, sizeof(double2) * N, ctx->stream);
        }
        k_bf_outer_add<<<oblocks, 256, 0, ctx->stream>>>(D, Lf.d_diag, Rf.d_diag, c, nh);
        ctx->launches++;
      }
      for (const BfTerm& a : Lf.flips) {
        if (Rf.has_diag) {
          if (!Rf.diag_const) return cleanup(QP_OK);  // a flip times a row-dependent diagonal is not uniform
          add(a.mask, a.cmask, a.cval, mul(c, mul(a.val, Rf.diag_c)));
        }
        for (const BfTerm& b : Rf.flips)
          add(a.mask | (b.mask << nb), a.cmask | (b.cmask << nb), a.cval | (b.cval << nb), mul(c, mul(a.val, b.val)));
      }
      if (Lf.has_diag)
        for (const BfTerm& b : Rf.flips) {
          if (!Lf.diag_const) return cleanup(QP_OK);
          add(b.mask << nb, b.cmask << nb, b.cval << nb, mul(c, mul(Lf.diag_c, b.val)));
        }
    }
    for (const Acc& a : acc) {
      if (a.val.x == 0.0 && a.val.y == 0.0) continue;
      if (a.val.y != 0.0) C.values_real = false;
      C.terms.push_back(BfTerm{a.mask, a.cmask, a.cval, a.val, l});
    }
    if ((int)C.terms.size() > BF_MAX_TERMS + BF_MAX_LOW - 4) return cleanup(QP_OK);
    if (D != nullptr) {
      v.diag_c[v.n_diag] = D;
      v.diag_op[v.n_diag++] = l;
      C.diags_real = false;
    }
  }
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return cleanup(qp_fail(ctx, QP_ERR_CUDA, "bit-flip form: diagonal assembly failed"));
  *ok = true;
  return cleanup(QP_OK);
}

// Tries to put the generator into bit-flip form.  *ok = false (nothing kept) if some operator is not a
// diagonal plus (conditional) uniform bit flips.
int32_t qp_bitflip_build(qp_gen_t g, bool* ok) {
  *ok = false;
  qp_ctx_t ctx = g->ctx;
  const int64_t n = g->n;
  if (n % 32 != 0 || n < 64 || n >= (int64_t(1) << 32)) return QP_OK;
  qp_bitflip_s* B = new qp_bitflip_s();
  BitflipView& v = B->view;
  memset(&v, 0, sizeof(v));
  v.n = n;
  int* d_flags = nullptr;
  MaskSet* d_set = nullptr;
  auto fail = [&](int32_t rc) {
    cudaFree(d_flags);
    cudaFree(d_set);
    qp_bitflip_free(B);
    return rc;
  };
  if (cudaMalloc(&d_flags, 2 * sizeof(int)) != cudaSuccess || cudaMalloc(&d_set, sizeof(MaskSet)) != cudaSuccess)
    return fail(qp_fail(ctx, QP_ERR_OOM, "bit-flip form: cudaMalloc failed"));
  const unsigned blocks = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 16);
  BfCollect C;
  bool collected = false;
  {
    bool any_lr = false;
    for (int l = 0; l < g->n_ops; ++l) any_lr = any_lr || g->ops[l]->leftright;
    const int32_t rc = any_lr ? bf_collect_lr(g, B, d_set, d_flags, C, &collected) : bf_collect_sparse(g, B, d_set, d_flags, C, &collected);
    if (rc != QP_OK || !collected) return fail(rc);
  }
  std::vector<BfTerm>& terms = C.terms;
  for (const BfTerm& t : terms)
    if (t.cmask != 0u) B->cond = true;
  const bool values_real = C.values_real, diags_real = C.diags_real;
  typedef BfTerm Term;
  cudaFree(d_set);
  d_set = nullptr;
  if (terms.empty()) return fail(QP_OK);  // purely diagonal generators gain nothing here
  // deterministic order (the hash set's is not): by operator, then by mask
  std::sort(terms.begin(), terms.end(), [](const Term& a, const Term& b) {
    if (a.op != b.op) return a.op < b.op;
    if (a.mask != b.mask) return a.mask < b.mask;
    if (a.cmask != b.cmask) return a.cmask < b.cmask;
    return a.cval < b.cval;
  });
  B->h_val.assign(BF_MAX_TERMS + BF_MAX_LOW, make_double2(0.0, 0.0));
  B->h_op.assign(BF_MAX_TERMS + BF_MAX_LOW, 0);
  const int max_low = getenv("QPROP_BITFLIP_SHUFFLES") ? std::min(BF_MAX_LOW, atoi(getenv("QPROP_BITFLIP_SHUFFLES"))) : BF_MAX_LOW;
  auto put_high = [&](const Term& t) {
    v.mask[v.n_high] = t.mask;
    v.cmask[v.n_high] = t.cmask;
    v.cval[v.n_high] = t.cval;
    B->h_val[v.n_high] = t.val;
    B->h_op[v.n_high++] = t.op;
  };
  std::vector<Term> low;
  for (const Term& t : terms) {
    if (t.mask < 32u && (int)low.size() < max_low) low.push_back(t);
    else if (v.n_high < BF_MAX_TERMS) put_high(t);
    else return fail(QP_OK);
  }
  // the load list is processed four terms at a time: fill it up with shuffle terms (an in-warp partner costs
  // the same as a load that hits L1) before padding it with zero-valued terms on the row itself
  while (v.n_high % 4 != 0 && !low.empty() && v.n_high < BF_MAX_TERMS) {
    put_high(low.back());
    low.pop_back();
  }
  while (v.n_high % 4 != 0 && v.n_high < BF_MAX_TERMS) put_high(Term{0u, 0u, 0u, make_double2(0.0, 0.0), 0});
  if (v.n_high % 4 != 0) return fail(QP_OK);
  for (const Term& t : low) {
    v.lmask[v.n_low] = t.mask;
    v.lcmask[v.n_low] = t.cmask;
    v.lcval[v.n_low] = t.cval;
    B->h_val[BF_MAX_TERMS + v.n_low] = t.val;
    B->h_op[BF_MAX_TERMS + v.n_low++] = t.op;
  }
  {
    double2* d_val = nullptr;
    int* d_op = nullptr;
    if (cudaMalloc(&d_val, sizeof(double2) * B->h_val.size()) != cudaSuccess || cudaMalloc(&d_op, sizeof(int) * B->h_op.size()) != cudaSuccess) {
      cudaFree(d_val);
      return fail(qp_fail(ctx, QP_ERR_OOM, "bit-flip form: cudaMalloc failed"));
    }
    B->owned.push_back(d_val);
    B->owned.push_back(d_op);
    cudaMemcpy(d_val, B->h_val.data(), sizeof(double2) * B->h_val.size(), cudaMemcpyHostToDevice);
    if (cudaMemcpy(d_op, B->h_op.data(), sizeof(int) * B->h_op.size(), cudaMemcpyHostToDevice) != cudaSuccess)
      return fail(qp_fail(ctx, QP_ERR_CUDA, "bit-flip form: upload of the terms failed"));
    v.tval = d_val;
    v.top = d_op;
  }
  // coded diagonals
  bool all_diags_real = v.n_diag > 0;
  for (int i = 0; i < v.n_diag; ++i) all_diags_real = all_diags_real && v.diag_r[i] != nullptr;
  if (all_diags_real && !getenv("QPROP_BITFLIP_NO_CODES")) {
    unsigned long long* d_keys = nullptr;
    if (cudaMalloc(&d_keys, sizeof(unsigned long long) * BF_HASH_CAP) != cudaSuccess)
      return fail(qp_fail(ctx, QP_ERR_OOM, "bit-flip form: cudaMalloc failed"));
    std::vector<double> h_tab;
    DiagTabs T;
    memset(&T, 0, sizeof(T));
    T.n_diag = v.n_diag;
    bool codes_ok = true;
    int64_t joint = 1;
    for (int i = 0; i < v.n_diag && codes_ok; ++i) {
      int h_flags[2] = {0, 0};
      cudaMemsetAsync(d_keys, 0xFF, sizeof(unsigned long long) * BF_HASH_CAP, ctx->stream);
      cudaMemsetAsync(d_flags, 0, 2 * sizeof(int), ctx->stream);
      k_bf_distinct<<<blocks, 256, 0, ctx->stream>>>(v.diag_r[i], n, d_keys, d_flags);
      ctx->launches++;
      std::vector<unsigned long long> h_keys(BF_HASH_CAP);
      cudaMemcpyAsync(h_keys.data(), d_keys, sizeof(unsigned long long) * BF_HASH_CAP, cudaMemcpyDeviceToHost, ctx->stream);
      cudaMemcpyAsync(h_flags, d_flags, sizeof(h_flags), cudaMemcpyDeviceToHost, ctx->stream);
      if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        cudaFree(d_keys);
        return fail(qp_fail(ctx, QP_ERR_CUDA, "bit-flip form: distinct-value scan failed"));
      }
      std::vector<double> vals;
      for (unsigned long long k : h_keys)
        if (k != ~0ull) {
          double x;
          memcpy(&x, &k, 8);
          vals.push_back(x);
        }
      bool has_nan = false;
      for (double x : vals) has_nan = has_nan || x != x;
      if (h_flags[0] != 0 || vals.empty() || (int)vals.size() > BF_DIAG_DISTINCT || has_nan) {
        codes_ok = false;
        break;
      }
      std::sort(vals.begin(), vals.end());
      T.d[i] = v.diag_r[i];
      T.count[i] = (int)vals.size();
      T.stride[i] = (int)joint;
      T.off[i] = (int)h_tab.size();
      joint *= (int64_t)vals.size();
      if (joint > BF_MAX_JOINT) codes_ok = false;
      h_tab.insert(h_tab.end(), vals.begin(), vals.end());
    }
    cudaFree(d_keys);
    if (codes_ok) {
      double* d_tab = nullptr;
      uint16_t* d_code = nullptr;
      if (cudaMalloc(&d_tab, sizeof(double) * h_tab.size()) != cudaSuccess || cudaMalloc(&d_code, sizeof(uint16_t) * n) != cudaSuccess) {
        cudaFree(d_tab);
        return fail(qp_fail(ctx, QP_ERR_OOM, "bit-flip form: cudaMalloc of the diagonal codes failed"));
      }
      B->owned.push_back(d_tab);
      B->owned.push_back(d_code);
      int h_flags[2] = {0, 0};
      cudaMemcpyAsync(d_tab, h_tab.data(), sizeof(double) * h_tab.size(), cudaMemcpyHostToDevice, ctx->stream);
      cudaMemsetAsync(d_flags, 0, 2 * sizeof(int), ctx->stream);
      k_bf_encode<<<blocks, 256, 0, ctx->stream>>>(T, d_tab, n, d_code, d_flags);
      ctx->launches++;
      cudaMemcpyAsync(h_flags, d_flags, sizeof(h_flags), cudaMemcpyDeviceToHost, ctx->stream);
      if (cudaStreamSynchronize(ctx->stream) != cudaSuccess || h_flags[0] != 0)
        return fail(qp_fail(ctx, QP_ERR_INTERNAL, "bit-flip form: encoding the diagonals failed"));
      v.dcode = d_code;
      v.dtab = d_tab;
      v.n_joint = (int)joint;
      for (int i = 0; i < v.n_diag; ++i) {
        v.dcount[i] = T.count[i];
        v.dstride[i] = T.stride[i];
        v.doff[i] = T.off[i];
        // the vectors are not needed any more
        for (void*& p : B->owned)
          if (p == (void*)v.diag_r[i]) {
            cudaFree(p);
            p = nullptr;
          }
        v.diag_r[i] = nullptr;
      }
    }
  }
  B->values_real = values_real;
  B->diags_real = diags_real;
  cudaFree(d_flags);
  g->bitflip = B;
  *ok = true;
  return QP_OK;
}

int64_t qp_bitflip_stored_bytes(const qp_bitflip_s* b) {
  int64_t bytes = 0;
  if (b->view.dcode != nullptr) return 2 * b->view.n;
  for (int i = 0; i < b->view.n_diag; ++i) bytes += (b->view.diag_r[i] ? 8 : 16) * b->view.n;
  return bytes;
}

// ---------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------

__device__ __forceinline__ double2 shfl_xor_c(double2 v, int m) {
  return make_double2(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m));
}

// One CTA per SM owns a contiguous slice range (near partners of neighbouring slices meet in L1); a warp takes a
// slice per round, a lane a row.  Per round and lane: LB gathers in flight (batches over the first BF_STATIC load
// terms, whose masks / products sit at compile-time positions of the constant bank), the shuffle terms, the
// diagonal look-up, the fused epilogue.  Nothing is carried from round to round, so the kernel fits 72-80 registers
// and 24-28 warps per SM hide the L1 / L2 latencies -- measured on config 2: 16 warps x 16 gathers with a
// cross-round prefetch (128 registers) 2140 prop_step!/s, 24 x 8: 2410, 28 x 8: 2480-2540, 32 x 4: 2400;
// with every far partner replaced by a near one (no L2 traffic at all) 2570: what is left is the L1 / shared-
// memory data pipe (60 wavefronts of gathers, 20 of shuffles, 20 of vector streams per slice).
// CK: where the products coefficient x value come from -- 0: shared memory, computed in the prologue from the
// device copy of the coefficients; 1 / 2: the constant bank, computed on the host (all real / complex).
// COND: some term is conditional (its product is dropped on the rows outside its sub-cube; the load is issued
// regardless -- predicating it on "the sub-cube excludes the whole warp" made config 4 7 % SLOWER, like every
// other attempt to put a predicate in front of these loads).
// NS: load terms at compile-time positions (16 or BF_STATIC; the unrolled batches beyond the list cost ~10 %).
template <int EPI, int CK, int COND, int THREADS, int LB, int NS>
__global__ void __launch_bounds__(THREADS, 1)
k_spmv_bitflip(const __grid_constant__ BitflipView v, const double2* __restrict__ coef, const double2* __restrict__ x,
               EpiArgs e, int rounds) {
  __shared__ double2 s_c[CK == 0 ? BF_MAX_TERMS + BF_MAX_LOW : 1];
  __shared__ double2 s_cd[BF_MAX_DIAG];
  extern __shared__ double2 s_dt[];  // coded diagonals: sum_i coefficient_i x value_i per joint code [n_joint]
  // programmatic dependent launch: the next term's launch and this set-up overlap the tail of the previous term
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (CK == 0) {
    for (int t = threadIdx.x; t < BF_MAX_TERMS + BF_MAX_LOW; t += THREADS) s_c[t] = cmul2(coef[v.top[t]], ld_stream(v.tval + t));  // static data: may sit above the wait
    if (threadIdx.x < v.n_diag) s_cd[threadIdx.x] = coef[v.diag_op[threadIdx.x]];
  }
  if (v.dcode != nullptr) {
    for (int c = threadIdx.x; c < v.n_joint; c += THREADS) {
      double re = 0.0, im = 0.0;
#pragma unroll
      for (int i = 0; i < BF_MAX_DIAG; ++i)
        if (i < v.n_diag) {
          const double d = v.dtab[v.doff[i] + (c / v.dstride[i]) % v.dcount[i]];
          const double2 u = CK == 0 ? coef[v.diag_op[i]] : v.cd[i];
          re = fma(u.x, d, re);
          if (CK != 1) im = fma(u.y, d, im);
        }
      s_dt[c] = make_double2(re, im);
    }
  }
  if (CK == 0 || v.dcode != nullptr) __syncthreads();
  asm volatile("griddepcontrol.wait;" ::: "memory");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = THREADS >> 5;
  const int64_t n_slices = v.n >> 5;
  double dr = 0, di = 0, nn = 0;
  // every warp of the grid runs the same number of rounds (a uniform trip count keeps the shuffles free of
  // re-convergence code); slices past the end (last CTA only) are computed on slice 0 and not stored -- predicating
  // their loads instead costs 13 % (the batches of gathers are broken up by branches)
  // 32-bit slice / row arithmetic (n < 2^32 is a condition of the format): two registers less than with 64-bit indices
  uint32_t s = blockIdx.x * (uint32_t)(rounds * nwarps) + (uint32_t)warp;
  for (int it = 0; it < rounds; ++it, s += (uint32_t)nwarps) {
    const bool active = s < (uint32_t)n_slices;
    const uint32_t r32 = (active ? s : 0u) * 32u + (uint32_t)lane;
    const int64_t row = (int64_t)r32;
    const double2 xown = ld_x(x + row);
    unsigned short code = 0;
    if (v.dcode != nullptr) code = __ldg(v.dcode + row);
    double hr = 0.0, hi = 0.0, hr2 = 0.0, hi2 = 0.0;
    double2 yv = make_double2(0.0, 0.0), av = yv;
    double2 xv[LB];
    // one term: product (from the constant bank or shared memory), condition, multiply-add
    auto accumulate = [&](double cr, double ci, uint32_t cm, uint32_t cv, const double2& xq, int chain) {
      if (COND) {
        const bool on = (r32 & cm) == cv;
        cr = on ? cr : 0.0;
        if (CK != 1) ci = on ? ci : 0.0;
      }
      if (CK == 1) {
        if (chain & 1) { hr2 = fma(cr, xq.x, hr2); hi2 = fma(cr, xq.y, hi2); }
        else { hr = fma(cr, xq.x, hr); hi = fma(cr, xq.y, hi); }
      } else {
        hr = fma(cr, xq.x, hr);
        hi = fma(cr, xq.y, hi);
        hr2 = fma(-ci, xq.y, hr2);
        hi2 = fma(ci, xq.x, hi2);
      }
    };
    // shuffle terms: the partner row is a lane of this warp.  They only need the lane's own x, so they run in
    // the shadow of the first batch of gathers (SHUF_EARLY) instead of after the last one
    auto shuffle_terms = [&]() {
#pragma unroll
      for (int q = 0; q < BF_MAX_LOW; ++q)
        if (q < v.n_low) {
          const double2 xs = shfl_xor_c(xown, (int)v.lmask[q]);
          if (CK == 0) {
            const double2 c = s_c[BF_MAX_TERMS + q];
            accumulate(c.x, c.y, v.lcmask[q], v.lcval[q], xs, 0);
          } else {
            accumulate(v.lcre[q], CK == 2 ? v.lcim[q] : 0.0, v.lcmask[q], v.lcval[q], xs, 0);
          }
        }
    };
#pragma unroll
    for (int b0 = 0; b0 < NS; b0 += LB) {
#pragma unroll
      for (int g = 0; g < LB / 4; ++g)
        if (b0 + 4 * g < v.n_high) {
#pragma unroll
          for (int q = 4 * g; q < 4 * g + 4; ++q) xv[q] = ld_x(x + (r32 ^ v.mask[b0 + q]));
        }
      if (SHUF_EARLY && b0 == 0) shuffle_terms();
      if (b0 == LB && THREADS < 1024) {  // the epilogue operands travel with the second batch
        if (EPI == EPI_MUL) {
          if (e.betac.x != 0.0 || e.betac.y != 0.0) yv = ld_noalloc(e.y + row);
        } else if (EPI == EPI_CHEB_MID || EPI == EPI_CHEB_LAST) {
          yv = ld_noalloc(e.y + row);
          av = ld_noalloc(e.acc + row);
        }
      }
#pragma unroll
      for (int g = 0; g < LB / 4; ++g)
        if (b0 + 4 * g < v.n_high) {
#pragma unroll
          for (int q = 4 * g; q < 4 * g + 4; ++q) {
            if (CK == 0) {
              const double2 c = s_c[b0 + q];
              accumulate(c.x, c.y, v.cmask[b0 + q], v.cval[b0 + q], xv[q], q);
            } else {
              accumulate(v.cre[b0 + q], CK == 2 ? v.cim[b0 + q] : 0.0, v.cmask[b0 + q], v.cval[b0 + q], xv[q], q);
            }
          }
        }
    }
    for (int t = NS; t < v.n_high; t += 4) {  // more load terms than static positions
#pragma unroll
      for (int q = 0; q < 4; ++q) xv[q] = ld_x(x + (r32 ^ v.mask[t + q]));
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (CK == 0) {
          const double2 c = s_c[t + q];
          accumulate(c.x, c.y, v.cmask[t + q], v.cval[t + q], xv[q], q);
        } else {
          accumulate(v.cre[t + q], CK == 2 ? v.cim[t + q] : 0.0, v.cmask[t + q], v.cval[t + q], xv[q], q);
        }
      }
    }
    if (!SHUF_EARLY) shuffle_terms();
    if (THREADS >= 1024) {  // 64 registers: the epilogue operands are requested once the gathers are consumed
      if (EPI == EPI_MUL) {
        if (e.betac.x != 0.0 || e.betac.y != 0.0) yv = ld_noalloc(e.y + row);
      } else if (EPI == EPI_CHEB_MID || EPI == EPI_CHEB_LAST) {
        yv = ld_noalloc(e.y + row);
        av = ld_noalloc(e.acc + row);
      }
    }
    double dre = 0.0, dim = 0.0;
    if (v.dcode != nullptr) {
      const double2 t = s_dt[code];
      dre = t.x;
      if (CK != 1) dim = t.y;
    } else {
      for (int i = 0; i < v.n_diag; ++i) {
        const double2 u = CK == 0 ? s_cd[i] : v.cd[i];
        if (v.diag_r[i] != nullptr) {
          const double d = ld_stream_f64(v.diag_r[i] + row);
          dre = fma(u.x, d, dre);
          if (CK != 1) dim = fma(u.y, d, dim);
        } else if (v.diag_c[i] != nullptr) {
          const double2 d = ld_stream(v.diag_c[i] + row);
          dre += u.x * d.x - u.y * d.y;
          dim += u.x * d.y + u.y * d.x;
        }
      }
    }
    hr = (hr + hr2) + (dre * xown.x - dim * xown.y);
    hi = (hi + hi2) + (dre * xown.y + dim * xown.x);
    if (active) epi_apply<EPI>(e, row, make_double2(hr, hi), xown, yv, av, dr, di, nn);
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

template <int EPI, int CK, int COND, int THREADS, int LB, int NS>
static int32_t bitflip_launch(qp_gen_t gen, const double2* x, const EpiArgs& e) {
  qp_ctx_t ctx = gen->ctx;
  qp_bitflip_s* B = gen->bitflip;
  BitflipView& v = B->view;
  if (CK != 0) {
    for (int t = 0; t < v.n_high; ++t) {
      const double2 u = gen->h_coef[B->h_op[t]], a = B->h_val[t];
      v.cre[t] = u.x * a.x - u.y * a.y;
      v.cim[t] = u.x * a.y + u.y * a.x;
    }
    for (int t = 0; t < v.n_low; ++t) {
      const double2 u = gen->h_coef[B->h_op[BF_MAX_TERMS + t]], a = B->h_val[BF_MAX_TERMS + t];
      v.lcre[t] = u.x * a.x - u.y * a.y;
      v.lcim[t] = u.x * a.y + u.y * a.x;
    }
    for (int i = 0; i < v.n_diag; ++i) v.cd[i] = gen->h_coef[v.diag_op[i]];
  }
  auto kern = k_spmv_bitflip<EPI, CK, COND, THREADS, LB, NS>;
  const int wpc = THREADS / 32;
  const int64_t n_slices = v.n >> 5;
  int64_t ctas = ctx->sm_count;
  int64_t spc = (n_slices + ctas - 1) / ctas;
  spc = (spc + wpc - 1) / wpc * wpc;  // whole rounds of the CTA's warps
  ctas = (n_slices + spc - 1) / spc;
  static const int pdl = getenv("QPROP_PDL") ? atoi(getenv("QPROP_PDL")) : 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)ctas);
  cfg.blockDim = dim3(THREADS);
  cfg.stream = ctx->stream;
  cfg.dynamicSmemBytes = v.dcode != nullptr ? sizeof(double2) * (size_t)v.n_joint : 0;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  EpiArgs e2 = e;
  e2.part = nullptr;
  static const int atomic = getenv("QPROP_ATOMIC_SUMS") ? atoi(getenv("QPROP_ATOMIC_SUMS")) : 0;
  const bool det = epi_has_sums(EPI) && e.chk != nullptr && !atomic;
  const int64_t n_warps = ctas * wpc;
  if (det) {
    QP_CHECK(qp_ctx_reserve_part(ctx, (size_t)3 * (size_t)n_warps));
    QP_CUDA(ctx, cudaMemsetAsync(ctx->d_part, 0, sizeof(double) * 3 * (size_t)n_warps, ctx->stream));
    e2.part = ctx->d_part;
  }
  const double2* coef = gen->d_coef;
  const int rounds = (int)(spc / wpc);
  QP_CUDA(ctx, cudaLaunchKernelEx(&cfg, kern, v, coef, x, e2, rounds));
  QP_LAUNCHED(ctx);
  if (det) QP_CHECK(qp_part_reduce(ctx, e2.part, n_warps, 1, e2.chk));
  return QP_OK;
}

template <int EPI, int CK, int COND>
static int32_t bitflip_launch_shape(qp_gen_t gen, const double2* x, const EpiArgs& e) {
  // CTA shape: 28 warps x 8 gathers (72 registers) or 32 warps x 4 (64) -- the one whose whole rounds waste fewer
  // SM slots (ties: 28 warps); QPROP_BITFLIP_THREADS = 896 / 1024 forces one
  static const int threads_env = getenv("QPROP_BITFLIP_THREADS") ? atoi(getenv("QPROP_BITFLIP_THREADS")) : 0;
  int threads = threads_env;
  if (threads != 896 && threads != 1024) {
    const int64_t n_slices = gen->n >> 5, per_sm = (n_slices + gen->ctx->sm_count - 1) / gen->ctx->sm_count;
    int64_t best = -1;
    for (int t : {896, 1024}) {
      const int wpc = t / 32;
      const int64_t spc = (per_sm + wpc - 1) / wpc * wpc, ctas = (n_slices + spc - 1) / spc;
      const int64_t slots = spc * gen->ctx->sm_count * ((ctas + gen->ctx->sm_count - 1) / gen->ctx->sm_count);
      // long term lists are latency-bound with 4 gathers in flight: 32 warps x 4 only if it saves 3 % of the slots
      if (best < 0 || (gen->bitflip->view.n_high > 16 ? slots * 100 < best * 97 : slots < best)) {
        best = slots;
        threads = t;
      }
    }
  }
  const bool few = gen->bitflip->view.n_high <= 16;
  if (threads == 1024) return few ? bitflip_launch<EPI, CK, COND, 1024, 4, 16>(gen, x, e) : bitflip_launch<EPI, CK, COND, 1024, 4, BF_STATIC>(gen, x, e);
  return few ? bitflip_launch<EPI, CK, COND, 896, 8, 16>(gen, x, e) : bitflip_launch<EPI, CK, COND, 896, 8, BF_STATIC>(gen, x, e);
}

template <int EPI>
static int32_t bitflip_launch_epi(qp_gen_t gen, const double2* x, const EpiArgs& e) {
  const qp_bitflip_s* B = gen->bitflip;
  // products on the host when the coefficients were set from the host and are not per-trajectory; real products
  // and real diagonal contributions: half the multiply-adds.  Coefficients that only exist on the device: one
  // general variant (products from shared memory)
  static const int device_coefs = getenv("QPROP_BITFLIP_DEVICE_COEFS") ? atoi(getenv("QPROP_BITFLIP_DEVICE_COEFS")) : 0;
  const bool host = gen->h_coef_valid && (int)gen->h_coef.size() == gen->n_ops && !device_coefs;
  if (!host) return bitflip_launch<EPI, 0, 1, 896, 8, BF_STATIC>(gen, x, e);
  bool real = B->values_real && B->diags_real;
  for (int l = 0; real && l < gen->n_ops; ++l) real = gen->h_coef[l].y == 0.0;
  if (B->cond) return real ? bitflip_launch_shape<EPI, 1, 1>(gen, x, e) : bitflip_launch_shape<EPI, 2, 1>(gen, x, e);
  return real ? bitflip_launch_shape<EPI, 1, 0>(gen, x, e) : bitflip_launch_shape<EPI, 2, 0>(gen, x, e);
}

int32_t qp_launch_bitflip(qp_gen_t gen, int epi, const double2* x, const EpiArgs& e) {
  switch (epi) {
    case EPI_MUL: return bitflip_launch_epi<EPI_MUL>(gen, x, e);
    case EPI_CHEB_FIRST: return bitflip_launch_epi<EPI_CHEB_FIRST>(gen, x, e);
    case EPI_CHEB_MID: return bitflip_launch_epi<EPI_CHEB_MID>(gen, x, e);
    case EPI_CHEB_LAST: return bitflip_launch_epi<EPI_CHEB_LAST>(gen, x, e);
    case EPI_CHEB_ONLY: return bitflip_launch_epi<EPI_CHEB_ONLY>(gen, x, e);
    case EPI_DOT: return bitflip_launch_epi<EPI_DOT>(gen, x, e);
  }
  return qp_fail(gen->ctx, QP_ERR_INTERNAL, "bad epilogue %d", epi);
}
