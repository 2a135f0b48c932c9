// Bit-flip (XOR-stencil) form of a generator, single states (QP_FORMAT_BITFLIP).
//
// Spin-1/2 Hamiltonians are very often "diagonal operators + sums of bit flips": the transverse-field
// Ising chain of BASELINE config 2 is H0 = -J sum Z_i Z_{i+1} (diagonal), H1 = sum X_i (row r couples to
// the 20 columns r XOR 2^i, every entry the SAME value), H2 = sum Z_i (diagonal); QAOA mixers and any
// transverse-field term look the same.  For such an operator nothing about an off-diagonal entry depends
// on the row: column = row XOR mask_t, value = v_t.  So there is no matrix stream at all -- the K <= 32
// (mask, value, operator) triples travel as a kernel parameter (constant bank), the diagonals are explicit
// vectors (8 B per row when real) -- and a fused term is
//
//     (H x)[r] = sum_l u_l ( d_l[r] x[r] + sum_{t in l} v_t x[r XOR mask_t] )
//
// with no code words, no table look-ups and no per-entry decode: masks below 32 are warp shuffles of the
// lane's own x[r], every other term is one coalesced 512 B load per warp.  The structure is DETECTED at
// qp_gen_create from the uploaded sparse matrices (every row of an operator must carry exactly the same
// set of masks with the same values; k_xor_check), never assumed; anything else keeps the dictionary /
// SELL / CSR formats.  Same fused epilogues as every other kernel (spmv.cuh).
//
// Replaces: mul!(C, A::Operator, B, alpha, beta), src/generators.jl:634-645, inside the Chebyshev term
// of src/cheby.jl:186-209 (config 2 of BASELINE.json).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>

#include "spmv.cuh"

constexpr int BF_MAX_TERMS = 32;
constexpr int BF_MAX_DIAG = 4;

constexpr int BF_MAX_LOW = 8;
constexpr int BF_MAX_JOINT = 2048;     // entries of the folded diagonal table (shared memory)
constexpr int BF_DIAG_DISTINCT = 256;  // distinct values per coded diagonal
constexpr int BF_HASH_CAP = 1024;

struct BitflipView {
  int64_t n;
  // terms served by loads (any mask; high[0, n_high)) and by warp shuffles (mask < 32; low[0, n_low))
  int n_high, n_low;
  uint32_t mask[BF_MAX_TERMS];
  double2 val[BF_MAX_TERMS];
  int op[BF_MAX_TERMS];
  uint32_t lmask[BF_MAX_LOW];
  double2 lval[BF_MAX_LOW];
  int lop[BF_MAX_LOW];
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
  // REALC launches only: the products coefficient x value (and the diagonals' coefficients), computed on the
  // host from the host copy of the coefficients -- they reach the DFMAs straight from the constant bank
  double cre[BF_MAX_TERMS];
  double lcre[BF_MAX_LOW];
  double cdr[BF_MAX_DIAG];
};

__device__ __forceinline__ double ld_stream_f64(const double* p) {
  double r;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
  return r;
}

struct qp_bitflip_s {
  BitflipView view;
  std::vector<void*> owned;  // device arrays of the diagonals
  bool all_real = false;     // every term value and every diagonal is real
};

void qp_bitflip_free(qp_bitflip_s* b) {
  if (!b) return;
  for (void* p : b->owned) cudaFree(p);
  delete b;
}

// ---------------------------------------------------------------------------------------
// detection
// ---------------------------------------------------------------------------------------

// flags[0] = 1 unless every row has at most one entry and that entry sits on the diagonal
__global__ void k_bf_check_diag(const uint32_t* __restrict__ ptr, const uint32_t* __restrict__ col, int64_t n, int* flags) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t p0 = ptr[r], p1 = ptr[r + 1];
    if (p1 - p0 > 1u || (p1 - p0 == 1u && col[p0] != (uint32_t)r)) flags[0] = 1;
  }
}

// diag[r] = the diagonal entry of row r (0 if absent); flags[1] = 1 if any imaginary part is non-zero
__global__ void k_bf_extract_diag(const uint32_t* __restrict__ ptr, const double2* __restrict__ val, int64_t n,
                                  double2* __restrict__ dc, double* __restrict__ dr, int* flags) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t p0 = ptr[r], p1 = ptr[r + 1];
    const double2 v = p1 > p0 ? val[p0] : make_double2(0.0, 0.0);
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

struct XorProbe {
  int k;
  uint32_t mask[BF_MAX_TERMS];
  unsigned long long vre[BF_MAX_TERMS], vim[BF_MAX_TERMS];  // value bits
};

// flags[0] = 1 unless every row carries exactly the masks of the probe (row 0), each once, with the same values
__global__ void k_bf_check_xor(const uint32_t* __restrict__ ptr, const uint32_t* __restrict__ col,
                               const double2* __restrict__ val, int64_t n, XorProbe pr, int* flags) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t p0 = ptr[r], p1 = ptr[r + 1];
    if ((int)(p1 - p0) != pr.k) {
      flags[0] = 1;
      continue;
    }
    unsigned seen = 0u;
    for (uint32_t p = p0; p < p1; ++p) {
      const uint32_t m = col[p] ^ (uint32_t)r;
      const double2 v = val[p];
      int hit = -1;
      for (int t = 0; t < pr.k; ++t)
        if (pr.mask[t] == m) hit = t;
      if (hit < 0 || (unsigned long long)__double_as_longlong(v.x) != pr.vre[hit] ||
          (unsigned long long)__double_as_longlong(v.y) != pr.vim[hit] || ((seen >> hit) & 1u)) {
        flags[0] = 1;
        break;
      }
      seen |= 1u << hit;
    }
  }
}

// Tries to put the generator into bit-flip form.  *ok = false (nothing kept) if any operator is neither
// purely diagonal nor a uniform XOR stencil.
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
  auto fail = [&](int32_t rc) {
    cudaFree(d_flags);
    qp_bitflip_free(B);
    return rc;
  };
  if (cudaMalloc(&d_flags, 2 * sizeof(int)) != cudaSuccess) return fail(qp_fail(ctx, QP_ERR_OOM, "bit-flip form: cudaMalloc failed"));
  const unsigned blocks = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 16);
  bool all_real = true;
  struct Term { uint32_t mask; double2 val; int op; };
  std::vector<Term> terms;
  for (int l = 0; l < g->n_ops; ++l) {
    qp_op_t op = g->ops[l];
    if (op->dense || op->leftright || op->nnz == 0) {
      if (op->nnz == 0 && !op->dense && !op->leftright) continue;  // an all-zero operator contributes nothing
      return fail(QP_OK);
    }
    int h_flags[2] = {0, 0};
    cudaMemsetAsync(d_flags, 0, 2 * sizeof(int), ctx->stream);
    if (op->nnz <= n) {  // candidate diagonal
      k_bf_check_diag<<<blocks, 256, 0, ctx->stream>>>(op->d_ptr, op->d_col, n, d_flags);
      ctx->launches++;
      cudaMemcpyAsync(h_flags, d_flags, sizeof(h_flags), cudaMemcpyDeviceToHost, ctx->stream);
      if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return fail(qp_fail(ctx, QP_ERR_CUDA, "bit-flip form: diagonal check failed"));
      if (h_flags[0] == 0) {
        if (v.n_diag >= BF_MAX_DIAG) return fail(QP_OK);
        double2* dc = nullptr;
        double* dr = nullptr;
        if (cudaMalloc(&dc, sizeof(double2) * n) != cudaSuccess || cudaMalloc(&dr, sizeof(double) * n) != cudaSuccess) {
          cudaFree(dc);
          return fail(qp_fail(ctx, QP_ERR_OOM, "bit-flip form: cudaMalloc of a diagonal failed"));
        }
        cudaMemsetAsync(d_flags, 0, 2 * sizeof(int), ctx->stream);
        k_bf_extract_diag<<<blocks, 256, 0, ctx->stream>>>(op->d_ptr, op->d_val, n, dc, dr, d_flags);
        ctx->launches++;
        cudaMemcpyAsync(h_flags, d_flags, sizeof(h_flags), cudaMemcpyDeviceToHost, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
          cudaFree(dc);
          cudaFree(dr);
          return fail(qp_fail(ctx, QP_ERR_CUDA, "bit-flip form: diagonal extraction failed"));
        }
        const bool real = h_flags[1] == 0;
        if (real) {
          cudaFree(dc);
          B->owned.push_back(dr);
          v.diag_r[v.n_diag] = dr;
        } else {
          cudaFree(dr);
          B->owned.push_back(dc);
          v.diag_c[v.n_diag] = dc;
          all_real = false;
        }
        v.diag_op[v.n_diag++] = l;
        continue;
      }
    }
    // candidate XOR stencil: k = nnz / n entries per row, the masks and values of row 0
    if (op->nnz % n != 0) return fail(QP_OK);
    const int64_t k = op->nnz / n;
    if (k < 1 || k > BF_MAX_TERMS || (int)terms.size() + k > BF_MAX_TERMS) return fail(QP_OK);
    std::vector<uint32_t> h_col((size_t)k);
    std::vector<double2> h_val((size_t)k);
    uint32_t h_ptr[2] = {0, 0};
    cudaMemcpy(h_ptr, op->d_ptr, sizeof(h_ptr), cudaMemcpyDeviceToHost);
    if ((int64_t)(h_ptr[1] - h_ptr[0]) != k) return fail(QP_OK);
    cudaMemcpy(h_col.data(), op->d_col + h_ptr[0], sizeof(uint32_t) * k, cudaMemcpyDeviceToHost);
    if (cudaMemcpy(h_val.data(), op->d_val + h_ptr[0], sizeof(double2) * k, cudaMemcpyDeviceToHost) != cudaSuccess)
      return fail(qp_fail(ctx, QP_ERR_CUDA, "bit-flip form: probe download failed"));
    XorProbe pr;
    memset(&pr, 0, sizeof(pr));
    pr.k = (int)k;
    for (int t = 0; t < (int)k; ++t) {
      pr.mask[t] = h_col[t];  // row 0: column XOR 0
      if (pr.mask[t] == 0u) return fail(QP_OK);  // a diagonal entry inside a stencil operator
      memcpy(&pr.vre[t], &h_val[t].x, 8);
      memcpy(&pr.vim[t], &h_val[t].y, 8);
      for (int s = 0; s < t; ++s)
        if (pr.mask[s] == pr.mask[t]) return fail(QP_OK);
    }
    cudaMemsetAsync(d_flags, 0, 2 * sizeof(int), ctx->stream);
    k_bf_check_xor<<<blocks, 256, 0, ctx->stream>>>(op->d_ptr, op->d_col, op->d_val, n, pr, d_flags);
    ctx->launches++;
    cudaMemcpyAsync(h_flags, d_flags, sizeof(h_flags), cudaMemcpyDeviceToHost, ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return fail(qp_fail(ctx, QP_ERR_CUDA, "bit-flip form: stencil check failed"));
    if (h_flags[0] != 0) return fail(QP_OK);
    for (int t = 0; t < (int)k; ++t) {
      terms.push_back(Term{pr.mask[t], h_val[t], l});
      if (h_val[t].y != 0.0) all_real = false;
    }
  }
  if (terms.empty()) return fail(QP_OK);  // purely diagonal generators gain nothing here
  const int max_low = getenv("QPROP_BITFLIP_SHUFFLES") ? std::min(BF_MAX_LOW, atoi(getenv("QPROP_BITFLIP_SHUFFLES"))) : BF_MAX_LOW;
  for (const Term& t : terms) {
    if (t.mask < 32u && v.n_low < max_low) {
      v.lmask[v.n_low] = t.mask;
      v.lval[v.n_low] = t.val;
      v.lop[v.n_low++] = t.op;
    } else {
      v.mask[v.n_high] = t.mask;
      v.val[v.n_high] = t.val;
      v.op[v.n_high++] = t.op;
    }
  }
  // the load list is processed four terms at a time: fill it up with shuffle terms (an in-warp partner costs
  // the same as a load that hits L1) before padding it
  while (v.n_high % 4 != 0 && v.n_low > 0 && v.n_high < BF_MAX_TERMS) {
    --v.n_low;
    v.mask[v.n_high] = v.lmask[v.n_low];
    v.val[v.n_high] = v.lval[v.n_low];
    v.op[v.n_high++] = v.lop[v.n_low];
  }
  while (v.n_high % 4 != 0 && v.n_high < BF_MAX_TERMS) {  // zero-valued padding terms on the row itself
    v.mask[v.n_high] = 0u;
    v.val[v.n_high] = make_double2(0.0, 0.0);
    v.op[v.n_high++] = 0;
  }
  if (v.n_high % 4 != 0) return fail(QP_OK);
  // coded diagonals
  bool diags_real = v.n_diag > 0;
  for (int i = 0; i < v.n_diag; ++i) diags_real = diags_real && v.diag_r[i] != nullptr;
  if (diags_real && !getenv("QPROP_BITFLIP_NO_CODES")) {
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
  B->all_real = all_real;
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
// slice per round, a lane a row.  Per round and lane: LB gathers in flight (two batches for the first sixteen load
// terms, whose masks / products sit at compile-time positions of the constant bank), the shuffle terms, the
// diagonal look-up, the fused epilogue.  Nothing is carried from round to round, so the kernel fits 72-80 registers
// and 24-28 warps per SM hide the L1 / L2 latencies -- measured on config 2: 16 warps x 16 gathers with a
// cross-round prefetch (128 registers) 2140 prop_step!/s, 24 x 8: 2410, 28 x 8: 2480-2500, 32 x 4: 2400;
// with every far partner replaced by a near one (no L2 traffic at all) 2570: what is left is the L1 / shared-
// memory data pipe (60 wavefronts of gathers, 20 of shuffles, 20 of vector streams per slice).
// REALC: every (coefficient x value) product and every (coefficient x diagonal) of this launch is real.
template <int EPI, int REALC, int THREADS, int LB>
__global__ void __launch_bounds__(THREADS, 1)
k_spmv_bitflip(const __grid_constant__ BitflipView v, const double2* __restrict__ coef, const double2* __restrict__ x,
                    EpiArgs e, int rounds) {
  __shared__ double2 s_c[BF_MAX_TERMS];
  __shared__ double2 s_cl[BF_MAX_LOW];
  __shared__ double2 s_cd[BF_MAX_DIAG];
  extern __shared__ double2 s_dt[];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (!REALC) {
    if (threadIdx.x < v.n_high) s_c[threadIdx.x] = cmul2(coef[v.op[threadIdx.x]], v.val[threadIdx.x]);
    if (threadIdx.x < v.n_low) s_cl[threadIdx.x] = cmul2(coef[v.lop[threadIdx.x]], v.lval[threadIdx.x]);
    if (threadIdx.x < v.n_diag) s_cd[threadIdx.x] = coef[v.diag_op[threadIdx.x]];
  }
  if (v.dcode != nullptr) {
    for (int c = threadIdx.x; c < v.n_joint; c += THREADS) {
      double re = 0.0, im = 0.0;
#pragma unroll
      for (int i = 0; i < BF_MAX_DIAG; ++i)
        if (i < v.n_diag) {
          const double d = v.dtab[v.doff[i] + (c / v.dstride[i]) % v.dcount[i]];
          if (REALC) {
            re = fma(v.cdr[i], d, re);
          } else {
            const double2 u = coef[v.diag_op[i]];
            re = fma(u.x, d, re);
            im = fma(u.y, d, im);
          }
        }
      s_dt[c] = make_double2(re, im);
    }
  }
  if (!REALC || v.dcode != nullptr) __syncthreads();
  asm volatile("griddepcontrol.wait;" ::: "memory");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = THREADS >> 5;
  const int64_t n_slices = v.n >> 5;
  double dr = 0, di = 0, nn = 0;
  // every warp of the grid runs the same number of rounds (a uniform trip count keeps the shuffles free of
  // re-convergence code); slices past the end (last CTA only) are computed on slice 0 and not stored -- predicating
  // their loads instead costs 13 % (the batches of gathers are broken up by branches)
  int64_t s = (int64_t)blockIdx.x * rounds * nwarps + warp;
  for (int it = 0; it < rounds; ++it, s += nwarps) {
    const bool active = s < n_slices;
    const int64_t row = (active ? s : 0) * 32 + lane;
    const uint32_t r32 = (uint32_t)row;
    const double2 xown = ld_x(x + row);
    unsigned short code = 0;
    if (v.dcode != nullptr) code = __ldg(v.dcode + row);
    double hr = 0.0, hi = 0.0, hr2 = 0.0, hi2 = 0.0;
    double2 yv = make_double2(0.0, 0.0), av = yv;
    double2 xv[LB];
#pragma unroll
    for (int b0 = 0; b0 < 16; b0 += LB) {
#pragma unroll
      for (int g = 0; g < LB / 4; ++g)
        if (b0 + 4 * g < v.n_high) {
#pragma unroll
          for (int q = 4 * g; q < 4 * g + 4; ++q) xv[q] = ld_x(x + (r32 ^ v.mask[b0 + q]));
        }
      if (b0 == 8 && THREADS < 1024) {  // the epilogue operands travel with the second batch
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
            if (REALC) {
              const double c = v.cre[b0 + q];
              if (q & 1) { hr2 = fma(c, xv[q].x, hr2); hi2 = fma(c, xv[q].y, hi2); }
              else { hr = fma(c, xv[q].x, hr); hi = fma(c, xv[q].y, hi); }
            } else {
              const double2 c = s_c[b0 + q];
              hr = fma(c.x, xv[q].x, hr);
              hi = fma(c.x, xv[q].y, hi);
              hr2 = fma(-c.y, xv[q].y, hr2);
              hi2 = fma(c.y, xv[q].x, hi2);
            }
          }
        }
    }
    for (int t = 16; t < v.n_high; t += 4) {  // more than sixteen load terms
#pragma unroll
      for (int q = 0; q < 4; ++q) xv[q] = ld_x(x + (r32 ^ v.mask[t + q]));
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double2 c = REALC ? make_double2(v.cre[t + q], 0.0) : s_c[t + q];
        hr = fma(c.x, xv[q].x, hr);
        hi = fma(c.x, xv[q].y, hi);
        if (!REALC) {
          hr2 = fma(-c.y, xv[q].y, hr2);
          hi2 = fma(c.y, xv[q].x, hi2);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < BF_MAX_LOW; ++q)
      if (q < v.n_low) {
        const double2 xs = shfl_xor_c(xown, (int)v.lmask[q]);
        if (REALC) {
          const double c = v.lcre[q];
          hr = fma(c, xs.x, hr);
          hi = fma(c, xs.y, hi);
        } else {
          const double2 c = s_cl[q];
          hr = fma(c.x, xs.x, hr);
          hi = fma(c.x, xs.y, hi);
          hr2 = fma(-c.y, xs.y, hr2);
          hi2 = fma(c.y, xs.x, hi2);
        }
      }
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
      if (!REALC) dim = t.y;
    } else {
      for (int i = 0; i < v.n_diag; ++i) {
        const double2 u = REALC ? make_double2(v.cdr[i], 0.0) : s_cd[i];
        if (v.diag_r[i] != nullptr) {
          const double d = ld_stream_f64(v.diag_r[i] + row);
          dre = fma(u.x, d, dre);
          if (!REALC) dim = fma(u.y, d, dim);
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

template <int EPI, int REALC, int THREADS, int LB>
static int32_t bitflip_launch(qp_gen_t gen, const double2* x, const EpiArgs& e) {
  qp_ctx_t ctx = gen->ctx;
  BitflipView& v = gen->bitflip->view;
  if (REALC) {
    for (int t = 0; t < v.n_high; ++t) v.cre[t] = gen->h_coef[v.op[t]].x * v.val[t].x;
    for (int t = 0; t < v.n_low; ++t) v.lcre[t] = gen->h_coef[v.lop[t]].x * v.lval[t].x;
    for (int i = 0; i < v.n_diag; ++i) v.cdr[i] = gen->h_coef[v.diag_op[i]].x;
  }
  auto kern = k_spmv_bitflip<EPI, REALC, THREADS, LB>;
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

template <int EPI>
static int32_t bitflip_launch_epi(qp_gen_t gen, const double2* x, const EpiArgs& e) {
  // real products: known on the host when the coefficients were set from the host and are not per-trajectory
  bool realc = gen->bitflip->all_real && gen->h_coef_valid && (int)gen->h_coef.size() == gen->n_ops;
  for (int l = 0; realc && l < gen->n_ops; ++l) realc = gen->h_coef[l].y == 0.0;
  // CTA shape: 24 / 28 warps x 8 gathers or 32 warps x 4 -- the one whose whole rounds waste the fewest SM slots
  // (ties: 28 warps); QPROP_BITFLIP_THREADS = 768 / 896 / 1024 forces one
  static const int threads_env = getenv("QPROP_BITFLIP_THREADS") ? atoi(getenv("QPROP_BITFLIP_THREADS")) : 0;
  int threads = threads_env;
  if (threads != 768 && threads != 896 && threads != 1024) {
    const int64_t n_slices = gen->n >> 5, per_sm = (n_slices + gen->ctx->sm_count - 1) / gen->ctx->sm_count;
    int64_t best = -1;
    for (int t : {896, 768, 1024}) {
      const int wpc = t / 32;
      const int64_t spc = (per_sm + wpc - 1) / wpc * wpc, ctas = (n_slices + spc - 1) / spc;
      const int64_t slots = spc * gen->ctx->sm_count * ((ctas + gen->ctx->sm_count - 1) / gen->ctx->sm_count);
      if (best < 0 || slots < best) {
        best = slots;
        threads = t;
      }
    }
  }
  if (threads == 1024) return realc ? bitflip_launch<EPI, 1, 1024, 4>(gen, x, e) : bitflip_launch<EPI, 0, 1024, 4>(gen, x, e);
  if (threads == 768) return realc ? bitflip_launch<EPI, 1, 768, 8>(gen, x, e) : bitflip_launch<EPI, 0, 768, 8>(gen, x, e);
  return realc ? bitflip_launch<EPI, 1, 896, 8>(gen, x, e) : bitflip_launch<EPI, 0, 896, 8>(gen, x, e);
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
