// Two-pass tile format for the trajectory-batched SpMM (host-side builder; plain C++17, no CUDA).
//
// The batched application  Y[r, b] = sum_l u_l^(b) sum_c (H_l)[r, c] X[c, b]  of a generator whose
// matrices come from tensor products of few-level operators is bounded by how often a value of X is
// fetched again (every column is referenced by ~30 rows; through L1/L2 that was 6 x
// the algorithmic traffic on config 3).  Here the index is split  r = hi * S + lo  (S a power of two
// near sqrt(N)) and every off-diagonal entry (r, c) is put in one of three classes:
//
//   A  r and c lie in the same block of S consecutive rows        (c / S == r / S)
//   B  r and c have the same position inside their blocks         (c % S == r % S)
//   O  anything else (couplings that straddle the split)
//
// Pass A runs over "A tiles" (S consecutive rows x 32 trajectories) and pass B over "B tiles" (the
// N/S rows {hi * S + lo : hi} of one lo x 32 trajectories).  A tile of X is brought into shared
// memory ONCE and every class-A (pass A) or class-B (pass B) reference is served from there; class
// O entries are gathered from global memory in pass B; the diagonals are kept as explicit vectors.
// Entries of different operators that sit in the same column are merged: one table entry holds the
// column offset and one real number per operator (generators whose operators are purely real or
// purely imaginary -- the factor i of an imaginary operator moves into its coefficient), so a
// shared column costs one load.  A row is a list of 16-bit codes into that table (code 0 = padding).
//
// Replaces (for B > 1 states sharing one generator): mul!(C, A::Operator, B, alpha, beta) of the
// reference, src/generators.jl:634-645, called once per Chebyshev term from src/cheby.jl:175,189.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

namespace qptile {

constexpr int TILE_MAX_OPS = 3;
constexpr int TILE_TRAJ = 32;          // trajectories per tile = one warp
constexpr int TILE_MAX_ROWS = 256;     // rows per tile: 256 x 32 x 16 B = 128 KB of shared memory
constexpr int TILE_MAX_TABLE = 2048;   // 16-bit codes; 2048 x 32 B = 64 KB of shared memory next to the 128 KB tile
constexpr uint32_t TILE_KIND_O = 1u << 8;

struct TileEntry {      // 32 bytes, read with two 16-byte shared-memory loads: {v0, off, km} {v1, v2}
  double v0;               // real representation of the value of operator 0 (0: operator absent)
  int32_t off;             // class A / B: byte offset inside the shared-memory tile (slot delta * 512);
                           // class O: row delta (c - r)
  uint32_t km;             // bits 0..2: operators present; TILE_KIND_O: class O
  double v1, v2;           // operators 1 and 2
  double v(int l) const { return l == 0 ? v0 : l == 1 ? v1 : v2; }
};
static_assert(sizeof(TileEntry) == 32, "TileEntry must be 32 bytes");

struct TileFormat {
  int64_t n = 0;
  int S = 0, NH = 0, n_ops = 0;
  unsigned imag_ops = 0;            // bit l: operator l is purely imaginary (v = Im)
  int WA = 0, WB = 0, WO = 0;       // codes per row: class A (pass A), class B and class O (pass B); multiples of 8
  std::vector<TileEntry> table;     // [n_table], entry 0 = padding
  std::vector<uint16_t> codesA;     // [n][WA]
  std::vector<uint16_t> codesB;     // [n][WB]
  std::vector<uint16_t> codesO;     // [n][WO]
  std::vector<double> diag;         // [n][TILE_MAX_OPS] real representation of the main diagonals
  int64_t n_A = 0, n_B = 0, n_O = 0, n_diag = 0;  // merged entries per class
  std::string why;                  // reason when build() returns false
};

inline int choose_split(int64_t n) {
  // S = power of two with n % S == 0 and both tile heights (S and N / S) <= 256; the most
  // balanced such split (smallest max(S, N / S)); ties go to the larger S
  int best = 0;
  int64_t best_cost = -1;
  for (int S = TILE_MAX_ROWS; S >= 2; S >>= 1) {
    if (n % S != 0) continue;
    const int64_t nh = n / S;
    if (nh > TILE_MAX_ROWS) break;  // a smaller S only makes N / S larger
    const int64_t cost = std::max<int64_t>(S, nh);
    if (best == 0 || cost < best_cost) {
      best = S;
      best_cost = cost;
    }
  }
  return best;
}

// mptr / colop / val: merged multi-operator CSR (operator index in the top 4 bits of the column
// word, 28 column bits), values as (re, im) pairs.
inline bool build(TileFormat& f, int64_t n, int n_ops, const uint32_t* mptr, const uint32_t* colop,
                  const double* val_reim, int S_forced = 0) {
  constexpr int COL_BITS = 28;
  constexpr uint32_t COL_MASK = (1u << COL_BITS) - 1u;
  f = TileFormat();
  f.n = n;
  f.n_ops = n_ops;
  if (n_ops < 1 || n_ops > TILE_MAX_OPS) { f.why = "more than 3 operators"; return false; }
  int S = S_forced > 0 ? S_forced : choose_split(n);
  if (S_forced <= 0 && S > 0) {
    // among all admissible splits take the one with the fewest class-O entries (a split between two
    // tensor factors: 4-level sites want S = 4^k); ties go to the most balanced one (choose_split)
    // (reasonably balanced splits only: both tile heights <= 4 sqrt(N), so that neither pass
    // degenerates into tiles of a few rows)
    int64_t best_o = -1;
    const int balanced = S;
    for (int cand = TILE_MAX_ROWS; cand >= 2; cand >>= 1) {
      if (n % cand != 0) continue;
      if (n / cand > TILE_MAX_ROWS) break;
      if (cand != balanced && (double)std::max<int64_t>(cand, n / cand) > 4.0 * std::sqrt((double)n)) continue;
      int64_t n_o = 0;
      for (int64_t r = 0; r < n; ++r)
        for (uint32_t k = mptr[r]; k < mptr[r + 1]; ++k) {
          const int64_t c = (int64_t)(colop[k] & COL_MASK);
          if (c != r && c / cand != r / cand && c % cand != r % cand) ++n_o;
        }
      if (best_o < 0 || n_o < best_o || (n_o == best_o && cand == balanced)) {
        best_o = n_o;
        S = cand;
      }
    }
  }
  if (S <= 0 || n % S != 0 || n / S > TILE_MAX_ROWS || S > TILE_MAX_ROWS) { f.why = "no two-level split of N with tiles of <= 256 rows"; return false; }
  f.S = S;
  f.NH = (int)(n / S);
  // every operator purely real or purely imaginary?
  unsigned has_re = 0, has_im = 0;
  const uint32_t nnz = mptr[n];
  for (uint32_t k = 0; k < nnz; ++k) {
    const int op = (int)(colop[k] >> COL_BITS);
    if (op >= n_ops) { f.why = "operator index out of range"; return false; }
    if (val_reim[2 * (size_t)k] != 0.0) has_re |= 1u << op;
    if (val_reim[2 * (size_t)k + 1] != 0.0) has_im |= 1u << op;
  }
  if (has_re & has_im) { f.why = "an operator has both real and imaginary values"; return false; }
  f.imag_ops = has_im;

  f.table.assign(1, TileEntry{0.0, 0, 0u, 0.0, 0.0});
  f.diag.assign((size_t)n * TILE_MAX_OPS, 0.0);
  using Key = std::tuple<uint32_t, int32_t, uint64_t, uint64_t, uint64_t>;
  std::map<Key, uint16_t> dict;
  std::vector<std::vector<uint16_t>> rowsA((size_t)n), rowsB((size_t)n), rowsO((size_t)n);
  struct Ent { int64_t col; int op; double v; };
  std::vector<Ent> row;
  auto bits = [](double d) { uint64_t u; std::memcpy(&u, &d, 8); return u; };
  for (int64_t r = 0; r < n; ++r) {
    row.clear();
    for (uint32_t k = mptr[r]; k < mptr[r + 1]; ++k) {
      const int op = (int)(colop[k] >> COL_BITS);
      const double v = ((has_im >> op) & 1u) ? val_reim[2 * (size_t)k + 1] : val_reim[2 * (size_t)k];
      row.push_back(Ent{(int64_t)(colop[k] & COL_MASK), op, v});
    }
    std::stable_sort(row.begin(), row.end(), [](const Ent& a, const Ent& b) { return a.col < b.col; });
    size_t i = 0;
    while (i < row.size()) {
      const int64_t c = row[i].col;
      double v[TILE_MAX_OPS] = {0.0, 0.0, 0.0};
      uint32_t mask = 0;
      for (; i < row.size() && row[i].col == c; ++i) {
        v[row[i].op] += row[i].v;  // duplicates within one operator add up (SparseArrays semantics)
        mask |= 1u << row[i].op;
      }
      if (c == r) {
        for (int l = 0; l < TILE_MAX_OPS; ++l) f.diag[(size_t)r * TILE_MAX_OPS + l] += v[l];
        ++f.n_diag;
        continue;
      }
      uint32_t km = mask;
      int32_t off;
      bool passA = false, isO = false;
      if (c / S == r / S) {  // class A: same block of S rows
        off = (int32_t)(c - r) * (TILE_TRAJ * 16);
        passA = true;
        ++f.n_A;
      } else if (c % S == r % S) {  // class B: same position in another block
        off = (int32_t)((c - r) / S) * (TILE_TRAJ * 16);
        ++f.n_B;
      } else {
        off = (int32_t)(c - r);
        km |= TILE_KIND_O;
        isO = true;
        ++f.n_O;
      }
      // pass A and pass B entries never share a table entry (their offsets mean different things)
      const Key key{km | (passA ? 1u << 16 : 0u), off, bits(v[0]), bits(v[1]), bits(v[2])};
      auto it = dict.find(key);
      uint16_t code;
      if (it == dict.end()) {
        if ((int)f.table.size() >= TILE_MAX_TABLE) { f.why = "more than 2047 distinct (offset, values) entries"; return false; }
        code = (uint16_t)f.table.size();
        f.table.push_back(TileEntry{v[0], off, km, v[1], v[2]});
        dict.emplace(key, code);
      } else {
        code = it->second;
      }
      (passA ? rowsA : isO ? rowsO : rowsB)[(size_t)r].push_back(code);
    }
  }
  size_t wa = 0, wb = 0, wo = 0;
  for (int64_t r = 0; r < n; ++r) {
    wa = std::max(wa, rowsA[(size_t)r].size());
    wb = std::max(wb, rowsB[(size_t)r].size());
    wo = std::max(wo, rowsO[(size_t)r].size());
  }
  f.WA = (int)((wa + 7) / 8 * 8);
  f.WB = (int)((wb + 7) / 8 * 8);
  f.WO = (int)((wo + 7) / 8 * 8);
  f.codesA.assign((size_t)n * f.WA, 0);
  f.codesB.assign((size_t)n * f.WB, 0);
  f.codesO.assign((size_t)n * f.WO, 0);
  for (int64_t r = 0; r < n; ++r) {
    std::copy(rowsA[(size_t)r].begin(), rowsA[(size_t)r].end(), f.codesA.begin() + (size_t)r * f.WA);
    std::copy(rowsB[(size_t)r].begin(), rowsB[(size_t)r].end(), f.codesB.begin() + (size_t)r * f.WB);
    std::copy(rowsO[(size_t)r].begin(), rowsO[(size_t)r].end(), f.codesO.begin() + (size_t)r * f.WO);
  }
  return true;
}

// CPU emulation of the two kernels' traversal of the format (test infrastructure for the builder):
// y[r*B + b] = sum_l u[l*B + b] * (H_l x)[r, b], complex numbers as (re, im) pairs.
inline void apply_host(const TileFormat& f, int64_t B, const double* u_reim /*[n_ops][B]*/,
                       const double* x_reim /*[n][B]*/, double* y_reim /*[n][B]*/) {
  const int64_t n = f.n;
  const int S = f.S;
  std::vector<double> t((size_t)n * B * 2, 0.0);
  auto coef = [&](int l, int64_t b, double& ur, double& ui) {
    ur = u_reim[2 * ((size_t)l * B + b)];
    ui = u_reim[2 * ((size_t)l * B + b) + 1];
    if ((f.imag_ops >> l) & 1u) {  // times i
      const double a = ur;
      ur = -ui;
      ui = a;
    }
  };
  for (int pass = 0; pass < 2; ++pass) {
    for (int64_t r = 0; r < n; ++r) {
      // slot of row r in its tile
      const int64_t slot = pass == 0 ? r % S : r / S;
      for (int64_t b = 0; b < B; ++b) {
        double pr[TILE_MAX_OPS] = {0, 0, 0}, pi[TILE_MAX_OPS] = {0, 0, 0};
        for (int list = 0; list < (pass == 0 ? 1 : 2); ++list) {
          const int W = pass == 0 ? f.WA : list == 0 ? f.WB : f.WO;
          const std::vector<uint16_t>& codes = pass == 0 ? f.codesA : list == 0 ? f.codesB : f.codesO;
          for (int j = 0; j < W; ++j) {
            const uint16_t code = codes[(size_t)r * W + j];
            if (code == 0) continue;
            const TileEntry& e = f.table[code];
            int64_t c;
            if (e.km & TILE_KIND_O) {
              c = r + e.off;
            } else {
              const int64_t s2 = slot + e.off / (TILE_TRAJ * 16);
              c = pass == 0 ? (r / S) * S + s2 : s2 * S + r % S;
            }
            const double xr = x_reim[2 * ((size_t)c * B + b)], xi = x_reim[2 * ((size_t)c * B + b) + 1];
            for (int l = 0; l < f.n_ops; ++l) {  // absent operators have v = 0 (the kernels multiply unconditionally)
              pr[l] += e.v(l) * xr;
              pi[l] += e.v(l) * xi;
            }
          }
        }
        if (pass == 0) {
          const double xr = x_reim[2 * ((size_t)r * B + b)], xi = x_reim[2 * ((size_t)r * B + b) + 1];
          for (int l = 0; l < f.n_ops; ++l) {
            pr[l] += f.diag[(size_t)r * TILE_MAX_OPS + l] * xr;
            pi[l] += f.diag[(size_t)r * TILE_MAX_OPS + l] * xi;
          }
        }
        double hr = 0, hi = 0;
        for (int l = 0; l < f.n_ops; ++l) {
          double ur, ui;
          coef(l, b, ur, ui);
          hr += ur * pr[l] - ui * pi[l];
          hi += ur * pi[l] + ui * pr[l];
        }
        double* dst = pass == 0 ? &t[2 * ((size_t)r * B + b)] : &y_reim[2 * ((size_t)r * B + b)];
        if (pass == 0) {
          dst[0] = hr;
          dst[1] = hi;
        } else {
          dst[0] = t[2 * ((size_t)r * B + b)] + hr;
          dst[1] = t[2 * ((size_t)r * B + b) + 1] + hi;
        }
      }
    }
  }
}

}  // namespace qptile
