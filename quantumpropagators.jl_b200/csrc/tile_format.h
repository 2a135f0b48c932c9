// Two-pass tile format for the trajectory-batched SpMM (host-side builder; plain C++17, no CUDA).
//
// The batched application  Y[r, b] = sum_l u_l^(b) sum_c (H_l)[r, c] X[c, b]  of a generator whose
// matrices come from tensor products of few-level operators is bounded by how often a value of X is
// fetched again (every column is referenced by ~30 rows; through L1/L2 that was 6 x
// the algorithmic traffic on config 3).  Here the index is split  r = hi * S + lo  (S a power of two,
// chosen between two tensor factors) and every off-diagonal entry (r, c) is put in one of three classes:
//
//   A  r and c lie in the same block of S consecutive rows        (c / S == r / S)
//   B  r and c have the same position inside their blocks         (c % S == r % S)
//   O  anything else (couplings that straddle the split)
//
// Pass A runs over "A tiles" (S consecutive rows x 32 trajectories) and pass B over "B tiles" (the
// N/S rows {hi * S + lo : hi} of one lo x 32 trajectories).  A tile of X is brought into shared
// memory ONCE and every class-A (pass A) or class-B (pass B) reference is served from there; class
// O entries are gathered from global memory in pass B; the diagonals are kept as explicit vectors.
//
// Entries of different operators that sit in the same column are merged (one load of X serves them
// all); operators must be purely real or purely imaginary (the factor i of an imaginary operator
// moves into its coefficient), so a value is ONE real number.  A merged entry has a KIND, and every
// row keeps one list of 16-bit codes per kind (code 0 = padding), so that the inner loops of the
// kernels contain no per-entry decisions:
//
//   S_l   one operator l only                        table entry 16 B {v, off}:   2 FMA into sum l
//   P     the operator pair (p1, p2) with |v2| = |v1| (quadrature controls: a + a^+ and i(a^+ - a))
//                                                    table entry 16 B {v, off, sign}: 4 FMA
//   G     anything else                              table entry 32 B {v0, off, v1, v2}: 2 n_ops FMA
//   O     class-O entries (always the 32-byte form), X from global memory
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
constexpr int TILE_MAX_TABLE = 2048;   // entries per table (16-bit codes; bounded by shared memory)
constexpr int TILE_ROW_BYTES = TILE_TRAJ * 16;

// kinds = lists of a row; pass A uses S0 S1 S2 P G, pass B the same plus O
enum Kind { K_S0 = 0, K_S1 = 1, K_S2 = 2, K_P = 3, K_G = 4, K_O = 5, N_KINDS = 6 };

struct Entry16 {   // kinds S_l and P
  double v;
  int32_t off;     // byte offset inside the shared-memory tile (slot delta * 512)
  uint32_t neg2;   // kind P: 0x80000000 if the second operator's value is -v (XORed into the sign), else 0
};
struct Entry32 {   // kinds G and O: {v0, off, pad} {v1, v2}
  double v0;
  int32_t off;     // G: byte offset inside the tile; O: row delta (c - r)
  uint32_t pad;
  double v1, v2;
  double v(int l) const { return l == 0 ? v0 : l == 1 ? v1 : v2; }
};
static_assert(sizeof(Entry16) == 16 && sizeof(Entry32) == 32, "table entry sizes");

struct TileFormat {
  int64_t n = 0;
  int S = 0, NH = 0, n_ops = 0;
  unsigned imag_ops = 0;            // bit l: operator l is purely imaginary (v = Im)
  int p1 = 0, p2 = 1;               // the operator pair of kind P
  int W[2][N_KINDS] = {{0}};        // codes per row, per pass and kind (multiples of 8; pass A has no O)
  int WT[2] = {0, 0};               // 16-byte code words per row and pass (all kinds back to back)
  int word0[2][N_KINDS] = {{0}};    // first word of a kind's list inside the row's words
  std::vector<Entry16> tab16;       // entry 0 = padding
  std::vector<Entry32> tab32;       // entry 0 = padding
  std::vector<uint16_t> codes[2];   // [n][WT * 8] per pass
  std::vector<double> diag;         // [n][TILE_MAX_OPS] real representation of the main diagonals
  int64_t count[2][N_KINDS] = {{0}};  // merged entries per pass and kind
  int64_t n_diag = 0;
  std::string why;                  // reason when build() returns false
  int64_t n_A() const { int64_t s = 0; for (int k = 0; k < K_O; ++k) s += count[0][k]; return s; }
  int64_t n_B() const { int64_t s = 0; for (int k = 0; k < K_O; ++k) s += count[1][k]; return s; }
  int64_t n_O() const { return count[1][K_O]; }
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
  // the pair of kind P: the last two operators (drift first, controls last: the quadrature pair)
  f.p1 = n_ops >= 2 ? n_ops - 2 : 0;
  f.p2 = n_ops >= 2 ? n_ops - 1 : 0;
  const uint32_t pair_mask = n_ops >= 2 ? ((1u << f.p1) | (1u << f.p2)) : 0u;

  f.tab16.assign(1, Entry16{0.0, 0, 0u});
  f.tab32.assign(1, Entry32{0.0, 0, 0u, 0.0, 0.0});
  f.diag.assign((size_t)n * TILE_MAX_OPS, 0.0);
  using Key16 = std::tuple<int32_t, uint64_t, uint32_t>;
  using Key32 = std::tuple<int, int32_t, uint64_t, uint64_t, uint64_t>;
  std::map<Key16, uint16_t> dict16;
  std::map<Key32, uint16_t> dict32;
  std::vector<std::vector<uint16_t>> lists[2][N_KINDS];
  for (int p = 0; p < 2; ++p)
    for (int k = 0; k < N_KINDS; ++k) lists[p][k].resize((size_t)n);
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
      for (; i < row.size() && row[i].col == c; ++i) v[row[i].op] += row[i].v;  // duplicates add up (SparseArrays semantics)
      uint32_t mask = 0;
      for (int l = 0; l < TILE_MAX_OPS; ++l)
        if (v[l] != 0.0) mask |= 1u << l;
      if (c == r) {
        for (int l = 0; l < TILE_MAX_OPS; ++l) f.diag[(size_t)r * TILE_MAX_OPS + l] += v[l];
        ++f.n_diag;
        continue;
      }
      if (mask == 0) continue;  // stored zeros
      int pass, kind;
      int32_t off;
      if (c / S == r / S) {  // class A: same block of S rows
        pass = 0;
        off = (int32_t)(c - r) * TILE_ROW_BYTES;
        kind = -1;
      } else if (c % S == r % S) {  // class B: same position in another block
        pass = 1;
        off = (int32_t)((c - r) / S) * TILE_ROW_BYTES;
        kind = -1;
      } else {
        pass = 1;
        off = (int32_t)(c - r);
        kind = K_O;
      }
      uint16_t code;
      if (kind != K_O) {
        const bool single = (mask & (mask - 1)) == 0;
        const bool pair = !single && mask == pair_mask && std::fabs(v[f.p1]) == std::fabs(v[f.p2]);
        if (single || pair) {
          const int l = single ? (mask == 1 ? 0 : mask == 2 ? 1 : 2) : f.p1;
          kind = single ? K_S0 + l : K_P;
          const uint32_t neg2 = (pair && (std::signbit(v[f.p1]) != std::signbit(v[f.p2]))) ? 0x80000000u : 0u;
          const Key16 key{off, bits(v[l]), neg2};
          auto it = dict16.find(key);
          if (it == dict16.end()) {
            if ((int)f.tab16.size() >= TILE_MAX_TABLE) { f.why = "more than 2047 distinct (offset, value) entries"; return false; }
            code = (uint16_t)f.tab16.size();
            f.tab16.push_back(Entry16{v[l], off, neg2});
            dict16.emplace(key, code);
          } else {
            code = it->second;
          }
        } else {
          kind = K_G;
        }
      }
      if (kind == K_G || kind == K_O) {
        const Key32 key{kind, off, bits(v[0]), bits(v[1]), bits(v[2])};
        auto it = dict32.find(key);
        if (it == dict32.end()) {
          if ((int)f.tab32.size() >= TILE_MAX_TABLE) { f.why = "more than 2047 distinct (offset, values) entries"; return false; }
          code = (uint16_t)f.tab32.size();
          f.tab32.push_back(Entry32{v[0], off, 0u, v[1], v[2]});
          dict32.emplace(key, code);
        } else {
          code = it->second;
        }
      }
      lists[pass][kind][(size_t)r].push_back(code);
      ++f.count[pass][kind];
    }
  }
  for (int p = 0; p < 2; ++p) {
    int words = 0;
    for (int k = 0; k < N_KINDS; ++k) {
      size_t w = 0;
      for (int64_t r = 0; r < n; ++r) w = std::max(w, lists[p][k][(size_t)r].size());
      f.W[p][k] = (int)((w + 7) / 8 * 8);
      f.word0[p][k] = words;
      words += f.W[p][k] / 8;
    }
    f.WT[p] = words;
    f.codes[p].assign((size_t)n * words * 8, 0);
    for (int k = 0; k < N_KINDS; ++k)
      for (int64_t r = 0; r < n; ++r)
        std::copy(lists[p][k][(size_t)r].begin(), lists[p][k][(size_t)r].end(),
                  f.codes[p].begin() + ((size_t)r * words + f.word0[p][k]) * 8);
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
      const int64_t slot = pass == 0 ? r % S : r / S;  // slot of row r in its tile
      auto tile_col = [&](int32_t off) {
        const int64_t s2 = slot + off / TILE_ROW_BYTES;
        return pass == 0 ? (r / S) * S + s2 : s2 * S + r % S;
      };
      for (int64_t b = 0; b < B; ++b) {
        double pr[TILE_MAX_OPS] = {0, 0, 0}, pi[TILE_MAX_OPS] = {0, 0, 0};
        for (int kind = 0; kind < N_KINDS; ++kind) {
          const uint16_t* cw = f.codes[pass].data() + ((size_t)r * f.WT[pass] + f.word0[pass][kind]) * 8;
          for (int j = 0; j < f.W[pass][kind]; ++j) {
            const uint16_t code = cw[j];
            if (code == 0) break;  // lists are packed at the front
            if (kind <= K_P) {
              const Entry16& e = f.tab16[code];
              const int64_t c = tile_col(e.off);
              const double xr = x_reim[2 * ((size_t)c * B + b)], xi = x_reim[2 * ((size_t)c * B + b) + 1];
              if (kind == K_P) {
                const double v2 = e.neg2 ? -e.v : e.v;
                pr[f.p1] += e.v * xr;
                pi[f.p1] += e.v * xi;
                pr[f.p2] += v2 * xr;
                pi[f.p2] += v2 * xi;
              } else {
                pr[kind] += e.v * xr;
                pi[kind] += e.v * xi;
              }
            } else {
              const Entry32& e = f.tab32[code];
              const int64_t c = kind == K_O ? r + e.off : tile_col(e.off);
              const double xr = x_reim[2 * ((size_t)c * B + b)], xi = x_reim[2 * ((size_t)c * B + b) + 1];
              for (int l = 0; l < f.n_ops; ++l) {  // absent operators have v = 0 (the kernels multiply unconditionally)
                pr[l] += e.v(l) * xr;
                pi[l] += e.v(l) * xi;
              }
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
        if (pass == 0) {
          t[2 * ((size_t)r * B + b)] = hr;
          t[2 * ((size_t)r * B + b) + 1] = hi;
        } else {
          y_reim[2 * ((size_t)r * B + b)] = t[2 * ((size_t)r * B + b)] + hr;
          y_reim[2 * ((size_t)r * B + b) + 1] = t[2 * ((size_t)r * B + b) + 1] + hi;
        }
      }
    }
  }
}

}  // namespace qptile
