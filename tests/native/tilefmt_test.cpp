// Test hook for the host-side builder of the two-pass tile format (csrc/tile_format.h): builds the
// format from a merged CSR matrix and applies it on the CPU exactly the way the kernels traverse it.
// Compiled by tests/test_tile_format.py with g++ (no CUDA needed).
#include "tile_format.h"

extern "C" int tilefmt_apply(long long n, int n_ops, const unsigned* mptr, const unsigned* colop, const double* val,
                             long long B, const double* u, const double* x, double* y, long long* stats, int S_forced) {
  qptile::TileFormat f;
  if (!qptile::build(f, n, n_ops, mptr, colop, val, S_forced)) return -1;
  qptile::apply_host(f, B, u, x, y);
  stats[0] = f.S;
  stats[1] = f.NH;
  stats[2] = f.WT[0];
  stats[3] = f.WT[1];
  stats[4] = (long long)(f.tab16.size() + f.tab32.size());
  stats[5] = f.n_A();
  stats[6] = f.n_B();
  stats[7] = f.n_O();
  stats[8] = f.n_diag;
  stats[9] = f.imag_ops;
  stats[10] = f.count[0][qptile::K_P] + f.count[1][qptile::K_P];
  stats[11] = f.count[0][qptile::K_G] + f.count[1][qptile::K_G];
  return 0;
}

extern "C" int tilefmt_choose_split(long long n) { return qptile::choose_split(n); }
