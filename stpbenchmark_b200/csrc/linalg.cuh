// Team-cooperative dense kernels on small fp64 matrices held in shared memory.
//
// All matrices are row-major with an ODD leading dimension `ld` (in doubles): a warp that
// walks a column then touches 16 distinct bank pairs per half-warp, so row and column
// accesses are both conflict-free for 8-byte words.
#pragma once
#include "team.cuh"

namespace stp {

STP_HD int odd_ld(int n) { return n | 1; }

// ---------------------------------------------------------------------------------------
// Symmetric sweep (Goodnight 1979) of the leading `npiv` pivots of an m x m symmetric matrix
// whose LOWER triangle (i >= j) lives in A.  With the bordered input
//     [ K   r ]            [ -K^-1      .      ]
//     [ r^T 0 ]   it ends  [ (K^-1 r)^T -r^T K^-1 r ]
// i.e. rows < npiv hold -K^-1, the border row holds alpha = K^-1 r and the corner holds
// minus the quadratic form.  piv[k] receives the k-th pivot (= the LDL^T diagonal), so
// logdet K = sum log piv.  One barrier per pivot; every step updates the whole triangle, so
// the work is perfectly parallel (no shrinking trailing block, no serial substitution).
// The strict upper triangle of A is never touched.
// `c0`, `c1`: two scratch vectors of m doubles (ping-pong copies of the pivot column).
// Returns 0, or 1 + k if pivot k is not positive / not finite (matrix not PD).
// ---------------------------------------------------------------------------------------
template <class Team>
STP_HD int sweep_bordered(const Team& tm, double* A, int ld, int m, int npiv, double* c0, double* c1,
                          double* piv) {
  for (int i = tm.rank(); i < m; i += tm.size()) c0[i] = A[i * ld];
  tm.sync();
  double* cur = c0;
  double* nxt = c1;
  const int W = tm.warp_width();
  for (int k = 0; k < npiv; ++k) {
    const double p = cur[k];
    if (!(p > 0.0) || !(p <= 1.79e308)) return 1 + k;  // uniform: every thread reads the same p
    if (tm.rank() == 0) piv[k] = p;
    const double invp = 1.0 / p;
    for (int i = tm.warp(); i < m; i += tm.nwarps()) {
      double* row = A + i * ld;
      if (i == k) {  // pivot row (warp-uniform branch): a_kj / p, and -1/p on the diagonal
        for (int j = tm.lane(); j <= i; j += W) row[j] = (j == k) ? -invp : cur[j] * invp;
      } else {       // generic rows, branch-free body: a_ij - a_ik a_kj / p, pivot column gets a_ik / p
        const double ti = cur[i] * invp;
        const bool next_row = (i == k + 1);
        for (int j = tm.lane(); j <= i; j += W) {
          double a = fma(-ti, cur[j], row[j]);
          a = (j == k) ? ti : a;
          row[j] = a;
          if (j == k + 1) nxt[i] = a;
          if (next_row) nxt[j] = a;
        }
      }
    }
    tm.sync();
    double* t = cur; cur = nxt; nxt = t;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------
// In-place Cholesky of the lower triangle: A = L L^T.  Right-looking with the column scaling
// deferred to the end (the loop works on l_ij * l_jj), hence one barrier per column.
// `sd` : n doubles of scratch.  Returns 0 or 1 + j for a non-positive pivot.
// ---------------------------------------------------------------------------------------
template <class Team>
STP_HD int cholesky_lower(const Team& tm, double* A, int ld, int n, double* sd) {
  const int W = tm.warp_width();
  for (int j = 0; j < n; ++j) {
    tm.sync();
    const double dj = A[j * ld + j];
    if (!(dj > 0.0) || !(dj <= 1.79e308)) return 1 + j;
    const double inv = 1.0 / dj;
    for (int i = j + 1 + tm.warp(); i < n; i += tm.nwarps()) {
      double* row = A + i * ld;
      const double lij = row[j] * inv;
      for (int k = j + 1 + tm.lane(); k <= i; k += W) row[k] -= lij * A[k * ld + j];
    }
  }
  tm.sync();
  for (int j = tm.rank(); j < n; j += tm.size()) sd[j] = sqrt(A[j * ld + j]);
  tm.sync();
  for (int i = tm.warp(); i < n; i += tm.nwarps()) {
    double* row = A + i * ld;
    for (int j = tm.lane(); j <= i; j += W) row[j] = (j == i) ? sd[j] : row[j] / sd[j];
  }
  tm.sync();
  return 0;
}

// ---------------------------------------------------------------------------------------
// In-place inverse of a lower-triangular matrix (unblocked, right-to-left columns):
//   Linv[j][j] = 1 / L[j][j];  Linv[j+1:, j] = -Linv[j+1:, j+1:] L[j+1:, j] / L[j][j].
// ---------------------------------------------------------------------------------------
template <class Team>
STP_HD void tri_inverse_lower(const Team& tm, double* A, int ld, int n, double* tmp) {
  for (int j = n - 1; j >= 0; --j) {
    const double invd = 1.0 / A[j * ld + j];
    for (int i = j + 1 + tm.rank(); i < n; i += tm.size()) {
      const double* row = A + i * ld;
      double s0 = 0.0, s1 = 0.0;
      int k = j + 1;
      for (; k + 1 <= i; k += 2) {
        s0 += row[k] * A[k * ld + j];
        s1 += row[k + 1] * A[(k + 1) * ld + j];
      }
      if (k <= i) s0 += row[k] * A[k * ld + j];
      tmp[i] = -(s0 + s1) * invd;
    }
    tm.sync();
    for (int i = j + tm.rank(); i < n; i += tm.size()) A[i * ld + j] = (i == j) ? invd : tmp[i];
    tm.sync();
  }
}

}  // namespace stp
