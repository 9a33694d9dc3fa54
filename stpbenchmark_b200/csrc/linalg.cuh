// Team-cooperative dense kernels on small fp64 matrices held in shared memory.
//
// All matrices are row-major with an ODD leading dimension `ld` (in doubles): a warp that
// walks a column then touches 16 distinct bank pairs per half-warp, so row and column
// accesses are both conflict-free for 8-byte words.
#pragma once
#include "team.cuh"

namespace stp {

STP_HD int odd_ld(int n) { return n | 1; }

// Row offset of a lower-triangular matrix: square storage (row i at i * ld) or packed rows (row i at i (i + 1) / 2,
// half the shared memory; rows stay contiguous, which is all the sweep / Cholesky / inverse / pool pass need).
template <bool TRI>
STP_HD int row_off(int i, int ld) { return TRI ? (i * (i + 1)) / 2 : i * ld; }

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
// Returns 0, or 1 + k if pivot k is not positive / not finite (matrix not PD).
//
// Implemented in its rank-2 (pivot-pair) form: pivots k and k + 1 are eliminated together,
//     a_ij <- a_ij - [a_ik a_ik'] Pinv [a_kj a_k'j]^T ,  a_iK <- a_iK Pinv ,  A_KK <- -Pinv ,   P = A_KK (2 x 2),
// which equals two successive single sweeps but touches every element once per PAIR: one load / two DFMA /
// one store per element and half the barriers.  Four ping-pong vectors hold the two pivot columns of this step
// (c0, c1: values BEFORE the step) and of the next one (published by the owners of columns k + 2, k + 3).
// An odd last pivot is handled by a single sweep step.  CQ > 0: every lane keeps the pivot-column entries of
// its CQ fixed columns (lane + q W) in registers; CQ == 0: generic strided loops (any m, any team).
// v0 .. v3: four scratch vectors of m doubles.
// ---------------------------------------------------------------------------------------
template <int CQ, bool TRI, class Team>
STP_HD int sweep_bordered_pair(const Team& tm, double* A, int ld, int m, int npiv, double* v0, double* v1, double* v2,
                               double* v3, double* piv) {
  const int W = tm.warp_width(), lane = tm.lane(), nw = tm.nwarps();
  constexpr int NQ = CQ > 0 ? CQ : 1;
  double* c0 = v0; double* c1 = v1; double* n0 = v2; double* n1 = v3;
  tm.sync();   // the caller's writes to A (and the previous users of v0 .. v3) must be complete
  // columns 0 and 1 of the initial matrix (symmetric fill: entry i of column c is A[max(i,c)][min(i,c)])
  for (int i = tm.rank(); i < m; i += tm.size()) {
    c0[i] = A[row_off<TRI>(i, ld)];
    if (m > 1) c1[i] = (i >= 1) ? A[row_off<TRI>(i, ld) + 1] : A[row_off<TRI>(1, ld)];
  }
  tm.sync();
  int k = 0;
  for (; k + 1 < npiv; k += 2) {
    const double p00 = c0[k], p01 = c0[k + 1], p11 = c1[k + 1];
    if (!(p00 > 0.0) || !(p00 <= 1.79e308)) return 1 + k;
    const double s = p11 - p01 * (p01 / p00);
    if (!(s > 0.0) || !(s <= 1.79e308)) return 2 + k;
    if (tm.rank() == 0) { piv[k] = p00; piv[k + 1] = s; }
    const double idet = 1.0 / (p00 * s);
    const double q00 = p11 * idet, q01 = -p01 * idet, q11 = p00 * idet;
    double cj0[NQ], cj1[NQ];
    if (CQ > 0) {
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int j = lane + q * W;
        cj0[q] = (j < m) ? c0[j] : 0.0;
        cj1[q] = (j < m) ? c1[j] : 0.0;
      }
    }
    if (CQ > 0) {
      // Register-cached columns: RU rows per iteration, all loads of the RU x CQ tile issued before the first DFMA,
      // so that the shared-memory latency is paid once per RU rows instead of once per row (the loop is
      // latency-bound: ~3 elements per lane per row).
      constexpr int RU = 1;   // measured on B200: RU = 2 / 4 are slower (20.0k / 17.8k vs 20.9k it/s): register pressure, no gain
      for (int ib = tm.warp(); ib < m; ib += nw * RU) {
        double a[RU][NQ], t0[RU], t1[RU];
        int ii[RU];
        bool pr[RU];
#pragma unroll
        for (int r = 0; r < RU; ++r) {
          const int i = ib + r * nw;
          ii[r] = i;
          const int ic = (i < m) ? i : 0;
          const double ci0 = c0[ic], ci1 = c1[ic];
          const bool prow0 = (i == k), prow1 = (i == k + 1);
          pr[r] = prow0 || prow1;
          t0[r] = prow0 ? q00 : (prow1 ? q01 : ci0 * q00 + ci1 * q01);
          t1[r] = prow0 ? q01 : (prow1 ? q11 : ci0 * q01 + ci1 * q11);
          const double* row = A + row_off<TRI>(ic, ld);
#pragma unroll
          for (int q = 0; q < NQ; ++q) {
            const int j = lane + q * W;
            a[r][q] = (i < m && j <= i) ? row[j] : 0.0;
          }
        }
#pragma unroll
        for (int r = 0; r < RU; ++r) {
          const int i = ii[r];
          if (i < m) {
            double* row = A + row_off<TRI>(i, ld);
            const bool nrow0 = (i == k + 2), nrow1 = (i == k + 3);
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
              const int j = lane + q * W;
              if (j <= i) {
                double v;
                if (pr[r]) v = (j == k) ? -t0[r] : ((j == k + 1) ? -t1[r] : t0[r] * cj0[q] + t1[r] * cj1[q]);
                else v = (j == k) ? t0[r] : ((j == k + 1) ? t1[r] : fma(-t1[r], cj1[q], fma(-t0[r], cj0[q], a[r][q])));
                row[j] = v;
                if (j == k + 2) n0[i] = v;
                if (j == k + 3) n1[i] = v;
                if (nrow0) n0[j] = v;
                if (nrow1) n1[j] = v;
              }
            }
          }
        }
      }
    } else {
      for (int i = tm.warp(); i < m; i += nw) {
        double* row = A + row_off<TRI>(i, ld);
        const double ci0 = c0[i], ci1 = c1[i];
        const bool prow0 = (i == k), prow1 = (i == k + 1);
        // generic rows: t = a_iK Pinv; pivot rows: the matching row of Pinv (so that t . c_j = (Pinv a_Kj)_row)
        const double t0 = prow0 ? q00 : (prow1 ? q01 : ci0 * q00 + ci1 * q01);
        const double t1 = prow0 ? q01 : (prow1 ? q11 : ci0 * q01 + ci1 * q11);
        const bool pr = prow0 || prow1;
        const bool nrow0 = (i == k + 2), nrow1 = (i == k + 3);
        for (int j = lane; j <= i; j += W) {
          const double a0 = c0[j], a1 = c1[j];
          double a;
          if (pr) a = (j == k) ? -t0 : ((j == k + 1) ? -t1 : t0 * a0 + t1 * a1);          // pivot rows
          else a = (j == k) ? t0 : ((j == k + 1) ? t1 : fma(-t1, a1, fma(-t0, a0, row[j])));
          row[j] = a;
          if (j == k + 2) n0[i] = a;
          if (j == k + 3) n1[i] = a;
          if (nrow0) n0[j] = a;
          if (nrow1) n1[j] = a;
        }
      }
    }
    tm.sync();
    double* t;
    t = c0; c0 = n0; n0 = t;
    t = c1; c1 = n1; n1 = t;
  }
  if (k < npiv) {  // odd remainder: one single-pivot step on column c0
    const double p = c0[k];
    if (!(p > 0.0) || !(p <= 1.79e308)) return 1 + k;
    if (tm.rank() == 0) piv[k] = p;
    const double invp = 1.0 / p;
    for (int i = tm.warp(); i < m; i += nw) {
      double* row = A + row_off<TRI>(i, ld);
      const double ti = c0[i] * invp;
      for (int j = lane; j <= i; j += W) {
        double a;
        if (i == k) a = (j == k) ? -invp : c0[j] * invp;
        else a = (j == k) ? ti : fma(-ti, c0[j], row[j]);
        row[j] = a;
      }
    }
    tm.sync();
  }
  return 0;
}


// Dispatcher used by the fit closures: pair sweep with register-cached columns when the matrix fits 3 column
// slots per lane, generic pair sweep otherwise.
template <bool TRI = false, class Team>
STP_HD int sweep_bordered4(const Team& tm, double* A, int ld, int m, int npiv, double* v0, double* v1, double* v2,
                           double* v3, double* piv) {
  if (tm.warp_width() > 1) {
    if (m <= 1 * tm.warp_width()) return sweep_bordered_pair<1, TRI>(tm, A, ld, m, npiv, v0, v1, v2, v3, piv);
    if (m <= 2 * tm.warp_width()) return sweep_bordered_pair<2, TRI>(tm, A, ld, m, npiv, v0, v1, v2, v3, piv);
    if (m <= 3 * tm.warp_width()) return sweep_bordered_pair<3, TRI>(tm, A, ld, m, npiv, v0, v1, v2, v3, piv);
  }
  return sweep_bordered_pair<0, TRI>(tm, A, ld, m, npiv, v0, v1, v2, v3, piv);
}

// ---------------------------------------------------------------------------------------
// Register-tiled contraction  out(i, k) = sum_{r = r0}^{r1 - 1} L(i, r) * [scale(r)] * R(r, k)  with
//   L(i, r) = Lp[i * l_si + r * l_sr]   (LT = false: l_sr == 1, row-major;  LT = true: l_si == 1, read down a column)
//   R(r, k) = Rp[r * ld_r + k]
// A warp owns RB consecutive rows, a lane CQ columns 32 apart: RB + CQ loads feed RB * CQ DFMAs with RB * CQ
// independent accumulators; L(i, r) is a warp-wide broadcast, R(r, k) is consecutive across lanes.
// `range(i_lo, i_hi, k_lo, k_hi, r0, r1)` gives one reduction range per tile: operands must be ZERO (not just
// ignored) outside their own support inside it.  `store(i, k, v)` is called for every in-bounds (i, k).
// The reduction loop is unrolled by 4 with running pointers, so the inner body is loads + DFMA only.
// ---------------------------------------------------------------------------------------
template <int RB, int CQ, bool LT, bool SCALE, class Team, class RangeF, class StoreF>
STP_HD void contract_p(const Team& tm, int n_rows, int n_cols, const double* Lp, int ld_l, const double* Rp, int ld_r,
                       const double* scale, RangeF range, StoreF store) {
  const int W = tm.warp_width(), nw = tm.nwarps();
  const int l_si = LT ? 1 : ld_l, l_sr = LT ? ld_l : 1;
  for (int ib = tm.warp() * RB; ib < n_rows; ib += nw * RB) {
    int ri[RB];
#pragma unroll
    for (int a = 0; a < RB; ++a) ri[a] = (ib + a < n_rows) ? ib + a : n_rows - 1;
    for (int kb = 0; kb < n_cols; kb += W * CQ) {
      int kk[CQ];
#pragma unroll
      for (int q = 0; q < CQ; ++q) { const int k = kb + tm.lane() + q * W; kk[q] = (k < n_cols) ? k : n_cols - 1; }
      int r0 = 0, r1 = 0;
      range(ib, ri[RB - 1], kb, (kb + W * CQ - 1 < n_cols) ? kb + W * CQ - 1 : n_cols - 1, r0, r1);
      double acc[RB][CQ];
#pragma unroll
      for (int a = 0; a < RB; ++a)
#pragma unroll
        for (int q = 0; q < CQ; ++q) acc[a][q] = 0.0;
      const double* lp[RB];
      const double* rp[CQ];
#pragma unroll
      for (int a = 0; a < RB; ++a) lp[a] = Lp + (size_t)ri[a] * l_si + (size_t)r0 * l_sr;
#pragma unroll
      for (int q = 0; q < CQ; ++q) rp[q] = Rp + (size_t)r0 * ld_r + kk[q];
      const double* sp = SCALE ? scale + r0 : nullptr;
      int r = r0;
      for (; r + 4 <= r1; r += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          double l[RB], rr[CQ];
#pragma unroll
          for (int a = 0; a < RB; ++a) l[a] = lp[a][u * l_sr];
#pragma unroll
          for (int q = 0; q < CQ; ++q) rr[q] = rp[q][u * ld_r];
          if (SCALE) {
            const double sc = sp[u];
#pragma unroll
            for (int q = 0; q < CQ; ++q) rr[q] *= sc;
          }
#pragma unroll
          for (int a = 0; a < RB; ++a)
#pragma unroll
            for (int q = 0; q < CQ; ++q) acc[a][q] = fma(l[a], rr[q], acc[a][q]);
        }
#pragma unroll
        for (int a = 0; a < RB; ++a) lp[a] += 4 * l_sr;
#pragma unroll
        for (int q = 0; q < CQ; ++q) rp[q] += 4 * ld_r;
        if (SCALE) sp += 4;
      }
      for (; r < r1; ++r) {
        double l[RB], rr[CQ];
#pragma unroll
        for (int a = 0; a < RB; ++a) { l[a] = *lp[a]; lp[a] += l_sr; }
#pragma unroll
        for (int q = 0; q < CQ; ++q) { rr[q] = *rp[q]; rp[q] += ld_r; }
        if (SCALE) {
          const double sc = *sp++;
#pragma unroll
          for (int q = 0; q < CQ; ++q) rr[q] *= sc;
        }
#pragma unroll
        for (int a = 0; a < RB; ++a)
#pragma unroll
          for (int q = 0; q < CQ; ++q) acc[a][q] = fma(l[a], rr[q], acc[a][q]);
      }
#pragma unroll
      for (int a = 0; a < RB; ++a)
#pragma unroll
        for (int q = 0; q < CQ; ++q) {
          const int k = kb + tm.lane() + q * W;
          if (ib + a < n_rows && k < n_cols) store(ib + a, k, acc[a][q]);
        }
    }
  }
}


// ---------------------------------------------------------------------------------------
// Fused Cholesky + triangular inverse, in place, ONE barrier per column.  With the deferred-scaling Cholesky
// (column k holds l_ik l_kk, d_k = l_kk^2) the multiplier f_i = a_ik / d_k = L[i][k] / L[k][k] drives both the
// trailing update a_ij -= f_i a_jk (j > k) and the right-looking inverse update x_ic -= f_i x_kc (c < k),
// x_ik = -f_i, so step k is one uniform pass over the whole row i > k:   a_i,: -= f_i * r_k,  r_k = row/column k
// (symmetric fill, published by its owners during step k - 1 like the sweep's pivot column).
// On exit A (lower) = L^-1; `Lout` (may be null; leading dimension ldo, any memory space) receives L (lower part
// only); `sd` (n doubles) = diag(L).  `v0`, `v1`: n doubles each.  Returns 0 or 1 + k for a non-positive pivot.
// ---------------------------------------------------------------------------------------
template <bool TRI = false, class Team>
STP_HD int cholesky_inverse_fused(const Team& tm, double* A, int ld, int n, double* v0, double* v1, double* sd,
                                  double* Lout, int ldo) {
  const int W = tm.warp_width(), lane = tm.lane(), nw = tm.nwarps();
  double* cur = v0;
  double* nxt = v1;
  tm.sync();   // the caller's writes to A (and the previous users of v0 / v1) must be complete
  for (int i = tm.rank(); i < n; i += tm.size()) cur[i] = A[row_off<TRI>(i, ld)];
  tm.sync();
  for (int k = 0; k < n; ++k) {
    const double dk = cur[k];
    if (!(dk > 0.0) || !(dk <= 1.79e308)) return 1 + k;
    const double inv = 1.0 / dk;
    const double lkk = sqrt(dk);
    const double isd = 1.0 / lkk;
    if (tm.rank() == 0) { sd[k] = lkk; if (Lout) Lout[k * ldo + k] = lkk; }
    for (int i = k + 1 + tm.warp(); i < n; i += nw) {
      double* row = A + row_off<TRI>(i, ld);
      const double aik = cur[i];                 // = A[i][k] before this step
      const double f = aik * inv;
      const bool next_row = (i == k + 1);
      if (Lout && lane == 0) Lout[i * ldo + k] = aik * isd;
      for (int c = lane; c <= i; c += W) {
        double a = (c == k) ? -f : fma(-f, cur[c], row[c]);
        row[c] = a;
        if (c == k + 1) nxt[i] = a;
        if (next_row) nxt[c] = a;
      }
    }
    tm.sync();
    double* t = cur; cur = nxt; nxt = t;
  }
  // rows hold the unscaled inverse (unit diagonal implied): scale row i by 1 / l_ii
  for (int i = tm.warp(); i < n; i += nw) {
    double* row = A + row_off<TRI>(i, ld);
    const double di = 1.0 / sd[i];
    for (int c = lane; c <= i; c += W) row[c] = ((c == i) ? 1.0 : row[c]) * di;
  }
  tm.sync();
  return 0;
}

}  // namespace stp
