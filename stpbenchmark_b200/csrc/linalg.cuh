// Team-cooperative dense kernels on small fp64 matrices held in shared memory.
//
// All matrices are row-major with an ODD leading dimension `ld` (in doubles): a warp that
// walks a column then touches 16 distinct bank pairs per half-warp, so row and column
// accesses are both conflict-free for 8-byte words.
#pragma once
#include "team.cuh"

namespace stp {

// Reciprocal / reciprocal square root of a positive, normal double for the dependent chains (pivots, BFGS
// curvatures): hardware seed (rcp / rsqrt.approx.ftz.f64, ~20 bits) + two Newton steps = 1 MUFU + 4 - 6 DFMA (~1 ulp)
// instead of the ~11-instruction IEEE division / sqrt + division with their special-case branches.  Only for
// arguments that were checked to be positive and finite.  Host code (emulation) uses the IEEE forms.
STP_HD double rcp_pos(double p) {
#if defined(__CUDA_ARCH__) && !defined(STP_IEEE_PIVOT_RCP)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(p));
  r = fma(r, fma(-p, r, 1.0), r);
  r = fma(r, fma(-p, r, 1.0), r);
  return r;
#else
  return 1.0 / p;
#endif
}
STP_HD double rsqrt_pos(double p) {
#if defined(__CUDA_ARCH__) && !defined(STP_IEEE_PIVOT_RCP)
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(p));
  const double hp = 0.5 * p;
  r = fma(r, fma(-hp * r, r, 0.5), r);
  r = fma(r, fma(-hp * r, r, 0.5), r);
  return r;
#else
  return 1.0 / sqrt(p);
#endif
}

STP_HD int odd_ld(int n) { return n | 1; }

// Extra diagonal jitter of attempt k of a Cholesky: what the reference's factorisations do through
// linear_operator's psd_safe_cholesky in fp64 (plain, then +1e-8, +1e-7, +1e-6 on the diagonal, then NotPSDError).
constexpr int kPsdSafeTries = 4;
STP_HD double psd_safe_jitter(int attempt) { return attempt == 0 ? 0.0 : (attempt == 1 ? 1e-8 : (attempt == 2 ? 1e-7 : 1e-6)); }

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
    const double s = p11 - p01 * (p01 * rcp_pos(p00));
    if (!(s > 0.0) || !(s <= 1.79e308)) return 2 + k;
    if (tm.rank() == 0) { piv[k] = p00; piv[k + 1] = s; }
    const double idet = rcp_pos(p00 * s);
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

// fp64 tensor-core MMA, D(8x8) += A(8x4) B(4x8):  lane l holds  a = A[l / 4][l % 4],  b = B[l % 4][l / 4],
// d0 = D[l / 4][2 (l % 4)], d1 = D[l / 4][2 (l % 4) + 1]   (DMMA in SASS; there is no fp64 tcgen05 path).
#if defined(__CUDACC__)
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}
#endif

// ---------------------------------------------------------------------------------------
// Blocked bordered sweep on the tensor cores (device only).  Same contract as sweep_bordered_pair (square storage,
// lower triangle + diagonal of the m x m matrix updated in place, strict upper triangle untouched, pivots 0 ..
// npiv - 1, rows >= npiv are border rows), but the pivots are taken EIGHT at a time.  With C = A[:, K] (m x 8,
// symmetric fill), D = C[K, :] = Lu Delta Lu^T (unit lower Lu, Delta = the 8 scalar pivots) and X = Lu^-1:
//     W = C X^T,  V = W Delta^-1,  T = V X = C D^-1,
//     A[K, K] <- -X^T Delta^-1 X,   A[i, K] <- T[i, :] (i not in K),   A[i, j] <- A[i, j] - V[i, :] . W[j, :] otherwise.
// (The explicit D^-1 never multiplies the trailing matrix: forming T C^T from an inverted block loses
// cond(D) eps -- 4 to 6 digits on long-lengthscale kernels -- while the V W^T form is as accurate as the
// pivot-by-pivot sweep; see tests/test_blocked_sweep_model.py.)
// Per block step:
//   (A) every warp eliminates the 8 x 8 block in registers (accumulator-fragment layout, warp shuffles only: no
//       barrier on the critical path of 8 dependent pivots), giving X and 1 / Delta;
//   (B) W, V, T by DMMA, 8 rows per warp; W replaces C in its panel (V = W / Delta is rebuilt from it where it is
//       needed), T is written straight into the pivot columns / rows of A (and into the next panel where it
//       belongs);                                                                                    -- barrier --
//   (C) rank-8 update of the lower triangle in 16 x 16 tiles, 8 DMMA per tile; the store pass also collects the
//       NEXT panel (columns K + 8, symmetric fill).                                                -- barrier --
// i.e. 2 barriers per 8 pivots instead of 4, and ~8x fewer issued instructions per MAC than the pair sweep.
// A trailing partial block is padded with identity pivots (its panel columns are zero).
// `scratch`: sweep_blocked_scratch_doubles(m) doubles of shared memory.  Returns 0 or 1 + k for a bad pivot.
// ---------------------------------------------------------------------------------------
constexpr int kPanelLd = 12;   // 8 panel columns + 4 pad doubles: conflict-free fragment loads (24 g + 2 q words)
STP_HD size_t sweep_blocked_scratch_doubles(int m) { return (size_t)2 * m * kPanelLd; }

// By-product of the blocked sweep of a positive-definite matrix (no border): its Cholesky factor and the inverse of
// that factor.  Right before block K is eliminated, the panel rows c < K hold (A[K, :K] A[:K, :K]^-1)^T =
// (L[K, :K] L[:K, :K]^-1)^T, so with D = Lu Delta Lu^T the product W = C Lu^-T already is, up to Delta^-1/2,
//     L[i, K_b]       =  W[i][b] / sqrt(Delta_b)      (i >= K_b),
//     L^-1[K_b, c]    = -W[c][b] / sqrt(Delta_b)      (c <  K),          L^-1[K_g, K_b] = X[g][b] / sqrt(Delta_g).
// The sweep therefore replaces the column-by-column fused Cholesky + inverse (n barriers, one pass over the triangle
// each) by n / 8 block steps.  Targets may be null; only the lower triangles (upper for LiT) are written.
struct SweepCholExport {
  double* L;     // lower Cholesky factor, row-major, leading dimension ld
  double* Li;    // L^-1 (lower)
  double* LiT;   // (L^-1)^T (upper)
  int ld;
};

#if defined(__CUDACC__)
// FACTOR_ONLY (with EXPORT): the caller wants L and L^-1 but not the swept matrix, so the trailing update skips the
// tiles that lie entirely above the pivot block -- the already inverted top-left part, which no later panel reads
// (a third of the update's traffic over the whole sweep).  A is then garbage on exit.
template <bool EXPORT = false, bool FACTOR_ONLY = false>
__device__ __forceinline__ int sweep_bordered_blocked(const CtaTeam& tm, double* A, int ld, int m, int npiv,
                                                      double* scratch, double* piv,
                                                      SweepCholExport ex = SweepCholExport{nullptr, nullptr, nullptr, 0}) {
  constexpr unsigned kFull = 0xffffffffu;
  constexpr int PL = kPanelLd;
  const int lane = tm.lane(), warp = tm.warp(), nw = tm.nwarps();
  const int g = lane >> 2, q = lane & 3;
  double* C = scratch;                           // current panel; rows outside K become W in step (B)
  double* Cn = scratch + (size_t)m * PL;         // next panel
  tm.sync();   // the caller's writes to A must be complete
  for (int idx = tm.rank(); idx < m * 8; idx += tm.size()) {
    const int i = idx >> 3, b = idx & 7;
    C[i * PL + b] = (b < npiv) ? ((i >= b) ? A[i * ld + b] : A[b * ld + i]) : 0.0;
  }
  tm.sync();
  const int nT = (m + 15) >> 4, n_tiles = nT * (nT + 1) / 2;
  for (int kb = 0; kb < npiv; kb += 8) {
    const int r = (npiv - kb < 8) ? (npiv - kb) : 8;
    const int kn = kb + 8;
    STP_PHASE(-1); STP_COUNT(17);
    // ---- (A) elimination of D in registers: am = trailing block, xm = X (row g, columns 2q, 2q + 1)
    double am0, am1;
    if (g < r) {
      am0 = C[(kb + g) * PL + 2 * q];
      am1 = C[(kb + g) * PL + 2 * q + 1];
    } else {
      am0 = (2 * q == g) ? 1.0 : 0.0;
      am1 = (2 * q + 1 == g) ? 1.0 : 0.0;
    }
    double xm0 = (2 * q == g) ? 1.0 : 0.0, xm1 = (2 * q + 1 == g) ? 1.0 : 0.0;
    double ipa = 1.0, ipb = 1.0;     // 1 / Delta at columns 2q, 2q + 1
    double ipq0 = 1.0, ipq1 = 1.0;   // 1 / Delta at q, q + 4
    double ipg = 1.0;                // 1 / Delta at g (EXPORT)
    int fail = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const double dk = (k & 1) ? am1 : am0;
      const double p = __shfl_sync(kFull, dk, 4 * k + (k >> 1));
      const double rk0 = __shfl_sync(kFull, am0, 4 * k + q);
      const double rk1 = __shfl_sync(kFull, am1, 4 * k + q);
      const double xk0 = __shfl_sync(kFull, xm0, 4 * k + q);
      const double xk1 = __shfl_sync(kFull, xm1, 4 * k + q);
      const double cg = __shfl_sync(kFull, dk, 4 * g + (k >> 1));
      // a bad pivot is only recorded (no branch in the dependent chain); the garbage that follows is never used
      if (fail == 0 && (!(p > 0.0) || !(p <= 1.79e308))) fail = 1 + kb + k;
      const double ip = rcp_pos(p);
      const double f = (g > k) ? cg * ip : 0.0;
      am0 = fma(-f, rk0, am0);
      am1 = fma(-f, rk1, am1);
      xm0 = fma(-f, xk0, xm0);
      xm1 = fma(-f, xk1, xm1);
      if (2 * q == k) ipa = ip;
      if (2 * q + 1 == k) ipb = ip;
      if (q == k) ipq0 = ip;
      if (q + 4 == k) ipq1 = ip;
      if (EXPORT && g == k) ipg = ip;
      if (warp == 0 && lane == 0 && k < r) piv[kb + k] = p;
    }
    if (fail) return fail;   // identical in every warp: the whole team leaves together
    STP_PHASE(2);
    // operand fragments of X:  bw = X[g][4s + q] (B operand of C X^T),  bt = X[4s + q][g] (B operand of V X)
    double bw0, bw1, bt0, bt1;
    {
      const int src = 4 * g + (q >> 1);
      const double x0 = __shfl_sync(kFull, xm0, src), x1 = __shfl_sync(kFull, xm1, src);
      const double y0 = __shfl_sync(kFull, xm0, src + 2), y1 = __shfl_sync(kFull, xm1, src + 2);
      bw0 = (q & 1) ? x1 : x0;
      bw1 = (q & 1) ? y1 : y0;
      const int st = 4 * q + (g >> 1);
      const double u0 = __shfl_sync(kFull, xm0, st), u1 = __shfl_sync(kFull, xm1, st);
      const double z0 = __shfl_sync(kFull, xm0, st + 16), z1 = __shfl_sync(kFull, xm1, st + 16);
      bt0 = (g & 1) ? u1 : u0;
      bt1 = (g & 1) ? z1 : z0;
    }
    if (warp == 0) {   // A[K, K] <- -X^T Delta^-1 X
      double s0 = 0.0, s1 = 0.0;
      dmma884(s0, s1, bt0, bt0 * ipq0);
      dmma884(s0, s1, bt1, bt1 * ipq1);
      if (g < r) {
        double* row = A + (size_t)(kb + g) * ld + kb;
        if (2 * q <= g) row[2 * q] = -s0;
        if (2 * q + 1 <= g) row[2 * q + 1] = -s1;
      }
      if (EXPORT && g < r) {   // L^-1[K, K] = Delta^-1/2 X
        const double sg = sqrt(ipg);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int b = 2 * q + e;
          if (b <= g) {
            const double v = (e ? xm1 : xm0) * sg;
            if (ex.Li) ex.Li[(size_t)(kb + g) * ex.ld + kb + b] = v;
            if (ex.LiT) ex.LiT[(size_t)(kb + b) * ex.ld + kb + g] = v;
          }
        }
      }
    }
    const double isa = EXPORT ? sqrt(ipa) : 0.0, isb = EXPORT ? sqrt(ipb) : 0.0;
    // ---- (B) W = C X^T, V = W / Delta, T = V X, 8 rows per warp
    for (int rt = warp; rt * 8 < m; rt += nw) {
      const int i = rt * 8 + g, ic = (i < m) ? i : m - 1;
      const double a0 = C[ic * PL + q], a1 = C[ic * PL + 4 + q];
      double w0 = 0.0, w1 = 0.0;
      dmma884(w0, w1, a0, bw0);
      dmma884(w0, w1, a1, bw1);
      const double v0 = w0 * ipa, v1 = w1 * ipb;
      const int src = 4 * g + (q >> 1);
      const double e0 = __shfl_sync(kFull, v0, src), e1 = __shfl_sync(kFull, v1, src);
      const double h0 = __shfl_sync(kFull, v0, src + 2), h1 = __shfl_sync(kFull, v1, src + 2);
      const double va0 = (q & 1) ? e1 : e0, va1 = (q & 1) ? h1 : h0;
      double t0 = 0.0, t1 = 0.0;
      dmma884(t0, t1, va0, bt0);
      dmma884(t0, t1, va1, bt1);
      if (EXPORT && i < npiv) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int b = 2 * q + e, j = kb + b;
          if (b < r) {
            const double v = (e ? w1 : w0) * (e ? isb : isa);
            if (i >= j) { if (ex.L) ex.L[(size_t)i * ex.ld + j] = v; }
            if (i < kb) {
              if (ex.Li) ex.Li[(size_t)j * ex.ld + i] = -v;
              if (ex.LiT) ex.LiT[(size_t)i * ex.ld + j] = -v;
            }
          }
        }
      }
      if (i < m && !((unsigned)(i - kb) < (unsigned)r)) {   // rows of K keep C (phase (A) of other warps reads them)
        C[i * PL + 2 * q] = w0;
        C[i * PL + 2 * q + 1] = w1;
        const bool nxt = (unsigned)(i - kn) < 8u && i < npiv;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int b = 2 * q + e, j = kb + b;
          const double tv = e ? t1 : t0;
          if (b < r) {
            if (i > j) A[(size_t)i * ld + j] = tv;
            else A[(size_t)j * ld + i] = tv;
            if (nxt) Cn[j * PL + (i - kn)] = tv;
          }
        }
      }
    }
    if (kn < npiv && kn + 8 > npiv) {   // the next block is partial: its unused panel columns must read as zero
      for (int idx = tm.rank(); idx < m * 8; idx += tm.size()) {
        const int b = idx & 7;
        if (kn + b >= npiv) Cn[(idx >> 3) * PL + b] = 0.0;
      }
    }
    tm.sync();
    STP_PHASE(3);
    // ---- (C) A -= V W^T on the lower triangle outside K, 16 x 16 tiles
    for (int t = warp; t < n_tiles; t += nw) {
      int I = 0, J = t;
      while (J > I) { J -= I + 1; ++I; }
      const int ib = I << 4, jb = J << 4;
      if (FACTOR_ONLY && ib + 16 <= kb) continue;
      double c[2][2][2];
      double af[2][2], bf[2][2];
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int i = ib + 8 * a + g, ic = (i < m) ? i : m - 1;
        af[a][0] = -(C[ic * PL + q] * ipq0);        // -V = -W / Delta (bitwise what a stored -V panel would hold)
        af[a][1] = -(C[ic * PL + 4 + q] * ipq1);
        const int j = jb + 8 * a + g, jc = (j < m) ? j : m - 1;
        bf[a][0] = C[jc * PL + q];
        bf[a][1] = C[jc * PL + 4 + q];
      }
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int i = ib + 8 * a + g, j = jb + 8 * b + 2 * q;
          const double* row = A + (size_t)((i < m) ? i : 0) * ld;
          c[a][b][0] = (i < m && j <= i) ? row[j] : 0.0;
          c[a][b][1] = (i < m && j + 1 <= i) ? row[j + 1] : 0.0;
        }
#pragma unroll
      for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int b = 0; b < 2; ++b) dmma884(c[a][b][0], c[a][b][1], af[a][s], bf[b][s]);
      const int kt = kb & ~15, nt = kn & ~15;
      const bool edge = (ib == kt) || (jb == kt) || ((kn < npiv) && ((ib == nt) || (jb == nt)));
      if (!edge) {
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            const int i = ib + 8 * a + g, j = jb + 8 * b + 2 * q;
            if (i < m) {
              double* row = A + (size_t)i * ld;
              if (j <= i) row[j] = c[a][b][0];
              if (j + 1 <= i) row[j + 1] = c[a][b][1];
            }
          }
      } else {
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int i = ib + 8 * a + g, j = jb + 8 * b + 2 * q + e;
              const bool ik = (unsigned)(i - kb) < (unsigned)r, jk = (unsigned)(j - kb) < (unsigned)r;
              if (i < m && j <= i && !ik && !jk) {
                const double v = c[a][b][e];
                A[(size_t)i * ld + j] = v;
                if ((unsigned)(j - kn) < 8u && j < npiv) Cn[i * PL + (j - kn)] = v;
                if ((unsigned)(i - kn) < 8u && i < npiv) Cn[j * PL + (i - kn)] = v;
              }
            }
      }
    }
    tm.sync();
    STP_PHASE(4);
    double* sw = C; C = Cn; Cn = sw;
  }
  return 0;
}
#endif

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
// Tensor-core form of the contraction (device only): out(i, k) = sum_r L(i, r) [scale(r)] R(r, k) with the legacy
// fp64 MMA, mma.sync.aligned.m8n8k4.row.col.f64 (DMMA in SASS -- there is no fp64 tcgen05 path).  A warp owns a
// 16 x 16 output block (2 x 2 MMA tiles, 8 accumulator registers per thread); per 4-deep reduction step it loads
// 2 A fragments + 2 B fragments (one double per thread each) and issues 4 DMMA = 1024 MAC, i.e. ~7x fewer
// instructions per MAC than the 4 x 2 register-tiled FMA loop.
// Reduction ranges are widened to multiples of 4 (triangular operands are zero-filled inside the n x n square, as
// for contract_p); the one 4-step that may cross n_red loads with clamped addresses and zeroed values.
// Out-of-range output rows / columns are clamped on load and never stored.
// ---------------------------------------------------------------------------------------
#if defined(__CUDACC__)

template <bool LT, bool SCALE, class RangeF, class StoreF>
__device__ __forceinline__ void contract_mma(const CtaTeam& tm, int n_rows, int n_cols, int n_red, const double* Lp, int ld_l,
                                             const double* Rp, int ld_r, const double* scale, RangeF range,
                                             StoreF store) {
  const int lane = tm.lane(), nw = tm.nwarps();
  const int g = lane >> 2, q = lane & 3;          // fragment coordinates: row / column group and k index
  const int TR = (n_rows + 15) >> 4, TC = (n_cols + 15) >> 4;
  for (int wt = tm.warp(); wt < TR * TC; wt += nw) {
    const int ib = (wt / TC) << 4, kb = (wt % TC) << 4;
    int r0 = 0, r1 = 0;
    range(ib, (ib + 15 < n_rows) ? ib + 15 : n_rows - 1, kb, (kb + 15 < n_cols) ? kb + 15 : n_cols - 1, r0, r1);
    r0 &= ~3;
    r1 = (r1 + 3) & ~3;
    double c[2][2][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) { c[a][b][0] = 0.0; c[a][b][1] = 0.0; }
    // rows of the two A tiles and columns of the two B tiles this thread loads (clamped into the matrix)
    int ia[2], jb[2];
#pragma unroll
    for (int a = 0; a < 2; ++a) { const int i = ib + 8 * a + g; ia[a] = (i < n_rows) ? i : n_rows - 1; }
#pragma unroll
    for (int b = 0; b < 2; ++b) { const int j = kb + 8 * b + g; jb[b] = (j < n_cols) ? j : n_cols - 1; }
    const int r_full = (r1 < (n_red & ~3)) ? r1 : (n_red & ~3);
    int r = r0;
    for (; r < r_full; r += 4) {
      const int rk = r + q;
      double af[2], bf[2];
#pragma unroll
      for (int a = 0; a < 2; ++a) af[a] = LT ? Lp[(size_t)rk * ld_l + ia[a]] : Lp[(size_t)ia[a] * ld_l + rk];
#pragma unroll
      for (int b = 0; b < 2; ++b) bf[b] = Rp[(size_t)rk * ld_r + jb[b]];
      if (SCALE) {
        const double sc = scale[rk];
        bf[0] *= sc; bf[1] *= sc;
      }
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) dmma884(c[a][b][0], c[a][b][1], af[a], bf[b]);
    }
    if (r < r1) {   // the 4-step that crosses n_red
      const bool in = (r + q) < n_red;
      const int rk = in ? r + q : n_red - 1;
      double af[2], bf[2];
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const double v = LT ? Lp[(size_t)rk * ld_l + ia[a]] : Lp[(size_t)ia[a] * ld_l + rk];
        af[a] = in ? v : 0.0;
      }
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const double v = Rp[(size_t)rk * ld_r + jb[b]];
        bf[b] = in ? v : 0.0;
      }
      if (SCALE) {
        const double sc = scale[rk];
        bf[0] *= sc; bf[1] *= sc;
      }
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) dmma884(c[a][b][0], c[a][b][1], af[a], bf[b]);
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int i = ib + 8 * a + g;
        const int j = kb + 8 * b + 2 * q;
        if (i < n_rows) {
          if (j < n_cols) store(i, j, c[a][b][0]);
          if (j + 1 < n_cols) store(i, j + 1, c[a][b][1]);
        }
      }
  }
}
#endif

// ---------------------------------------------------------------------------------------
// The same contraction with the operands STAGED THROUGH SHARED MEMORY (the product path of the variational epoch).
// contract_mma lets every warp stream its own operand rows from global memory: a 16x16 tile reads 2 x 16 x n doubles,
// n^3 / 8 doubles per contraction, and every 4-step waits for an L2 round trip (the work squares of a campaign are
// L2 / HBM residents).  Here the CTA walks the output in blocks of up to 4 x 4 tiles (64 x 64); for
// each block the K range is cut into chunks of kStageRC reduction indices whose two operand panels are copied by ALL
// threads with cp.async (coalesced rows, 8-byte copies: the leading dimensions are odd) into a three-deep ring, so
// the copy of chunk c + 2 is in flight while chunk c feeds the tensor cores: one barrier per chunk, ~10x less global
// traffic per MAC, and no load latency on the DMMA chain.  Panel strides are = 4 mod 8 doubles, so
// the m8n8k4 fragment loads (4 rows x 4 consecutive doubles per half warp, or the transpose) are bank-conflict free.
// Every tile accumulates the SAME DMMA sequence over the same 4-steps as contract_mma (ranges widened to multiples
// of 4, zeros beyond n_red): the two forms are bit-identical; -DSTP_NO_STAGE selects the old one for A/B runs.
// `stage`: shared memory, contract_stage_doubles() doubles, free for the duration of the call.
// ---------------------------------------------------------------------------------------
#if !defined(STP_STAGE_MAXT)
#define STP_STAGE_MAXT 2
#endif
#if !defined(STP_STAGE_CT)
#define STP_STAGE_CT 4
#endif
#if !defined(STP_STAGE_RT)
#define STP_STAGE_RT 4
#endif
constexpr int kStageMaxT = STP_STAGE_MAXT;      // tiles per warp and block
constexpr int kStageCT = STP_STAGE_CT;          // column tiles of a block
constexpr int kStageRT = STP_STAGE_RT;          // row tiles of a block
#if !defined(STP_STAGE_RC_LOG)
#define STP_STAGE_RC_LOG 5
#endif
constexpr int kStageRCLog = STP_STAGE_RC_LOG;
constexpr int kStageRC = 1 << kStageRCLog;      // reduction indices per chunk
constexpr int kStageLdK = (kStageCT > kStageRT ? kStageCT : kStageRT) * 16 + 4;   // stride of a K-major panel (row = reduction index): = 4 mod 8
constexpr int kStageLdR = kStageRC + 4;         // stride of a row-major L panel (row = output row): 20 = 4 mod 8
constexpr int kStagePanel = (kStageRT * 16 * kStageLdR > kStageRC * kStageLdK) ? kStageRT * 16 * kStageLdR : kStageRC * kStageLdK;
constexpr int kStageDepth = 3;
STP_HD size_t contract_stage_doubles() { return (size_t)kStageDepth * 2 * kStagePanel; }

#if defined(__CUDACC__)
__device__ __forceinline__ void cp_async_f64(double* dst_shared, const double* src_global) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_shared);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src_global) : "memory");
}
__device__ __forceinline__ void cp_async_2f64(double* dst_shared, const double* src_global) {   // both 16-byte aligned
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_shared);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src_global) : "memory");
}
// two consecutive doubles of which the first `nvalid` (0, 1, 2) exist; the others are zero-filled
__device__ __forceinline__ void cp_async_pair(double* dst, const double* src, int nvalid) {
  if (nvalid >= 2) cp_async_2f64(dst, src);
  else {
    if (nvalid == 1) cp_async_f64(dst, src); else dst[0] = 0.0;
    dst[1] = 0.0;
  }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// WIDE: 16-byte copies (two doubles per cp.async: half the LSU instructions, which are what bounds the 8-byte form);
// needs even leading dimensions and 16-byte aligned operand bases / stage.
template <bool LT, bool SCALE, bool WIDE = false, class RangeF, class StoreF>
__device__ __forceinline__ void contract_staged(const CtaTeam& tm, int n_rows, int n_cols, int n_red, const double* Lp,
                                                int ld_l, const double* Rp, int ld_r, const double* scale, RangeF range,
                                                StoreF store, double* stage) {
  const int lane = tm.lane(), nw = tm.nwarps(), wid = tm.warp(), tid = tm.rank(), nt = tm.size();
  const int g = lane >> 2, q = lane & 3;
  const int TR = (n_rows + 15) >> 4, TC = (n_cols + 15) >> 4;
  // block shape: <= 4 x kStageCT tiles, a warp owns <= kStageMaxT of them, both edges evened out over the matrix
  const int cap_t = kStageMaxT * nw;
  const int ncb = (TC + kStageCT - 1) / kStageCT, BTC = (TC + ncb - 1) / ncb;
  int btr_max = cap_t / BTC; btr_max = btr_max > kStageRT ? kStageRT : (btr_max < 1 ? 1 : btr_max);
  const int nrb = (TR + btr_max - 1) / btr_max, BTR = (TR + nrb - 1) / nrb;
  for (int rb = 0; rb < TR; rb += BTR)
    for (int cb = 0; cb < TC; cb += BTC) {
      const int btr = (TR - rb < BTR) ? TR - rb : BTR, btc = (TC - cb < BTC) ? TC - cb : BTC;
      const int ntile = btr * btc;
      // the block's reduction range: union over its tiles (every thread computes it: integer work only)
      int R0 = 0x7fffffff, R1 = 0;
      for (int t = 0; t < ntile; ++t) {
        const int ib = (rb + t / btc) << 4, kb = (cb + t % btc) << 4;
        int r0 = 0, r1 = 0;
        range(ib, (ib + 15 < n_rows) ? ib + 15 : n_rows - 1, kb, (kb + 15 < n_cols) ? kb + 15 : n_cols - 1, r0, r1);
        r0 &= ~3; r1 = (r1 + 3) & ~3;
        if (r1 > r0) { R0 = r0 < R0 ? r0 : R0; R1 = r1 > R1 ? r1 : R1; }
      }
      // this warp's tiles (a block with an empty range still stores its zeros: triangular outputs rely on it)
      int tib[kStageMaxT], tkb[kStageMaxT], tr0[kStageMaxT], tr1[kStageMaxT];
      double c[kStageMaxT][2][2][2];
#pragma unroll
      for (int m = 0; m < kStageMaxT; ++m) {
        const int t = wid + m * nw;
        tr0[m] = 0; tr1[m] = 0; tib[m] = 0; tkb[m] = 0;
        if (t < ntile) {
          tib[m] = (t / btc) << 4; tkb[m] = (t % btc) << 4;      // offsets inside the block
          const int ib = (rb << 4) + tib[m], kb = (cb << 4) + tkb[m];
          int r0 = 0, r1 = 0;
          range(ib, (ib + 15 < n_rows) ? ib + 15 : n_rows - 1, kb, (kb + 15 < n_cols) ? kb + 15 : n_cols - 1, r0, r1);
          tr0[m] = r0 & ~3; tr1[m] = (r1 + 3) & ~3;
        }
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int b = 0; b < 2; ++b) { c[m][a][b][0] = 0.0; c[m][a][b][1] = 0.0; }
      }
      const int i0 = rb << 4, k0 = cb << 4, rows = btr << 4, cols = btc << 4;
      const int nchunk = (R1 > R0) ? (R1 - R0 + kStageRC - 1) / kStageRC : 0;
      // Copy plan of this thread, fixed for the block: in a K-major panel (row = reduction index) the thread owns
      // column `tid & (pw - 1)` of the rows (tid >> sh) + j (nt >> sh), the width padded to a power of two; in the
      // row-major L panel it owns reduction index tid & 15 of the rows (tid >> 4) + j (nt >> 4).  Per chunk only the
      // two source pointers move.
      const int shl = rows > 64 ? 7 : (rows > 32 ? 6 : 5), shr = cols > 64 ? 7 : (cols > 32 ? 6 : 5);
      const int rk = tid & ((1 << shr) - 1), rrow = tid >> shr, rstep = nt >> shr;
      const bool r_on = rk < cols, r_in = k0 + rk < n_cols;
      const double* gR = Rp + (size_t)(R0 + rrow) * ld_r + k0 + rk;
      const int dR = rrow * kStageLdK + rk;
      const int lk = LT ? (tid & ((1 << shl) - 1)) : (tid & (kStageRC - 1)), lrow = LT ? (tid >> shl) : (tid >> kStageRCLog);
      const int lstep = LT ? (nt >> shl) : (nt >> kStageRCLog);
      const bool l_on = LT ? lk < rows : true, l_in = LT ? (i0 + lk < n_rows) : true;
      const double* gL = LT ? Lp + (size_t)(R0 + lrow) * ld_l + i0 + lk : Lp + (size_t)(i0 + lrow) * ld_l + R0 + lk;
      const int dL = LT ? lrow * kStageLdK + lk : lrow * kStageLdR + lk;
      // WIDE plan: the same with column PAIRS (K-major panels) / reduction-index pairs (row-major L panel)
      const int wrk = (tid & ((1 << (shr - 1)) - 1)) * 2, wrrow = tid >> (shr - 1), wrstep = nt >> (shr - 1);
      const int wlk = LT ? (tid & ((1 << (shl - 1)) - 1)) * 2 : (tid & (kStageRC / 2 - 1)) * 2;
      const int wlrow = LT ? (tid >> (shl - 1)) : (tid >> (kStageRCLog - 1));
      const int wlstep = LT ? (nt >> (shl - 1)) : (nt >> (kStageRCLog - 1));
      auto issue_wide = [&](int ch) {
        double* sL = stage + (size_t)(ch % kStageDepth) * (2 * kStagePanel);
        double* sR = sL + kStagePanel;
        const int r = R0 + ch * kStageRC;
        if (wrk < cols) {
          const int nv = n_cols - (k0 + wrk);                  // valid columns of the pair: >= 2, 1 or <= 0
          const double* src = Rp + (size_t)(r + wrrow) * ld_r + k0 + wrk;
          double* dst = sR + wrrow * kStageLdK + wrk;
          const int lim = n_red - r;
          for (int rr = wrrow; rr < kStageRC; rr += wrstep, src += (size_t)wrstep * ld_r, dst += wrstep * kStageLdK)
            cp_async_pair(dst, src, rr < lim ? nv : 0);
        }
        if (LT) {
          if (wlk < rows) {
            const int nv = n_rows - (i0 + wlk);
            const double* src = Lp + (size_t)(r + wlrow) * ld_l + i0 + wlk;
            double* dst = sL + wlrow * kStageLdK + wlk;
            const int lim = n_red - r;
            for (int rr = wlrow; rr < kStageRC; rr += wlstep, src += (size_t)wlstep * ld_l, dst += wlstep * kStageLdK)
              cp_async_pair(dst, src, rr < lim ? nv : 0);
          }
        } else {
          const int nv = n_red - (r + wlk);                    // valid reduction indices of the pair
          const double* src = Lp + (size_t)(i0 + wlrow) * ld_l + r + wlk;
          double* dst = sL + wlrow * kStageLdR + wlk;
          const int lim = n_rows - i0;
          for (int i = wlrow; i < rows; i += wlstep, src += (size_t)wlstep * ld_l, dst += wlstep * kStageLdR)
            cp_async_pair(dst, src, i < lim ? nv : 0);
        }
        cp_async_commit();
      };
      auto issue_narrow = [&](int ch) {
        double* sL = stage + (size_t)(ch % kStageDepth) * (2 * kStagePanel);
        double* sR = sL + kStagePanel;
        const int r = R0 + ch * kStageRC;
        if (r_on) {
          const double* src = gR + (size_t)ch * kStageRC * ld_r;
          double* dst = sR + dR;
          const int lim = r_in ? n_red - r : 0;            // rows rr < lim are inside the matrix
          for (int rr = rrow; rr < kStageRC; rr += rstep, src += (size_t)rstep * ld_r, dst += rstep * kStageLdK) {
            if (rr < lim) cp_async_f64(dst, src); else *dst = 0.0;
          }
        }
        if (LT) {       // L(i, r) = Lp[r * ld_l + i]: K-major panel sL[rr][i]
          if (l_on) {
            const double* src = gL + (size_t)ch * kStageRC * ld_l;
            double* dst = sL + dL;
            const int lim = l_in ? n_red - r : 0;
            for (int rr = lrow; rr < kStageRC; rr += lstep, src += (size_t)lstep * ld_l, dst += lstep * kStageLdK) {
              if (rr < lim) cp_async_f64(dst, src); else *dst = 0.0;
            }
          }
        } else {        // row-major panel sL[i][rr]
          const double* src = gL + ch * kStageRC;
          double* dst = sL + dL;
          const bool red_in = r + lk < n_red;
          const int lim = red_in ? n_rows - i0 : 0;        // rows i < lim are inside the matrix
          for (int i = lrow; i < rows; i += lstep, src += (size_t)lstep * ld_l, dst += lstep * kStageLdR) {
            if (i < lim) cp_async_f64(dst, src); else *dst = 0.0;
          }
        }
        cp_async_commit();
      };
      auto issue = [&](int ch) { if (WIDE) issue_wide(ch); else issue_narrow(ch); };
      __syncthreads();                        // the ring is free (previous block / previous user of `stage`)
      if (nchunk > 0) issue(0);
      if (nchunk > 1) issue(1);
      for (int ch = 0; ch < nchunk; ++ch) {
        if (ch + 1 < nchunk) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncthreads();                      // chunk ch is visible; everyone is done with chunk ch - 1
        if (ch + 2 < nchunk) issue(ch + 2);   // into the slot chunk ch - 1 occupied
        const double* sL = stage + (size_t)(ch % kStageDepth) * (2 * kStagePanel);
        const double* sR = sL + kStagePanel;
        const int r = R0 + ch * kStageRC;
#pragma unroll
        for (int m = 0; m < kStageMaxT; ++m) {
          if (tr1[m] <= tr0[m]) continue;
          const double* pa = LT ? sL + q * kStageLdK + tib[m] + g : sL + (tib[m] + g) * kStageLdR + q;
          const double* pb = sR + q * kStageLdK + tkb[m] + g;
#pragma unroll
          for (int kk = 0; kk < kStageRC; kk += 4) {
            if (r + kk < tr0[m] || r + kk >= tr1[m]) continue;
            double af[2], bf[2];
#pragma unroll
            for (int a = 0; a < 2; ++a) af[a] = LT ? pa[kk * kStageLdK + 8 * a] : pa[8 * a * kStageLdR + kk];
#pragma unroll
            for (int b = 0; b < 2; ++b) bf[b] = pb[kk * kStageLdK + 8 * b];
            if (SCALE) {
              const int rk2 = r + kk + q;
              const double sc = scale[rk2 < n_red ? rk2 : n_red - 1];
              bf[0] *= sc; bf[1] *= sc;
            }
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
              for (int b = 0; b < 2; ++b) dmma884(c[m][a][b][0], c[m][a][b][1], af[a], bf[b]);
          }
        }
      }
#pragma unroll
      for (int m = 0; m < kStageMaxT; ++m) {
        if (wid + m * nw >= ntile) continue;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            const int i = i0 + tib[m] + 8 * a + g;
            const int j = k0 + tkb[m] + 8 * b + 2 * q;
            if (i < n_rows) {
              if (j < n_cols) store(i, j, c[m][a][b][0]);
              if (j + 1 < n_cols) store(i, j + 1, c[m][a][b][1]);
            }
          }
      }
    }
  __syncthreads();   // `stage` may be reused by the caller
}
#endif

// Contraction dispatch: tensor cores on the device, the register-tiled FMA loop in the host emulation.
template <bool LT, bool SCALE, class RangeF, class StoreF>
STP_HD void contract_nn(const SoloTeam& tm, int n_rows, int n_cols, int, const double* Lp, int ld_l, const double* Rp,
                        int ld_r, const double* scale, RangeF range, StoreF store) {
  contract_p<4, 2, LT, SCALE>(tm, n_rows, n_cols, Lp, ld_l, Rp, ld_r, scale, range, store);
}
#if defined(__CUDACC__)
template <bool LT, bool SCALE, class RangeF, class StoreF>
__device__ __forceinline__ void contract_nn(const CtaTeam& tm, int n_rows, int n_cols, int n_red, const double* Lp, int ld_l,
                                            const double* Rp, int ld_r, const double* scale, RangeF range,
                                            StoreF store) {
#if defined(STP_NO_DMMA)
  contract_p<4, 2, LT, SCALE>(tm, n_rows, n_cols, Lp, ld_l, Rp, ld_r, scale, range, store);
#else
  contract_mma<LT, SCALE>(tm, n_rows, n_cols, n_red, Lp, ld_l, Rp, ld_r, scale, range, store);
#endif
}
#endif

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
    const double inv = rcp_pos(dk);
    const double isd = rsqrt_pos(dk);
    const double lkk = dk * isd;
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
