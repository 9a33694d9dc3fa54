// Exact-GP pieces of the BO hot path, one campaign per team (CTA):
//   exact_mll_fg      K1+K3+K4+K5  loss = -(log N(y|c, K+s2 I) + log-priors)/n and its analytic
//                                   gradient (what botorch's closure hands to L-BFGS-B at
//                                   optim.py:31-32; formulas SURVEY.md App. A.3)
//   exact_factorize   K1+K3+K4      L^-1 and alpha = (K+s2 I)^-1 (y-c) at the fitted theta
//   exact_acquire     K2+K10+K12+K13 fused cross-covariance, predictive mean/variance
//                                   (models.py:387-395), LogEI (runners.py:88-90) and the
//                                   first-index argmax over the remaining pool (runners.py:95)
// theta layout (botorch parameter order): [noise, constant, outputscale, lengthscale_0..d-1].
#pragma once
#include "linalg.cuh"
#include "logei.cuh"
#include "team.cuh"

namespace stp {

constexpr int kMaxD = 8;          // input dimensions supported by the unrolled ARD loops
constexpr double kLog2Pi = 1.83787706640934548356;

struct ExactPrior {               // LogNormal(loc, scale) on noise, outputscale, lengthscale
  double noise_loc, noise_scale, os_loc, os_scale, ls_loc, ls_scale;
};

STP_HD_NOINLINE double lognormal_logpdf(double x, double loc, double scale) {
  const double lx = log(x);
  const double z = lx - loc;
  return -lx - log(scale) - 0.5 * kLog2Pi - z * z / (2.0 * scale * scale);
}
STP_HD_NOINLINE double lognormal_dlogpdf(double x, double loc, double scale) {
  return -(1.0 + (log(x) - loc) / (scale * scale)) / x;
}

// Shared-memory carve-up of one exact-GP campaign with room for `cap` training points.
struct ExactSmem {
  double* A;      // (cap+1) x ld : lower = working matrix, strict upper = noise-free K (fit only)
  double* Xt;     // d x cap      : training inputs, transposed (conflict-free across points)
  double* y;      // cap
  double* c0;     // cap+5
  double* c1;     // cap+5
  double* c2;     // cap+9  (pair sweep: next pivot columns)
  double* c3;     // cap+9
  double* piv;    // cap+5 (pivots; reused as tmp by the triangular inverse)
  double* alpha;  // cap
  double* tile;   // acquisition: cap x TC cross-covariances + TC x (kMaxD) staged candidates;
                  // fit: the three panels of the blocked sweep (never live together)
  double* part;   // acquisition: nwarps x TC x 2 partial sums (ALIASES `red`: never live together)
  double* small;  // 64 doubles: theta (<=11), gradient (<=11), misc
  double* red;    // kRedDoubles reduction scratch
  int ld, cap;
  int kern;          // kKernRBF / kKernMatern52 (set by the caller after the carve-up; 0 = the reference's RBF)
  int acq_kind;      // kAcqLogEI / kAcqUCB / kAcqEI
  double acq_beta;   // UCB's beta
};

STP_HD size_t exact_tile_doubles(int cap, int tc) {
  const size_t acq = (size_t)((cap + 15) & ~15) * tc + (size_t)tc * kMaxD, swp = sweep_blocked_scratch_doubles(cap + 1);
  return acq > swp ? acq : swp;
}
STP_HD size_t exact_smem_doubles(int cap, int d, int tc, int nwarps) {
  const int ld = odd_ld(cap + 1);
  return (size_t)(cap + 1) * ld + (size_t)d * cap + cap + 5 * (size_t)(cap + 9) + cap +
         exact_tile_doubles(cap, tc) + 64 + 16 * kMaxWarps;  // nwarps*tc*2 <= 16*kMaxWarps
}

STP_HD ExactSmem exact_smem_carve(double* base, int cap, int d, int tc, int nwarps) {
  ExactSmem s;
  s.kern = 0; s.acq_kind = 0; s.acq_beta = 0.0;
  s.cap = cap;
  s.ld = odd_ld(cap + 1);
  double* p = base;
  s.A = p;      p += (size_t)(cap + 1) * s.ld;
  s.Xt = p;     p += (size_t)d * cap;
  s.y = p;      p += cap;
  s.c0 = p;     p += cap + 9;   // (cap + 1) rounded up to a multiple of the register-block edge (<= 8)
  s.c1 = p;     p += cap + 9;
  s.c2 = p;     p += cap + 9;
  s.c3 = p;     p += cap + 9;
  s.piv = p;    p += cap + 9;
  s.alpha = p;  p += cap;
  s.tile = p;   p += exact_tile_doubles(cap, tc);
  s.small = p;  p += 64;
  s.red = p;
  s.part = p;   // block reductions and the acquisition partials are never in flight together
  return s;
}

// Large capacities (benchmark_predictions.py:19-75 fits 80 % of a dataset, n up to 480): the
// working square no longer fits in shared memory next to the vectors, so it lives in a per-campaign GLOBAL workspace
// (L2-resident: 1.9 MB at n = 480) while the vectors, the training inputs, the sweep panels / pool tile and the
// optimiser state stay in shared memory.  Every routine addresses the square through s.A / s.ld, so the arithmetic
// (and its order) is the one of the shared-memory path; barriers order the global accesses of a CTA.
STP_HD size_t exact_large_work_doubles(int cap) { return (size_t)(cap + 1) * odd_ld(cap + 1); }
STP_HD size_t exact_large_smem_doubles(int cap, int d, int tc) {
  return (size_t)d * cap + cap + 5 * (size_t)(cap + 9) + cap + exact_tile_doubles(cap, tc) + 64 + 16 * kMaxWarps;
}
STP_HD ExactSmem exact_smem_carve_large(double* base, double* square, int cap, int d, int tc) {
  ExactSmem s;
  s.kern = 0; s.acq_kind = 0; s.acq_beta = 0.0;
  s.cap = cap;
  s.ld = odd_ld(cap + 1);
  s.A = square;
  double* p = base;
  s.Xt = p;     p += (size_t)d * cap;
  s.y = p;      p += cap;
  s.c0 = p;     p += cap + 9;
  s.c1 = p;     p += cap + 9;
  s.c2 = p;     p += cap + 9;
  s.c3 = p;     p += cap + 9;
  s.piv = p;    p += cap + 9;
  s.alpha = p;  p += cap;
  s.tile = p;   p += exact_tile_doubles(cap, tc);
  s.small = p;  p += 64;
  s.red = p;
  s.part = p;
  return s;
}

// Copy one campaign's training set into shared memory (X row-major [n,d] -> Xt [d][cap]).
template <class Team>
STP_HD void exact_load(const Team& tm, const ExactSmem& s, int n, int d, const double* X, const double* y) {
  for (int idx = tm.rank(); idx < n * d; idx += tm.size()) {
    const int i = idx / d, k = idx - i * d;
    s.Xt[k * s.cap + i] = X[idx];
  }
  for (int i = tm.rank(); i < n; i += tm.size()) s.y[i] = y[i];
  tm.sync();
}

// Availability byte of pool point i.  Read through the L2 on the device: in the persistent campaign kernel the mask
// row may have been written by another SM earlier in the same launch (this SM's L1 could hold a stale line).
STP_HD uint8_t mask_byte(const uint8_t* mask, int i) {
#if defined(__CUDA_ARCH__)
  return __ldcg(mask + i);
#else
  return mask[i];
#endif
}

// Stationary ARD kernels as functions of r2 = ||(x - x') / l||^2.  RBF is what the reference uses (models.py:342-358);
// Matern-5/2 (gpytorch MaternKernel(nu = 2.5): os (1 + sqrt5 r + 5/3 r^2) exp(-sqrt5 r), r = sqrt(max(r2, 1e-30)))
// is the alternative north_star names (row f4).
enum { kKernRBF = 0, kKernMatern52 = 1 };
STP_HD_NOINLINE double matern52_of_r2(double r2, double os) {
  const double r = sqrt(r2 > 1e-30 ? r2 : 1e-30), sr = 2.23606797749978969641 * r;
  return os * (1.0 + sr + (5.0 / 3.0) * r * r) * exp(-sr);
}
// d k / d l_k = kernel_dfactor * k * (x_k - x'_k)^2 / l_k^3:  1 for RBF, (5/3)(1 + sqrt5 r) / (1 + sqrt5 r + 5/3 r^2) for Matern-5/2
STP_HD_NOINLINE double matern52_dfactor(double r2) {
  const double r = sqrt(r2 > 1e-30 ? r2 : 1e-30), sr = 2.23606797749978969641 * r;
  return (5.0 / 3.0) * (1.0 + sr) / (1.0 + sr + (5.0 / 3.0) * r * r);
}
STP_HD double kernel_of_r2(int kern, double r2, double os) {
  return kern == kKernMatern52 ? matern52_of_r2(r2, os) : os * exp(-0.5 * r2);
}

// noise-free kernel value between training points i and j (direct differences, App. A.1)
STP_HD double ard_rbf_tt(const double* Xt, int cap, int d, const double* invl, int i, int j, double os, int kern = 0) {
  double r2 = 0.0;
#pragma unroll
  for (int k = 0; k < kMaxD; ++k)
    if (k < d) {
      const double t = (Xt[k * cap + i] - Xt[k * cap + j]) * invl[k];
      r2 += t * t;
    }
  return kernel_of_r2(kern, r2, os);
}

// Calls f(i, j) once for every pair of the strict lower triangle 0 <= j < i < n.  Row i of a triangle has i
// entries, so a warp that walks rows leaves most lanes idle on the short ones (60 % lane utilisation at n = 60);
// here row F is folded with row n - F into one run of n entries, walked by the lanes of one warp (93 - 100 %).
template <class Team, class Fn>
STP_HD void for_each_lower_pair(const Team& tm, int n, Fn f) {
  const int W = tm.warp_width();
  for (int F = 1 + tm.warp(); 2 * F <= n; F += tm.nwarps()) {
    const int len = (2 * F == n) ? F : n;          // the middle row of an even n stands alone
    for (int c = tm.lane(); c < len; c += W) {
      if (c < F) f(F, c);
      else f(n - F, c - F);
    }
  }
}

// Bordered sweep of the closure: blocked tensor-core form on the device, pair sweep in the host emulation.
template <class Team>
STP_HD int exact_sweep(const Team& tm, const ExactSmem& s, int n) {
  return sweep_bordered4(tm, s.A, s.ld, n + 1, n, s.c0, s.c1, s.c2, s.c3, s.piv);
}
#if defined(__CUDACC__) && !defined(STP_NO_BLOCKED_SWEEP)
__device__ __forceinline__ int exact_sweep(const CtaTeam& tm, const ExactSmem& s, int n) {
  return sweep_bordered_blocked(tm, s.A, s.ld, n + 1, n, s.tile, s.piv);
}
#endif

// ---------------------------------------------------------------------------------------
// One L-BFGS-B closure call.  Every thread returns the same status and loss; the gradient
// (d+3 doubles) is left in `g` (shared memory), valid after the call for every thread.
// status: 0 ok, 1 matrix not positive definite, 2 non-finite parameter / objective.
// Work: n^3/2 MAC (sweep) + n^2 (3d+8)/2 flop (kernel build + gradient contraction).
// ---------------------------------------------------------------------------------------
template <class Team>
STP_HD int exact_mll_fg(const Team& tm, const ExactSmem& s, int n, int d, const double* theta,
                        const ExactPrior& pr, double* f_out, double* g) {
  const int ld = s.ld, cap = s.cap, W = tm.warp_width();
  const double noise = theta[0], cst = theta[1], os = theta[2];
  double* invl = s.small + 32;  // d doubles
  tm.sync();
  STP_OWNED(tm, k, d) invl[k] = 1.0 / theta[3 + k];
  int bad = !(noise > 0.0) || !(os > 0.0) || !(cst == cst);
  STP_ROLLED for (int k = 0; k < d; ++k) bad |= !(theta[3 + k] > 0.0);
  if (bad) { *f_out = NAN; return 2; }
  tm.sync();
  STP_PHASE(-1); STP_COUNT(16);
  // K1: lower = K + noise I (to be swept), strict upper = K (kept for the gradient), border = y - c
  // The sweep works in place, so a retry with more jitter (psd_safe_cholesky) rebuilds the matrix; it never happens
  // on a healthy fit and costs one extra Gram build when it does.
  int fail = 1;
  STP_ROLLED for (int attempt = 0; attempt < kPsdSafeTries && fail; ++attempt) {
    const double diag = os + noise + psd_safe_jitter(attempt);
    for_each_lower_pair(tm, n, [&](int i, int j) {
      const double kij = ard_rbf_tt(s.Xt, cap, d, invl, i, j, os, s.kern);
      s.A[i * ld + j] = kij;
      s.A[j * ld + i] = kij;
    });
    for (int j = tm.rank(); j <= n; j += tm.size()) {
      if (j < n) s.A[j * ld + j] = diag;
      s.A[n * ld + j] = (j < n) ? (s.y[j] - cst) : 0.0;
    }
    tm.sync();
    STP_PHASE(1);
    // K3+K4: bordered sweep -> -Kinv, alpha, -quad, pivots
    fail = exact_sweep(tm, s, n);
    if (fail) tm.sync();      // every thread has left the sweep before the matrix is rebuilt
  }
  if (fail) { *f_out = NAN; return 1; }
  STP_PHASE(-1);
  const double* arow = s.A + n * ld;  // alpha
  // K5: reductions
  double acc[kMaxD + 5];
#pragma unroll
  for (int q = 0; q < kMaxD + 5; ++q) acc[q] = 0.0;
  // acc[0] = sum log piv, acc[1] = sum W_ii, acc[2] = sum alpha, acc[3] = sum_{i>j} W_ij K_ij,
  // acc[4+k] = sum_{i>j} W_ij K_ij (x_ik - x_jk)^2, acc[4+d] = sum of the log-priors (one parameter per thread)
  STP_OWNED(tm, k, d + 2) {   // one call site: the inlined log() sequences are what made this part large
    const double x = (k == 0) ? noise : ((k == 1) ? os : theta[1 + k]);
    const double loc = (k == 0) ? pr.noise_loc : ((k == 1) ? pr.os_loc : pr.ls_loc);
    const double sc = (k == 0) ? pr.noise_scale : ((k == 1) ? pr.os_scale : pr.ls_scale);
    acc[4 + d] += lognormal_logpdf(x, loc, sc);
  }
  STP_ROLLED for (int i = tm.rank(); i < n; i += tm.size()) {
    const double ai = arow[i];
    acc[0] += log(s.piv[i]);
    acc[1] += ai * ai + s.A[i * ld + i];
    acc[2] += ai;
  }
  const int kern = s.kern;
  for_each_lower_pair(tm, n, [&](int i, int j) {
    const double t = (arow[i] * arow[j] + s.A[i * ld + j]) * s.A[j * ld + i];
    acc[3] += t;
    double dxv[kMaxD];
    double r2 = 0.0;
#pragma unroll
    for (int k = 0; k < kMaxD; ++k) {
      dxv[k] = (k < d) ? s.Xt[k * cap + i] - s.Xt[k * cap + j] : 0.0;
      if (kern != kKernRBF && k < d) { const double u = dxv[k] * invl[k]; r2 += u * u; }
    }
    const double tl = (kern == kKernRBF) ? t : t * matern52_dfactor(r2);
#pragma unroll
    for (int k = 0; k < kMaxD; ++k)
      if (k < d) acc[4 + k] += tl * dxv[k] * dxv[k];
  });
  STP_PHASE(5);
  tm.sum_vec(acc, 5 + d);
  const double quad = -s.A[n * ld + n];
  const double logN = -0.5 * (quad + acc[0] + n * kLog2Pi);
  const double total = logN + acc[4 + d];
  const double scale = -1.0 / n;
  const double f = total * scale;
  tm.sync();
  // parameter p of (noise, constant, outputscale, lengthscale_1..d) by thread p; one division / one prior call site:
  //   g0 = scale (tr W / 2 + dlogp(noise)), g1 = scale sum alpha, g2 = scale (S_K / os + tr W / 2 + dlogp(os)),
  //   g_{3+k} = scale (S_k / l_k^3 + dlogp(l_k))
  STP_OWNED(tm, p, d + 3) {
    if (p == 1) g[1] = scale * acc[2];
    else {
      const double x = (p == 0) ? noise : ((p == 2) ? os : theta[p]);
      const double loc = (p == 0) ? pr.noise_loc : ((p == 2) ? pr.os_loc : pr.ls_loc);
      const double sc = (p == 0) ? pr.noise_scale : ((p == 2) ? pr.os_scale : pr.ls_scale);
      double data = 0.5 * acc[1];
      if (p >= 2) {
        const double q = acc[1 + p] / ((p == 2) ? os : x * x * x);
        data = (p == 2) ? q + data : q;
      }
      g[p] = scale * (data + lognormal_dlogpdf(x, loc, sc));
    }
  }
  tm.sync();
  STP_PHASE(6);
  *f_out = f;
  int st = !(f == f) || !(fabs(f) <= 1.79e308);
  STP_ROLLED for (int k = 0; k < d + 3; ++k) st |= !(g[k] == g[k]);
  return st ? 2 : 0;
}

// ---------------------------------------------------------------------------------------
// Factorise at the fitted theta: A(lower) <- L^-1 with K + noise I = L L^T, alpha <- K^-1 (y-c).
// Returns 0 or 1 (not PD).
// ---------------------------------------------------------------------------------------
template <class Team>
STP_HD int exact_factorize(const Team& tm, const ExactSmem& s, int n, int d, const double* theta) {
  const int ld = s.ld, cap = s.cap, W = tm.warp_width();
  const double noise = theta[0], cst = theta[1], os = theta[2];
  double* invl = s.small + 32;
  tm.sync();
  STP_OWNED(tm, k, d) invl[k] = 1.0 / theta[3 + k];
  tm.sync();
  int fail = 1;
  STP_ROLLED for (int attempt = 0; attempt < kPsdSafeTries && fail; ++attempt) {   // psd_safe_cholesky
    const double diag = os + noise + psd_safe_jitter(attempt);
    for (int i = tm.warp(); i < n; i += tm.nwarps()) {
      double* row = s.A + i * ld;
      for (int j = tm.lane(); j <= i; j += W)
        row[j] = (j == i) ? diag : ard_rbf_tt(s.Xt, cap, d, invl, i, j, os, s.kern);
    }
    fail = cholesky_inverse_fused(tm, s.A, ld, n, s.c1, s.c2, s.piv, (double*)nullptr, 0);
    if (fail) tm.sync();
  }
  if (fail) return 1;
  // z = Linv (y - c)  -> c0 ;  alpha = Linv^T z
  for (int i = tm.rank(); i < n; i += tm.size()) {
    const double* row = s.A + i * ld;
    double z = 0.0;
    for (int j = 0; j <= i; ++j) z += row[j] * (s.y[j] - cst);
    s.c0[i] = z;
  }
  tm.sync();
  for (int j = tm.rank(); j < n; j += tm.size()) {
    double a = 0.0;
    for (int i = j; i < n; ++i) a += s.A[i * ld + j] * s.c0[i];
    s.alpha[j] = a;
  }
  tm.sync();
  return 0;
}

// Partial sums of one candidate tile (TC candidates, cross-covariances in s.tile[j * TC + c]):
//   part[(p * TC + c) * 2 + 0], p < n_ssq : pieces of || Linv k_c ||^2      part[(p * TC + c) * 2 + 1], p < n_mu : of alpha . k_c
// Generic form: warp w takes the rows i = w, w + nwa, ...; a lane is a candidate.
template <class Team>
STP_HD void acq_tile_partials(const Team& tm, const ExactSmem& s, int n, int& n_ssq, int& n_mu) {
  const int ld = s.ld, TC = tm.warp_width(), nw = tm.nwarps();
  const int nwa = nw < 8 ? nw : 8;   // at most 8 warps: the partials live in the 512-double reduction scratch
  n_ssq = nwa; n_mu = nwa;
  if (tm.warp() < nwa) {
    const int c = tm.lane();
    double ssq = 0.0, mu = 0.0;
    for (int i = tm.warp(); i < n; i += nwa) {
      const double* row = s.A + i * ld;
      double v0 = 0.0, v1 = 0.0;
      int j = 0;
      for (; j + 1 <= i; j += 2) {
        v0 += row[j] * s.tile[j * TC + c];
        v1 += row[j + 1] * s.tile[(j + 1) * TC + c];
      }
      if (j <= i) v0 += row[j] * s.tile[j * TC + c];
      const double v = v0 + v1;
      ssq += v * v;
      mu += s.alpha[i] * s.tile[i * TC + c];
    }
    s.part[(tm.warp() * TC + c) * 2 + 0] = ssq;
    s.part[(tm.warp() * TC + c) * 2 + 1] = mu;
  }
}
#if defined(__CUDACC__) && !defined(STP_NO_DMMA)
// Tensor-core form: V = Linv K_tile in 16 x 16 DMMA tiles (heaviest row tiles first), the squares summed per row
// tile; rows n .. 16 ceil(n / 16) - 1 of the tile are zero (exact_acquire fills them).
__device__ __forceinline__ void acq_tile_partials(const CtaTeam& tm, const ExactSmem& s, int n, int& n_ssq, int& n_mu) {
  const int RT = (n + 15) >> 4;
  if (RT > 8) { acq_tile_partials<CtaTeam>(tm, s, n, n_ssq, n_mu); return; }
  constexpr int TC = 32;
  const int ld = s.ld, nw = tm.nwarps(), warp = tm.warp(), lane = tm.lane();
  const int g = lane >> 2, q = lane & 3;
  const int nwa = nw < 8 ? nw : 8;
  n_ssq = RT; n_mu = nwa;
  if (warp < nwa) {
    double mu = 0.0;
    for (int i = warp; i < n; i += nwa) mu += s.alpha[i] * s.tile[i * TC + lane];
    s.part[(warp * TC + lane) * 2 + 1] = mu;
  }
  for (int t = warp; t < 2 * RT; t += nw) {
    const int I = RT - 1 - (t >> 1), jb = (t & 1) << 4, ib = I << 4;
    double c[2][2][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) { c[a][b][0] = 0.0; c[a][b][1] = 0.0; }
    const int i0 = ib + g, i1 = ib + 8 + g;
    const double* r0 = s.A + (size_t)((i0 < n) ? i0 : n - 1) * ld;
    const double* r1 = s.A + (size_t)((i1 < n) ? i1 : n - 1) * ld;
    const double* tb = s.tile + jb + g;
    for (int k0 = 0; k0 < ib; k0 += 4) {         // strictly below the diagonal block: no masks
      const int kk = k0 + q;
      const double a0 = r0[kk], a1 = r1[kk];
      const double b0 = tb[kk * TC], b1 = tb[kk * TC + 8];
      dmma884(c[0][0][0], c[0][0][1], a0, b0);
      dmma884(c[0][1][0], c[0][1][1], a0, b1);
      dmma884(c[1][0][0], c[1][0][1], a1, b0);
      dmma884(c[1][1][0], c[1][1][1], a1, b1);
    }
#pragma unroll
    for (int k0 = 0; k0 < 16; k0 += 4) {          // the diagonal block: Linv is lower triangular
      const int kk = ib + k0 + q;
      const double a0 = (i0 < n && kk <= i0) ? r0[kk] : 0.0, a1 = (i1 < n && kk <= i1) ? r1[kk] : 0.0;
      const double b0 = tb[kk * TC], b1 = tb[kk * TC + 8];
      dmma884(c[0][0][0], c[0][0][1], a0, b0);
      dmma884(c[0][1][0], c[0][1][1], a0, b1);
      dmma884(c[1][0][0], c[1][0][1], a1, b0);
      dmma884(c[1][1][0], c[1][1][1], a1, b1);
    }
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        double sq = 0.0;
        if (i0 < n) sq += c[0][b][e] * c[0][b][e];
        if (i1 < n) sq += c[1][b][e] * c[1][b][e];
        sq += __shfl_xor_sync(0xffffffffu, sq, 4);
        sq += __shfl_xor_sync(0xffffffffu, sq, 8);
        sq += __shfl_xor_sync(0xffffffffu, sq, 16);
        if (g == 0) s.part[(I * TC + jb + 8 * b + 2 * q + e) * 2 + 0] = sq;
      }
  }
}
#endif

// ---------------------------------------------------------------------------------------
// Fused pool pass.  For every pool point c (N of them, row-major [N,d] in global memory):
//   k_j = os exp(-1/2 ||(x_c - x_j)/l||^2),  mean = cst + k.alpha,
//   var = os - ||Linv k||^2 + noise          (observation noise included, App. D-3)
//   acq = LogEI(mean, var, best_f)
// and the argmax over the points whose mask byte is non-zero.  Tiles of TC = warp-width
// candidates are staged in shared memory; each warp owns an interleaved quarter of the rows
// of Linv, so Linv[i][j] is a broadcast read and the tile column is conflict-free.
// Optional per-point outputs (may be null): mean_out, var_out, acq_out [N].
// best_f may be NAN when only the predictive is wanted (acq_out then holds NaN).
// ---------------------------------------------------------------------------------------
template <class Team>
STP_HD void exact_acquire(const Team& tm, const ExactSmem& s, int n, int d, const double* theta,
                          const double* pool, int N, const uint8_t* mask, double best_f,
                          int* best_idx, double* best_acq, double* mean_out, double* var_out,
                          double* acq_out) {
  const int cap = s.cap, TC = tm.warp_width();
  const double noise = theta[0], cst = theta[1], os = theta[2];
  const double* invl = s.small + 32;
  const int n16 = (n + 15) & ~15;                             // rows n .. n16 - 1 of the tile are zero-filled
  double* xc = s.tile + (size_t)((cap + 15) & ~15) * TC;  // [kMaxD][TC]
  double run_v = -INFINITY;
  int run_i = -1;
  for (int base = 0; base < N; base += TC) {
    const int cnt = (N - base < TC) ? (N - base) : TC;
    tm.sync();
    for (int idx = tm.rank(); idx < cnt * d; idx += tm.size()) {
      const int c = idx / d, k = idx - c * d;
      xc[k * TC + c] = pool[(size_t)base * d + idx];
    }
    tm.sync();
    for (int idx = tm.rank(); idx < n16 * TC; idx += tm.size()) {
      const int j = idx / TC, c = idx - j * TC;
      double v = 0.0;
      if (c < cnt && j < n) {
        double r2 = 0.0;
#pragma unroll
        for (int k = 0; k < kMaxD; ++k)
          if (k < d) {
            const double t = (xc[k * TC + c] - s.Xt[k * cap + j]) * invl[k];
            r2 += t * t;
          }
        v = kernel_of_r2(s.kern, r2, os);
      }
      s.tile[j * TC + c] = v;
    }
    tm.sync();
    int n_ssq, n_mu;
    acq_tile_partials(tm, s, n, n_ssq, n_mu);
    tm.sync();
    for (int c = tm.rank(); c < cnt; c += tm.size()) {
      double ssq = 0.0, mu = 0.0;
      for (int w = 0; w < n_ssq; ++w) ssq += s.part[(w * TC + c) * 2 + 0];
      for (int w = 0; w < n_mu; ++w) mu += s.part[(w * TC + c) * 2 + 1];
      const double mean = cst + mu;
      const double var = os - ssq + noise;
      const double a = acq_value(s.acq_kind, s.acq_beta, mean, var, best_f);
      const int gi = base + c;
      if (mean_out) mean_out[gi] = mean;
      if (var_out) var_out[gi] = var;
      if (acq_out) acq_out[gi] = a;
      if ((!mask || mask_byte(mask, gi)) && acq_better(a, gi, run_v, run_i)) { run_v = a; run_i = gi; }
    }
  }
  tm.argmax(run_v, run_i);
  *best_idx = run_i;
  *best_acq = run_v;
}

}  // namespace stp
