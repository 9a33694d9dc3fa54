// Whitened SVGP (VarGP / VarTGP / VarLGP) pieces of the BO hot path, one campaign per team:
//   var_epoch     K7+K8+K9  one pass of the loop body optim.py:150-170: ELBO forward, the
//                           hand-derived backward, the natural-gradient
//                           step on (theta1, Theta2) and the Adam step on (Z, c, l, os, noise[, rho])
//   var_prepare   K3+K4     Linv_K, G = L_P^-1 Linv_K, w = Linv_K^T mu for the pool pass
//   var_acquire   K2+K10-13 fused cross-covariance, q(f) mean / variance at every pool point,
//                           analytic LogEI (VarGP) or the S-sample Monte-Carlo mean of LogEI
//                           (VarTGP / VarLGP, models.py:95-111 / 280-297 + runners.py:88-93)
//                           and the first-index argmax.
//
// Layout.  All n x n work matrices are ROW-PER-DATA-POINT ("t" = transposed w.r.t. the maths of
// App. A.4): At[c][k] = A[k][c] etc., so that every O(n^3) contraction reads one operand as a
// warp-wide broadcast and the other as consecutive doubles.  They live in a per-campaign
// workspace in global memory (L1/L2-resident: 7 squares of (cap+1) x ld doubles); vectors and the
// transposed point sets live in shared memory.
#pragma once
#include "exact.cuh"
#include "linalg.cuh"
#include "logei.cuh"
#include "team.cuh"

namespace stp {

enum { kLikGaussian = 0, kLikStudentT = 1, kLikLaplace = 2 };
constexpr int kGH = 20;
constexpr double kVarJitter = 1e-6;  // gpytorch's variational Cholesky jitter in fp64

struct VarHyperPrior {   // LogNormal on lengthscale / outputscale / noise, Gamma(conc, rate) on nu
  double ls_loc, ls_scale, os_loc, os_scale, noise_loc, noise_scale, df_conc, df_rate;
};

// digamma for x > 0: upward recurrence to x >= 10, then the asymptotic series (|err| < 1e-15).
STP_HD double digamma_pos(double x) {
  double r = 0.0;
  while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
  const double f = 1.0 / (x * x);
  const double t = f * (-1.0 / 12.0 + f * (1.0 / 120.0 + f * (-1.0 / 252.0 + f * (1.0 / 240.0 + f * (-1.0 / 132.0 +
                   f * (691.0 / 32760.0 + f * (-1.0 / 12.0)))))));
  return r + log(x) - 0.5 / x + t;
}

STP_HD double softplus(double x) { return x > 30.0 ? x : log1p(exp(x)); }
STP_HD double sigmoid(double x) { return 1.0 / (1.0 + exp(-x)); }

// ---- Philox4x32-10 counter-based generator -> standard normals (Box-Muller, fp64) ----------
STP_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                          uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Two independent N(0,1) draws for (seed, stream, counter): stream = campaign, counter = (candidate, pair).
STP_HD void philox_normal2(uint64_t seed, uint64_t stream, uint64_t counter, double& z0, double& z1) {
  uint32_t o[4];
  philox4x32_10((uint32_t)counter, (uint32_t)(counter >> 32), (uint32_t)stream, (uint32_t)(stream >> 32),
                (uint32_t)seed, (uint32_t)(seed >> 32), o);
  const double u0 = (((uint64_t)o[0] << 21) ^ (uint64_t)(o[1] >> 11)) * (1.0 / 9007199254740992.0);  // [0,1)
  const double u1 = (((uint64_t)o[2] << 21) ^ (uint64_t)(o[3] >> 11)) * (1.0 / 9007199254740992.0);
  const double rad = sqrt(-2.0 * log(1.0 - u0));   // 1 - u0 in (0, 1]
#if defined(__CUDA_ARCH__)
  double sn, cs;
  sincospi(2.0 * u1, &sn, &cs);                    // cos / sin of 2 pi u1 without the 2 pi product and its range reduction
  z0 = rad * cs;
  z1 = rad * sn;
#else
  const double ang = 6.283185307179586476925 * u1;
  z0 = rad * cos(ang);
  z1 = rad * sin(ang);
#endif
}

// ------------------------------------------------------------------------------------------
// State of one variational campaign (pointers into global memory, caller-owned).
// Adam moments are laid out like the parameters: [Z (n*d, row-major, stride d) | c | ls (d) | os | noise | rho].
// ------------------------------------------------------------------------------------------
struct VarModel {
  double* Z;      // [cap, d]
  double* hyp;    // [d + 4]: c, ls_0..d-1, os, noise, rho (rho unused unless Student-T)
  double* nat1;   // [cap]
  double* nat2;   // [cap, cap] row-major, stride cap
  double* adam_m; // [cap*d + d + 4]
  double* adam_v; // [cap*d + d + 4]
};

struct VarSmem {
  double* Xt;    // d x cap   training inputs (transposed)
  double* Zt;    // d x cap   inducing inputs (transposed), refreshed every epoch
  double* ys;    // cap       standardised targets
  double* mu;    // cap+1
  double* c0;    // cap+1
  double* c1;    // cap+1
  double* c2;    // cap+1  (pair sweep: next pivot columns)
  double* c3;    // cap+1
  double* piv;   // cap+1
  double* fm;    // cap       q(f) mean at the training points
  double* fv;    // cap       q(f) variance
  double* dm;    // cap
  double* dv;    // cap
  double* gmu;   // cap
  double* gsm;   // cap       G_S mu
  double* gZ;    // cap x d   gradient w.r.t. Z (row-major)
  double* small; // 96
  double* red;   // 16 * kMaxWarps
  double* M0;    // (cap+1) x ld: the step-sequential routines (sweep, Cholesky, triangular inverse) run here
  double* panels; // the two panels of the blocked (DMMA) sweep, or nullptr: the pair sweep is used then
  double* stage;  // contract_stage_doubles(): operand ring of the staged contractions; aliases M0 (when M0 is in
                  // shared memory) and the sweep panels: all dead between the factorisations of an epoch and the next
  int cap, ld;
};

// Above this capacity the square M0 no longer fits in shared memory next to the vectors: it then lives in the
// campaign's global workspace (an eighth square, L2-resident); the routines are the same, only slower per step.
constexpr int kVarSmemSquareMax = 144;
// Leading dimension of the work squares (global AND shared: the epoch addresses all of them with one stride): odd
// where the working square is in shared memory (bank-conflict-free row and column walks), EVEN above -- the squares
// are then all global, banks do not matter, and even rows are what the 16-byte cp.async copies of contract_staged need.
STP_HD int var_ld(int cap) { return cap > kVarSmemSquareMax ? ((cap + 2) & ~1) : odd_ld(cap + 1); }
// The part of var_smem_doubles() in front of the working square / stage region (vectors and point sets).
STP_HD size_t var_smem_vector_doubles(int cap, int d) {
  const size_t v = (size_t)2 * d * cap + (size_t)cap * 4 + 6 * (size_t)(cap + 1) + (size_t)cap * d + 96 + 16 * kMaxWarps;
  return (v + 1) & ~(size_t)1;      // the square / stage region starts on a 16-byte boundary
}

// The same without gZ (the last vector): all var_prepare / var_acquire need.  Large capacities start the candidate tile
// of the pool pass here, which is what lets n_cap = 481 (benchmark_predictions.py's 80 % split of 600 points) fit.
STP_HD size_t var_smem_acq_vector_doubles(int cap, int d) {
  const size_t v = (size_t)2 * d * cap + (size_t)cap * 4 + 6 * (size_t)(cap + 1) + 96 + 16 * kMaxWarps;
  return (v + 1) & ~(size_t)1;
}

STP_HD size_t var_smem_doubles(int cap, int d, bool with_panels = false) {
  // the working square + the sweep panels double as the operand ring of the staged contractions
  size_t square = (cap <= kVarSmemSquareMax ? (size_t)(cap + 1) * var_ld(cap) : 0) + (with_panels ? sweep_blocked_scratch_doubles(cap + 1) : 0);
  if (cap > kVarSmemSquareMax && square < contract_stage_doubles()) square = contract_stage_doubles();   // staged contractions: BIG only
  return square + var_smem_vector_doubles(cap, d);
}

// in_smem: where the working square lives (-1: decided by the capacity).  The kernels pass a compile-time constant
// so that every access through M0 is compiled for ONE address space (a pointer that may be shared or global costs
// generic loads / stores in the step-sequential routines: measured 6 % of the VarTGP bench).
STP_HD VarSmem var_smem_carve(double* base, int cap, int d, bool with_panels = false, int in_smem = -1) {
  VarSmem s;
  s.cap = cap; s.ld = var_ld(cap);
  double* p = base;
  s.Xt = p; p += (size_t)d * cap;
  s.Zt = p; p += (size_t)d * cap;
  s.ys = p; p += cap;
  s.mu = p; p += cap + 1;
  s.piv = p; p += cap + 1;
  // the scratch columns of the step-sequential routines (dead once the factorisations of an epoch are done) share
  // their storage with the per-point vectors of the likelihood pass (written after them)
  s.c0 = s.fm = p; p += cap + 1;
  s.c1 = s.fv = p; p += cap + 1;
  s.c2 = s.dm = p; p += cap + 1;
  s.c3 = s.dv = p; p += cap + 1;
  s.gmu = p; p += cap;
  s.gsm = p; p += cap;
  s.small = p; p += 96;
  s.red = p; p += 16 * kMaxWarps;
  s.gZ = p; p += (size_t)cap * d;   // last: the pool pass never touches it and lets its candidate tile start here
  p = base + var_smem_vector_doubles(cap, d);
  const bool square_here = in_smem < 0 ? (cap <= kVarSmemSquareMax) : (in_smem != 0);
  s.M0 = square_here ? p : nullptr;   // nullptr: bound to the workspace by var_bind()
  s.stage = p;
  if (square_here) p += (size_t)(cap + 1) * s.ld;
  s.panels = with_panels ? p : nullptr;
  return s;
}

// Per-campaign global workspace: 7 squares of (cap+1) rows x ld doubles (+ 1 for M0 above kVarSmemSquareMax).
constexpr int kVarSquares = 7;
STP_HD size_t var_work_doubles(int cap) {
  return (size_t)(kVarSquares + (cap > kVarSmemSquareMax ? 1 : 0)) * (cap + 1) * var_ld(cap);
}

struct VarWork {
  double* R[kVarSquares];
  double* M0g;   // the eighth square (large capacities only)
  int ld;
};
STP_HD VarWork var_work_carve(double* base, int cap) {
  VarWork w;
  w.ld = var_ld(cap);
  for (int q = 0; q < kVarSquares; ++q) w.R[q] = base + (size_t)q * (cap + 1) * w.ld;
  w.M0g = cap > kVarSmemSquareMax ? base + (size_t)kVarSquares * (cap + 1) * w.ld : nullptr;
  return w;
}
STP_HD void var_bind(VarSmem& s, const VarWork& w) { if (!s.M0) s.M0 = w.M0g; }

// Bordered sweep of P = -2 Theta_2 (in s.M0): blocked tensor-core form when the launch carved its panels (device,
// shared-memory square, room for two CTAs per SM), pair sweep otherwise.
template <class Team>
STP_HD int var_sweep(const Team& tm, const VarSmem& s, int n) {
  return sweep_bordered4(tm, s.M0, s.ld, n + 1, n, s.c0, s.c1, s.c2, s.c3, s.piv);
}
#if defined(__CUDACC__)
__device__ __forceinline__ int var_sweep(const CtaTeam& tm, const VarSmem& s, int n) {
  if (s.panels) return sweep_bordered_blocked(tm, s.M0, s.ld, n + 1, n, s.panels, s.piv);
  return sweep_bordered4(tm, s.M0, s.ld, n + 1, n, s.c0, s.c1, s.c2, s.c3, s.piv);
}
#endif

// The O(n^3) contractions of the epoch.  STAGED (device, capacities above kVarSmemSquareMax: the n = 256 squares of
// BASELINE config #5 live in HBM / L2): operand panels staged through shared memory by contract_staged.  Otherwise
// every warp streams its operands straight from the L2-resident work squares (contract_mma) -- for n <= 144 the
// staging overhead (one barrier per 16 reduction indices, the copy issue) costs more than it saves: measured 8.6 ms
// against 9.4 ms per 40 epochs at n = 50 and 21.8 against 25.4 at n = 89, but 111 against 90 at n = 256
// (profiles/r02_var_ab.md).  Host emulation: the register-tiled loop.
template <bool STAGED, bool LT, bool SCALE, class Team, class RangeF, class StoreF>
STP_HD void var_contract(const Team& tm, const VarSmem&, int n_rows, int n_cols, int n_red, const double* Lp, int ld_l,
                         const double* Rp, int ld_r, const double* scale, RangeF range, StoreF store) {
  contract_nn<LT, SCALE>(tm, n_rows, n_cols, n_red, Lp, ld_l, Rp, ld_r, scale, range, store);
}
#if defined(__CUDACC__) && !defined(STP_NO_DMMA) && !defined(STP_NO_STAGE)
template <bool STAGED, bool LT, bool SCALE, class RangeF, class StoreF>
__device__ __forceinline__ void var_contract(const CtaTeam& tm, const VarSmem& s, int n_rows, int n_cols, int n_red,
                                             const double* Lp, int ld_l, const double* Rp, int ld_r, const double* scale,
                                             RangeF range, StoreF store) {
  if (STAGED) contract_staged<LT, SCALE, true>(tm, n_rows, n_cols, n_red, Lp, ld_l, Rp, ld_r, scale, range, store, s.stage);
  else contract_nn<LT, SCALE>(tm, n_rows, n_cols, n_red, Lp, ld_l, Rp, ld_r, scale, range, store);
}
#endif

// Cholesky factor L of the positive-definite matrix `fill(i, j, extra)` (j <= i; `extra` = the diagonal jitter of
// the psd_safe_cholesky retry) and its inverse, written zero-filled to global memory: Lout (lower), LIout = L^-1
// (lower), LTout = L^-T (upper); any of them may be null.  Works in s.M0.  Returns 0 or 1 (not PD after 3 retries).
template <class Team, class FillF>
STP_HD int var_factor_fused(const Team& tm, const VarSmem& s, int n, FillF fill, double* Lout, double* LIout,
                            double* LTout, int ldo) {
  const int ld = s.ld, W = tm.warp_width(), nw = tm.nwarps();
  double* M0 = s.M0;
  if (Lout)
    for (int i = tm.warp(); i < n; i += nw)
      for (int j = i + 1 + tm.lane(); j < n; j += W) Lout[i * ldo + j] = 0.0;
  int chol_fail = 1;
  STP_ROLLED for (int attempt = 0; attempt < kPsdSafeTries && chol_fail; ++attempt) {
    const double extra = psd_safe_jitter(attempt);
    for (int i = tm.warp(); i < n; i += nw) {
      double* row = M0 + i * ld;
      for (int j = tm.lane(); j <= i; j += W) row[j] = fill(i, j, extra);
    }
    chol_fail = cholesky_inverse_fused(tm, M0, ld, n, s.c0, s.c1, s.c2, Lout, ldo);
    if (chol_fail) tm.sync();
  }
  if (chol_fail) return 1;
  for (int i = tm.warp(); i < n; i += nw)
    for (int j = tm.lane(); j < n; j += W) {
      if (LIout) LIout[i * ldo + j] = (j <= i) ? M0[i * ld + j] : 0.0;
      if (LTout) LTout[i * ldo + j] = (j >= i) ? M0[j * ld + i] : 0.0;
    }
  tm.sync();
  return 0;
}
template <class Team, class FillF>
STP_HD int var_factor(const Team& tm, const VarSmem& s, int n, FillF fill, double* Lout, double* LIout, double* LTout,
                      int ldo) {
  return var_factor_fused(tm, s, n, fill, Lout, LIout, LTout, ldo);
}
#if defined(__CUDACC__)
// Device form: the blocked tensor-core sweep of the matrix exports L and L^-1 as a by-product (SweepCholExport):
// n / 8 block steps of two barriers instead of n column steps.
template <class FillF>
__device__ __forceinline__ int var_factor(const CtaTeam& tm, const VarSmem& s, int n, FillF fill, double* Lout,
                                          double* LIout, double* LTout, int ldo) {
  if (!s.panels) return var_factor_fused(tm, s, n, fill, Lout, LIout, LTout, ldo);
  const int ld = s.ld, nw = tm.nwarps();
  double* M0 = s.M0;
  for (int i = tm.warp(); i < n; i += nw)       // the parts the export never writes
    for (int j = tm.lane(); j < n; j += 32) {
      if (j > i) { if (Lout) Lout[i * ldo + j] = 0.0; if (LIout) LIout[i * ldo + j] = 0.0; }
      if (j < i && LTout) LTout[i * ldo + j] = 0.0;
    }
  int fail = 1;
  STP_ROLLED for (int attempt = 0; attempt < kPsdSafeTries && fail; ++attempt) {
    const double extra = psd_safe_jitter(attempt);
    for (int i = tm.warp(); i < n; i += nw) {
      double* row = M0 + i * ld;
      for (int j = tm.lane(); j <= i; j += 32) row[j] = fill(i, j, extra);
    }
    fail = sweep_bordered_blocked<true, true>(tm, M0, ld, n, n, s.panels, s.piv, SweepCholExport{Lout, LIout, LTout, ldo});
    if (fail) tm.sync();
  }
  tm.sync();
  return fail ? 1 : 0;
}
#endif

// log p(y | f) and its derivatives for one quadrature node.
struct LikConst {   // per-epoch constants of the likelihood
  int lik;
  double noise, nu, lg_const;   // lg_const: lgamma((nu+1)/2) - lgamma(nu/2) - 0.5 log(nu pi) - 0.5 log noise
  double dnu_const;             // 0.5 psi((nu+1)/2) - 0.5 psi(nu/2) - 1/(2 nu)
  double inv_noise, inv_nu, inv_sqrt_noise;   // reciprocals of the per-epoch constants: one division per node is left
};

STP_HD void lik_node(const LikConst& L, double r, double& lp, double& dlf, double& dln, double& dlnu) {
  if (L.lik == kLikStudentT) {
    // with den = nu noise + r^2:  z2n = r^2 / (nu noise),  1 / (1 + z2n) = nu noise / den  -- one division per node
    const double r2 = r * r;
    const double nn = L.nu * L.noise;
    const double inv_den = 1.0 / (nn + r2);
    const double z2n = r2 * (L.inv_nu * L.inv_noise);
    const double l1p = log1p(z2n);
    lp = L.lg_const - 0.5 * (L.nu + 1.0) * l1p;
    dlf = (L.nu + 1.0) * r * inv_den;
    dln = -0.5 * L.inv_noise + 0.5 * (L.nu + 1.0) * r2 * L.inv_noise * inv_den;
    dlnu = L.dnu_const - 0.5 * l1p + 0.5 * (L.nu + 1.0) * (z2n * L.inv_nu) * (nn * inv_den);
  } else {  // Laplace, b = sqrt(noise)
    const double a = fabs(r);
    lp = L.lg_const - a * L.inv_sqrt_noise;          // lg_const = -log(2 b) for the Laplace likelihood
    dlf = (r > 0.0 ? 1.0 : (r < 0.0 ? -1.0 : 0.0)) * L.inv_sqrt_noise;
    dln = -0.5 * L.inv_noise + 0.5 * a * L.inv_noise * L.inv_sqrt_noise;
    dlnu = 0.0;
  }
}

struct VarEpochOut {
  double loss;
  int status;   // 0 ok, 1 not PD (P or K_ZZ), 2 non-finite
};

// ------------------------------------------------------------------------------------------
// One training epoch.  gh_x / gh_w: the 20 Gauss-Hermite nodes / weights (fp32-rounded, App. D-9).
// `t` is the 1-based Adam step.  lr_ngd is the reference's `lr` argument (scaled by n inside, App.
// D-10); lr_adam is the hard-coded 0.01 of optim.py:140.  If `apply` is 0 the parameters are left
// untouched and the raw gradients are exported through `dbg` (tests): layout
// [g_nat1 (cap) | g_nat2 (cap*cap) | gZ (cap*d) | g_c | g_ls (d) | g_os | g_noise | g_rho].
// ------------------------------------------------------------------------------------------
template <bool STAGED = false, class Team>
STP_HD VarEpochOut var_epoch(const Team& tm, const VarSmem& s, const VarWork& w, const VarModel& M, int n, int d,
                             int lik, const VarHyperPrior& pr, const double* gh_x, const double* gh_w, int t,
                             double lr_ngd, double lr_adam, int apply, double* dbg) {
  const int ld = w.ld, cap = s.cap, W = tm.warp_width(), nw = tm.nwarps();
  VarEpochOut out; out.loss = NAN; out.status = 0;
  double* S = w.R[0];    // P -> -S (lower, bordered) -> full S ; later G_S ; later Cbar_t
  double* LK = w.R[1];   // L_K (lower)
  double* LI = w.R[2];   // Linv_K (lower, strict upper zero)
  double* LT = w.R[6];   // Linv_K^T (upper, strict lower zero)
  double* Ct = w.R[3];   // Ct[c][j] = k(x_c, z_j)
  double* At = w.R[4];   // At[c][k] ; later Phi ; later Kbar'
  double* Dt = w.R[5];   // Tt -> Dt -> Abar_t ; later Lbar ; later U
  double* hyp = s.small;          // c, ls[d], os, noise, rho  (d + 4)
  double* invl = s.small + 16;    // d
  double* ghyp = s.small + 32;    // gradient of hyp (d + 4)
  const int nh = d + 4;
  tm.sync();
  STP_PHASE(-1); STP_COUNT(48);
  for (int k = tm.rank(); k < nh; k += tm.size()) hyp[k] = M.hyp[k];
  for (int idx = tm.rank(); idx < n * d; idx += tm.size()) {
    const int i = idx / d, k = idx - i * d;
    s.Zt[k * cap + i] = M.Z[idx];
  }
  tm.sync();
  const double cst = hyp[0], os = hyp[1 + d], noise = hyp[2 + d], rho = hyp[3 + d];
  for (int k = tm.rank(); k < d; k += tm.size()) invl[k] = 1.0 / hyp[1 + k];
  int bad = !(os == os) || !(noise == noise) || !(cst == cst);
  for (int k = 0; k < d; ++k) bad |= !(hyp[1 + k] == hyp[1 + k]) || hyp[1 + k] == 0.0;
  if (bad) { out.status = 2; return out; }
  tm.sync();

  // (1) P = -2 Theta2, bordered with theta1; sweep IN SHARED MEMORY -> -S, mu = S theta1, pivots
  double* M0 = s.M0;
  for (int i = tm.warp(); i <= n; i += nw) {
    double* row = M0 + i * ld;
    if (i < n) for (int j = tm.lane(); j <= i; j += W) row[j] = -2.0 * M.nat2[(size_t)i * cap + j];
    else for (int j = tm.lane(); j <= n; j += W) row[j] = (j < n) ? M.nat1[j] : 0.0;
  }
  tm.sync();
  if (var_sweep(tm, s, n)) { out.status = 1; return out; }
  double kl_acc[4] = {0.0, 0.0, 0.0, 0.0};   // sum log piv(P), tr S, mu.mu
  for (int i = tm.rank(); i < n; i += tm.size()) {
    const double m = M0[n * ld + i];
    s.mu[i] = m;
    kl_acc[0] += log(s.piv[i]);
    kl_acc[1] += -M0[i * ld + i];
    kl_acc[2] += m * m;
  }
  tm.sum_vec(kl_acc, 3);
  // full symmetric S = -(swept) into global memory (operand of the Tt contraction)
  for (int i = tm.warp(); i < n; i += nw)
    for (int j = tm.lane(); j < n; j += W) S[i * ld + j] = (j <= i) ? -M0[i * ld + j] : -M0[j * ld + i];
  tm.sync();
  STP_PHASE(33);
  // (2) K_ZZ + jitter: fused Cholesky + in-place inverse in shared memory (one barrier per column), L_K streamed
  //     out (strict upper zero-filled); then Linv (lower) and Linv^T (upper), zero-filled, to global memory; Ct
  for (int c = tm.warp(); c < n; c += nw) {
    double* row = Ct + c * ld;
    for (int j = tm.lane(); j < n; j += W) {
      double r2 = 0.0;
#pragma unroll
      for (int k = 0; k < kMaxD; ++k)
        if (k < d) {
          const double tt = (s.Xt[k * cap + c] - s.Zt[k * cap + j]) * invl[k];
          r2 += tt * tt;
        }
      row[j] = os * exp(-0.5 * r2);
    }
  }
  {
    const double* Zt = s.Zt;
    const double diag0 = os + kVarJitter;
    // the factorisations only touch the lower triangle of the working square: when it lives in shared memory its strict
    // upper triangle keeps the noise-free K_ZZ entries for the kernel-gradient pass of step (13) (one exp per pair saved)
    double* keep = STAGED ? nullptr : M0;
    if (var_factor(tm, s, n, [=](int i, int j, double extra) {
          if (j < i) {
            const double k = ard_rbf_tt(Zt, cap, d, invl, i, j, os);
            if (keep) keep[j * ld + i] = k;
            return k;
          }
          return diag0 + extra;
        }, LK, LI, LT, ld)) { out.status = 1; return out; }
  }
  STP_PHASE(34); STP_PHASE(35);
  // (3) At[c][k] = sum_{j<=k} Ct[c][j] Linv^T[j][k]
  var_contract<STAGED, false, false>(tm, s, n, n, n, Ct, ld, LT, ld, (const double*)nullptr,
                 [=](int, int, int, int k_hi, int& r0, int& r1) { r0 = 0; r1 = k_hi + 1; },
                 [=](int c, int k, double v) { At[c * ld + k] = v; });
  tm.sync();
  STP_PHASE(36);
  // (4) Tt = At S ; Dt = Tt - At ; q(f): m_c = cst + At[c].mu ; v_c = os + jitter + At[c].Dt[c]
  var_contract<STAGED, false, false>(tm, s, n, n, n, At, ld, S, ld, (const double*)nullptr,
                 [=](int, int, int, int, int& r0, int& r1) { r0 = 0; r1 = n; },
                 [=](int c, int k, double v) { Dt[c * ld + k] = v - At[c * ld + k]; });
  tm.sync();
  for (int c = tm.warp(); c < n; c += nw) {
    const double* arow = At + c * ld;
    const double* drow = Dt + c * ld;
    double pm = 0.0, pv = 0.0;
    for (int k = tm.lane(); k < n; k += W) {
      const double a = arow[k];
      pm += a * s.mu[k];
      pv += a * drow[k];
    }
    pm = tm.warp_sum(pm);
    pv = tm.warp_sum(pv);
    if (tm.lane() == 0) { s.fm[c] = cst + pm; s.fv[c] = os + kVarJitter + pv; }
  }
  tm.sync();
  STP_PHASE(37);
  // (5) expected log-likelihood and its derivatives per training point
  // likelihood constants (lgamma / digamma: thousands of fp64 instructions): thread 0 evaluates them, the others read
  // them from shared memory instead of re-deriving them 256 times over on the shared fp64 pipe
  LikConst LC; LC.lik = lik; LC.noise = noise; LC.nu = 0.0; LC.lg_const = 0.0; LC.dnu_const = 0.0;
  LC.inv_noise = 1.0 / noise; LC.inv_nu = 0.0; LC.inv_sqrt_noise = 1.0 / sqrt(noise);
  if (lik == kLikLaplace) LC.lg_const = -log(2.0 * sqrt(noise));
  double nu = 0.0;
  if (lik == kLikStudentT) {
    double* lc = s.small + 48;
    if (tm.rank() == 0) {
      const double nu0 = 2.0 + softplus(rho);
      lc[0] = nu0;
      lc[1] = lgamma(0.5 * (nu0 + 1.0)) - lgamma(0.5 * nu0) - 0.5 * log(nu0 * 3.14159265358979323846) - 0.5 * log(noise);
      lc[2] = 0.5 * digamma_pos(0.5 * (nu0 + 1.0)) - 0.5 * digamma_pos(0.5 * nu0) - 0.5 / nu0;
    }
    tm.sync();
    nu = lc[0];
    LC.nu = nu; LC.lg_const = lc[1]; LC.dnu_const = lc[2]; LC.inv_nu = 1.0 / nu;
  }
  const double inv_n = 1.0 / n;
  double lk_acc[4] = {0.0, 0.0, 0.0, 0.0};   // ell, d ell / d noise, d ell / d nu, sum dv
  if (lik == kLikGaussian) {
    for (int i = tm.rank(); i < n; i += tm.size()) {
      const double m = s.fm[i], v = s.fv[i], yy = s.ys[i];
      const double r = yy - m;
      const double ell = -0.5 * ((r * r + v) / noise + log(noise) + kLog2Pi);
      const double dm_ = r / noise;
      const double dv_ = -0.5 / noise;
      const double dn_ = 0.5 * ((r * r + v) / (noise * noise) - 1.0 / noise);
      s.dm[i] = -inv_n * dm_;
      s.dv[i] = -inv_n * dv_;
      lk_acc[0] += ell; lk_acc[1] += dn_; lk_acc[3] += -inv_n * dv_;
    }
  } else {
    // Gauss-Hermite quadrature: four lanes share a training point, five of the 20 nodes each (a node costs a log1p
    // and four divisions); a warp takes eight points at a time, the partial sums meet through two shuffles.
    const double kInvSqrtPi = 0.56418958354775628695;
    const int G4 = W >= 4 ? 4 : 1;                 // lanes per point (1 in the host emulation)
    const int ppw = W / G4;                        // points per warp and pass
    const int sub = tm.lane() % G4, slot = tm.lane() / G4;
    for (int i0 = tm.warp() * ppw; i0 < n; i0 += nw * ppw) {
      const int i = i0 + slot;
      const bool on = i < n;
      const double m = on ? s.fm[i] : 0.0, v = on ? s.fv[i] : 1.0, yy = on ? s.ys[i] : 0.0;
      const double sq = sqrt(2.0 * v);
      double ell = 0.0, dm_ = 0.0, dv_ = 0.0, dn_ = 0.0, dnu_ = 0.0;
      for (int q = sub; q < kGH; q += G4) {
        const double f = sq * gh_x[q] + m;
        double lp, dlf, dln, dlnu;
        lik_node(LC, yy - f, lp, dlf, dln, dlnu);
        const double wq = kInvSqrtPi * gh_w[q];
        ell += wq * lp;
        dm_ += wq * dlf;
        dv_ += wq * dlf * gh_x[q];
        dn_ += wq * dln;
        dnu_ += wq * dlnu;
      }
      ell = tm.quad_sum(ell); dm_ = tm.quad_sum(dm_); dv_ = tm.quad_sum(dv_);
      dn_ = tm.quad_sum(dn_); dnu_ = tm.quad_sum(dnu_);
      dv_ /= sq;
      if (on && sub == 0) {
        s.dm[i] = -inv_n * dm_;
        s.dv[i] = -inv_n * dv_;
        lk_acc[0] += ell; lk_acc[1] += dn_; lk_acc[2] += dnu_; lk_acc[3] += -inv_n * dv_;
      }
    }
  }
  tm.sum_vec(lk_acc, 4);
  // loss
  const double logdetS = -kl_acc[0];
  const double kl = 0.5 * (-logdetS + kl_acc[1] + kl_acc[2] - n);
  double lprior = lognormal_logpdf(os, pr.os_loc, pr.os_scale) + lognormal_logpdf(noise, pr.noise_loc, pr.noise_scale);
  for (int k = 0; k < d; ++k) lprior += lognormal_logpdf(hyp[1 + k], pr.ls_loc, pr.ls_scale);
  if (lik == kLikStudentT)
    lprior += pr.df_conc * log(pr.df_rate) + (pr.df_conc - 1.0) * log(nu) - pr.df_rate * nu - lgamma(pr.df_conc);
  out.loss = -(lk_acc[0] - kl + lprior) * inv_n;
  if (!(out.loss == out.loss) || !(fabs(out.loss) <= 1.79e308)) { out.status = 2; return out; }

  STP_PHASE(38);
  // (6) variational gradients: g_mu = A dm + mu/n ; G_S = A diag(dv) A^T + (I - P)/(2n)  (into R0)
  for (int k = tm.rank(); k < n; k += tm.size()) {
    double a = 0.0;
    for (int c = 0; c < n; ++c) a += At[c * ld + k] * s.dm[c];
    s.gmu[k] = a + s.mu[k] * inv_n;
  }
  double* GS = S;
  tm.sync();   // every reader of S (step 4) is done
  {
    const double* dvv = s.dv;
    const double* nat2 = M.nat2;
    var_contract<STAGED, true, true>(tm, s, n, n, n, At, ld, At, ld, dvv,
                   [=](int, int i_hi, int k_lo, int, int& r0, int& r1) { r0 = 0; r1 = (k_lo > i_hi) ? 0 : n; },
                   [=](int i, int k, double v) {
                     if (k <= i) {
                       const double P = -2.0 * nat2[(size_t)i * cap + k];
                       const double g = v + ((i == k ? 1.0 : 0.0) - P) * (0.5 * inv_n);
                       GS[i * ld + k] = g;
                       GS[k * ld + i] = g;
                     }
                   });
  }
  tm.sync();
  for (int i = tm.warp(); i < n; i += nw) {
    double a = 0.0;
    for (int k = tm.lane(); k < n; k += W) a += GS[i * ld + k] * s.mu[k];
    a = tm.warp_sum(a);
    if (tm.lane() == 0) s.gsm[i] = a;
  }
  tm.sync();
  if (dbg) {
    for (int i = tm.rank(); i < n; i += tm.size()) dbg[i] = s.gmu[i] - 2.0 * s.gsm[i];
    for (int idx = tm.rank(); idx < n * n; idx += tm.size()) {
      const int i = idx / n, k = idx - i * n;
      dbg[cap + (size_t)i * cap + k] = GS[i * ld + k];
    }
  }
  if (apply) {
    // natural-gradient step (optim.py:132-134, 169): theta <- theta - lr * n * grad
    const double step = lr_ngd * n;
    for (int i = tm.rank(); i < n; i += tm.size()) M.nat1[i] -= step * (s.gmu[i] - 2.0 * s.gsm[i]);
    for (int i = tm.warp(); i < n; i += nw)
      for (int k = tm.lane(); k < n; k += W) M.nat2[(size_t)i * cap + k] -= step * GS[i * ld + k];
  }
  tm.sync();
  STP_PHASE(39);
  // (7) Abar_t[c][k] = mu_k dm_c + 2 Dt[c][k] dv_c   (in place over Dt)
  for (int c = tm.warp(); c < n; c += nw) {
    double* drow = Dt + c * ld;
    const double dmc = s.dm[c], dvc2 = 2.0 * s.dv[c];
    for (int k = tm.lane(); k < n; k += W) drow[k] = s.mu[k] * dmc + drow[k] * dvc2;
  }
  tm.sync();
  // (8) Cbar_t[c][j] = sum_{i>=j} Abar_t[c][i] Linv[i][j]   (into R0; G_S is dead)
  //     and ET[j][c] = Cbar_t[c][j] Ct[c][j], the transposed elementwise product step (13) walks by rows (into R6;
  //     Linv_K^T is dead after step 3)
  double* CBt = S;
  double* ET = LT;
  var_contract<STAGED, false, false>(tm, s, n, n, n, Dt, ld, LI, ld, (const double*)nullptr,
                 [=](int, int, int k_lo, int, int& r0, int& r1) { r0 = k_lo; r1 = n; },
                 [=](int c, int j, double v) { CBt[c * ld + j] = v; ET[j * ld + c] = v * Ct[c * ld + j]; });
  tm.sync();
  STP_PHASE(40);
  // (9) Lbar[i][k] = -sum_c Cbar_t[c][i] At[c][k], i >= k, zero above   (into R5, Abar_t is dead)
  double* LB = Dt;
  var_contract<STAGED, true, false>(tm, s, n, n, n, CBt, ld, At, ld, (const double*)nullptr,
                 [=](int, int i_hi, int k_lo, int, int& r0, int& r1) { r0 = 0; r1 = (k_lo > i_hi) ? 0 : n; },
                 [=](int i, int k, double v) { LB[i * ld + k] = (k <= i) ? -v : 0.0; });
  tm.sync();
  STP_PHASE(41);
  // (10) Phi[i][k] = sum_{r>=i} L[r][i] Lbar[r][k], i >= k, diagonal halved, zero above   (into R4, At is dead)
  double* PH = At;
  var_contract<STAGED, true, false>(tm, s, n, n, n, LK, ld, LB, ld, (const double*)nullptr,
                 [=](int i_lo, int i_hi, int k_lo, int, int& r0, int& r1) { r0 = i_lo; r1 = (k_lo > i_hi) ? i_lo : n; },
                 [=](int i, int k, double v) { PH[i * ld + k] = (k < i) ? v : (k == i ? 0.5 * v : 0.0); });
  tm.sync();
  STP_PHASE(42);
  // (11) U[i][k] = sum_{k<=r<=i} Phi[i][r] Linv[r][k], zero above   (into R5, Lbar is dead)
  double* U = Dt;
  var_contract<STAGED, false, false>(tm, s, n, n, n, PH, ld, LI, ld, (const double*)nullptr,
                 [=](int, int i_hi, int k_lo, int, int& r0, int& r1) { r0 = k_lo; r1 = (k_lo > i_hi) ? k_lo : i_hi + 1; },
                 [=](int i, int k, double v) { U[i * ld + k] = (k <= i) ? v : 0.0; });
  tm.sync();
  STP_PHASE(43);
  // (12) Kbar'[j][k] = sum_{i>=max(j,k)} Linv[i][j] U[i][k]   (into R4, Phi is dead)
  //      and its transpose (into R1; L_K is dead after step 10) so that step (13) reads both by rows
  double* KB = At;
  double* KBT = LK;
  var_contract<STAGED, true, false>(tm, s, n, n, n, LI, ld, U, ld, (const double*)nullptr,
                 [=](int j_lo, int, int k_lo, int, int& r0, int& r1) { r0 = (j_lo > k_lo) ? j_lo : k_lo; r1 = n; },
                 [=](int j, int k, double v) { KB[j * ld + k] = v; KBT[k * ld + j] = v; });
  tm.sync();
  STP_PHASE(44);
  // (13) kernel-parameter gradients.  Kbar = sym(Kbar'); K0 and C are re-evaluated on the fly.
  double acc[kMaxD + 4];
#pragma unroll
  for (int q = 0; q < kMaxD + 4; ++q) acc[q] = 0.0;   // acc[0]: sum (Kbar.K0 + Cbar.C); acc[1+k]: lengthscale sums
  for (int idx = tm.rank(); idx < n * d; idx += tm.size()) s.gZ[idx] = 0.0;
  tm.sync();
  for (int i = tm.warp(); i < n; i += nw) {
    double gz[kMaxD];
#pragma unroll
    for (int k = 0; k < kMaxD; ++k) gz[k] = 0.0;
    for (int j = tm.lane(); j < n; j += W) {
      // K part: pair (i, j) of inducing points, i != j (the diagonal has no dependence on l or Z)
      if (j != i) {
        const double kb = 0.5 * (KB[i * ld + j] + KBT[i * ld + j]);
        const double k0 = STAGED ? ard_rbf_tt(s.Zt, cap, d, invl, i, j, os) : ((i < j) ? M0[i * ld + j] : M0[j * ld + i]);
        const double tK = kb * k0;
        acc[0] += tK;
#pragma unroll
        for (int k = 0; k < kMaxD; ++k)
          if (k < d) {
            const double dz = s.Zt[k * cap + i] - s.Zt[k * cap + j];
            acc[1 + k] += tK * dz * dz;
            gz[k] += 2.0 * tK * dz;
          }
      } else {
        acc[0] += KB[i * ld + i] * os;
      }
      // C part: inducing point i against data point j : (Cbar . C)[i][j] = Cbar_t[j][i] Ct[j][i]
      const double tC = ET[i * ld + j];
      acc[0] += tC;
#pragma unroll
      for (int k = 0; k < kMaxD; ++k)
        if (k < d) {
          const double dz = s.Zt[k * cap + i] - s.Xt[k * cap + j];
          acc[1 + k] += tC * dz * dz;
          gz[k] += tC * dz;
        }
    }
#pragma unroll
    for (int k = 0; k < kMaxD; ++k)
      if (k < d) {
        const double gsum = tm.warp_sum(gz[k]);
        if (tm.lane() == 0) s.gZ[i * d + k] = -gsum * invl[k] * invl[k];
      }
  }
  tm.sum_vec(acc, 1 + d);
  STP_PHASE(45);
  double csum = 0.0;
  for (int i = tm.rank(); i < n; i += tm.size()) csum += s.dm[i];
  csum = tm.sum(csum);
  tm.sync();
  if (tm.rank() == 0) {
    ghyp[0] = csum;
    for (int k = 0; k < d; ++k) {
      const double l = hyp[1 + k];
      ghyp[1 + k] = acc[1 + k] / (l * l * l) - inv_n * lognormal_dlogpdf(l, pr.ls_loc, pr.ls_scale);
    }
    ghyp[1 + d] = lk_acc[3] + acc[0] / os - inv_n * lognormal_dlogpdf(os, pr.os_loc, pr.os_scale);
    ghyp[2 + d] = -inv_n * lk_acc[1] - inv_n * lognormal_dlogpdf(noise, pr.noise_loc, pr.noise_scale);
    double grho = 0.0;
    if (lik == kLikStudentT) {
      const double dnu = -inv_n * lk_acc[2] - inv_n * ((pr.df_conc - 1.0) / nu - pr.df_rate);
      grho = dnu * sigmoid(rho);
    }
    ghyp[3 + d] = grho;
  }
  tm.sync();
  if (dbg) {
    double* g = dbg + cap + (size_t)cap * cap;
    for (int idx = tm.rank(); idx < n * d; idx += tm.size()) g[idx] = s.gZ[idx];
    for (int k = tm.rank(); k < nh; k += tm.size()) g[(size_t)cap * d + k] = ghyp[k];
  }
  if (apply) {
    // (14) Adam (torch.optim.Adam defaults; lr = 0.01 hard-coded at optim.py:140)
    const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
    // the two bias corrections (an fp64 pow each) once per CTA, not once per thread
    double* bc = s.small + 52;
    if (tm.rank() == 0) { bc[0] = 1.0 - pow(b1, (double)t); bc[1] = sqrt(1.0 - pow(b2, (double)t)); }
    tm.sync();
    const double bc1 = bc[0], bc2s = bc[1];
    const double step_size = lr_adam / bc1;
    const int nadam = (lik == kLikStudentT) ? nh : nh - 1;
    for (int idx = tm.rank(); idx < n * d + nadam; idx += tm.size()) {
      const int isz = idx < n * d;
      const double g = isz ? s.gZ[idx] : ghyp[idx - n * d];
      const size_t slot = isz ? (size_t)idx : (size_t)cap * d + (idx - n * d);
      double m = M.adam_m[slot], v = M.adam_v[slot];
      m = m + (g - m) * (1.0 - b1);              // exp_avg.lerp_(grad, 1 - beta1)
      v = v * b2 + (1.0 - b2) * g * g;
      M.adam_m[slot] = m; M.adam_v[slot] = v;
      const double denom = sqrt(v) / bc2s + eps;
      const double upd = step_size * (m / denom);
      if (isz) M.Z[idx] -= upd; else M.hyp[idx - n * d] -= upd;
    }
  }
  tm.sync();
  STP_PHASE(46);
  return out;
}

// ------------------------------------------------------------------------------------------
// Prepare the pool pass at the trained state:
//   R1 <- Linv_K mirrored, R2 <- G = L_P^-1 Linv_K (lower), s.gmu <- w = Linv_K^T mu, s.mu <- mu.
// v_* = os + jitter + ||G k||^2 - ||Linv_K k||^2, m_* = cst + k.w   (App. A.4 step 3 at a candidate)
// Returns 0, 1 (not PD) or 2 (non-finite hyper-parameter).
// ------------------------------------------------------------------------------------------
template <bool STAGED = false, class Team>
STP_HD int var_prepare(const Team& tm, const VarSmem& s, const VarWork& w, const VarModel& M, int n, int d,
                       bool pack_in_smem = false) {
  const int ld = w.ld, cap = s.cap, W = tm.warp_width(), nw = tm.nwarps();
  double* LI = w.R[1];    // Linv_K mirrored
  double* G = w.R[2];
  double* LPI = w.R[3];
  double* hyp = s.small;
  double* invl = s.small + 16;
  const int nh = d + 4;
  tm.sync();
  for (int k = tm.rank(); k < nh; k += tm.size()) hyp[k] = M.hyp[k];
  for (int idx = tm.rank(); idx < n * d; idx += tm.size()) {
    const int i = idx / d, k = idx - i * d;
    s.Zt[k * cap + i] = M.Z[idx];
  }
  tm.sync();
  const double os = hyp[1 + d];
  for (int k = tm.rank(); k < d; k += tm.size()) invl[k] = 1.0 / hyp[1 + k];
  int bad = !(os == os) || !(hyp[2 + d] == hyp[2 + d]) || !(hyp[0] == hyp[0]);
  for (int k = 0; k < d; ++k) bad |= !(hyp[1 + k] == hyp[1 + k]) || hyp[1 + k] == 0.0;
  if (bad) return 2;
  tm.sync();
  // chol(P)^-1 and chol(K_ZZ + jitter I)^-1 (factorised in s.M0, written out zero-filled)
  {
    const double* nat2 = M.nat2;
    const double* Zt = s.Zt;
    if (var_factor(tm, s, n, [=](int i, int j, double extra) {
          return -2.0 * nat2[(size_t)i * cap + j] + ((j == i) ? extra : 0.0);
        }, (double*)nullptr, LPI, (double*)nullptr, ld)) return 1;
    if (var_factor(tm, s, n, [=](int i, int j, double extra) {
          return (j == i) ? os + kVarJitter + extra : ard_rbf_tt(Zt, cap, d, invl, i, j, os);
        }, (double*)nullptr, LI, (double*)nullptr, ld)) return 1;
  }
  // mu = S theta1 = LPI^T (LPI theta1)
  for (int i = tm.rank(); i < n; i += tm.size()) {
    double a = 0.0;
    for (int j = 0; j <= i; ++j) a += LPI[i * ld + j] * M.nat1[j];
    s.c1[i] = a;
  }
  tm.sync();
  for (int j = tm.rank(); j < n; j += tm.size()) {
    double a = 0.0;
    for (int i = j; i < n; ++i) a += LPI[i * ld + j] * s.c1[i];
    s.mu[j] = a;
  }
  tm.sync();
  // w = Linv_K^T mu ; G[i][k] = sum_{k<=r<=i} LPI[i][r] Linv_K[r][k]  (lower, zero above)
  for (int j = tm.rank(); j < n; j += tm.size()) {
    double a = 0.0;
    for (int i = j; i < n; ++i) a += LI[i * ld + j] * s.mu[i];
    s.gmu[j] = a;
  }
  var_contract<STAGED, false, false>(tm, s, n, n, n, LPI, ld, LI, ld, (const double*)nullptr,
                 [=](int, int i_hi, int k_lo, int, int& r0, int& r1) { r0 = k_lo; r1 = (k_lo > i_hi) ? k_lo : i_hi + 1; },
                 [=](int i, int k, double v) { G[i * ld + k] = (k <= i) ? v : 0.0; });
  tm.sync();
  if (pack_in_smem) {
    // The pool pass multiplies every candidate tile by Linv_K and by G: keep BOTH triangles in the shared-memory
    // square (free after the factorisations) -- Linv_K in the lower triangle, G transposed in the upper one, shifted
    // by one column so that the two diagonals do not collide (ld >= n + 1):  M0[i][k] = Linv_K[i][k],
    // M0[k][i + 1] = G[i][k]  (k <= i).  No global load is left in the pool pass.
    double* M0 = s.M0;
    for (int i = tm.warp(); i < n; i += nw)
      for (int k = tm.lane(); k <= i; k += W) {
        M0[i * ld + k] = LI[i * ld + k];
        M0[k * ld + i + 1] = G[i * ld + k];
      }
    tm.sync();
  }
  return 0;
}

// Likelihood scale used by the acquisition (App. A.5): sqrt(max(var_lik, 1e-12)).
STP_HD double var_lik_variance(int lik, double noise, double rho) {
  if (lik == kLikStudentT) { const double nu = 2.0 + softplus(rho); return noise * nu / (nu - 2.0); }
  if (lik == kLikLaplace) return 2.0 * noise;
  return noise;
}

// Partial-sum slots per candidate of the pool pass: one per 16-row tile of the triangular factors (tensor-core form)
// or per warp (generic form), whichever is larger.  Shared-memory need of var_acquire, in doubles: tile
// (cap16 + kMaxD) x TC, partial sums (2 slots + 8) x TC.
STP_HD int var_acq_slots(int cap) { const int rt = (cap + 15) >> 4; return rt > 8 ? rt : 8; }
STP_HD size_t var_acq_smem_doubles(int cap, int tc) {
  return (size_t)(((cap + 15) & ~15) + kMaxD) * tc + (size_t)(2 * var_acq_slots(cap) + 8) * tc;
}

// Partial sums of one candidate tile (TC candidates, cross-covariances in tile[j * TC + c]):
//   part[p * TC + c], p < n_slots: pieces of ||Linv_K k_c||^2;  part[(slots + p) * TC + c]: of ||G k_c||^2;
//   part[(2 slots + w) * TC + c], w < min(nwarps, 8): of gmu . k_c.
// Generic form: warp w takes the rows i = w, w + nw, ...; a lane is a candidate.
template <class Team>
STP_HD void var_tile_partials(const Team& tm, const VarSmem& s, const VarWork& w, int n, const double* tile, double* part,
                              int slots, bool packed, int& n_slots) {
  const int ld = w.ld, TC = tm.warp_width(), nw = tm.nwarps();
  const int nwa = nw < 8 ? nw : 8;
  const double* LI = w.R[1];
  const double* G = w.R[2];
  const double* M0 = s.M0;
  n_slots = nwa;
  if (tm.warp() < nwa) {
    const int c = tm.lane();
    double sa = 0.0, sb = 0.0, mu = 0.0;
    for (int i = tm.warp(); i < n; i += nwa) {
      double a = 0.0, b = 0.0;
      for (int j = 0; j <= i; ++j) {
        const double kj = tile[j * TC + c];
        a += (packed ? M0[i * ld + j] : LI[i * ld + j]) * kj;
        b += (packed ? M0[j * ld + i + 1] : G[i * ld + j]) * kj;
      }
      sa += a * a;
      sb += b * b;
      mu += s.gmu[i] * tile[i * TC + c];
    }
    part[tm.warp() * TC + c] = sa;
    part[(slots + tm.warp()) * TC + c] = sb;
    part[(2 * slots + tm.warp()) * TC + c] = mu;
  }
}
#if defined(__CUDACC__) && !defined(STP_NO_DMMA)
// Tensor-core form: V = [Linv_K; G] K_tile in 16 x 16 DMMA tiles (both factors are lower triangular: the reduction of
// row tile I stops at its diagonal block, which is masked), squares summed per row tile.  PACKED: both factors come
// from the shared-memory square (see var_prepare); otherwise from their global (L2-resident) copies (M = 256).
template <bool PACKED>
__device__ __forceinline__ void var_tile_partials_mma(const CtaTeam& tm, const VarSmem& s, const VarWork& w, int n,
                                                      const double* tile, double* part, int slots, int& n_slots) {
  constexpr int TC = 32;
  const int ld = w.ld, nw = tm.nwarps(), warp = tm.warp(), lane = tm.lane();
  const int g = lane >> 2, q = lane & 3;
  const int RT = (n + 15) >> 4;
  const int nwa = nw < 8 ? nw : 8;
  const double* LI = PACKED ? s.M0 : w.R[1];
  const double* G = PACKED ? s.M0 : w.R[2];
  n_slots = RT;
  if (warp < nwa) {
    double mu = 0.0;
    for (int i = warp; i < n; i += nwa) mu += s.gmu[i] * tile[i * TC + lane];
    part[(2 * slots + warp) * TC + lane] = mu;
  }
  for (int t = warp; t < 4 * RT; t += nw) {
    const int mat = t & 1, jb = ((t >> 1) & 1) << 4, I = RT - 1 - (t >> 2), ib = I << 4;   // heaviest row tiles first
    double c[2][2][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) { c[a][b][0] = 0.0; c[a][b][1] = 0.0; }
    const int i0 = ib + g, i1 = ib + 8 + g;
    const int r0 = (i0 < n) ? i0 : n - 1, r1 = (i1 < n) ? i1 : n - 1;
    // element (i, k) of the factor: row-major lower triangle, or (PACKED, G) the shifted transposed upper triangle
    const double* p0 = (PACKED && mat) ? G + r0 + 1 : (mat ? G : LI) + (size_t)r0 * ld;
    const double* p1 = (PACKED && mat) ? G + r1 + 1 : (mat ? G : LI) + (size_t)r1 * ld;
    const int ks = (PACKED && mat) ? ld : 1;
    const double* tb = tile + jb + g;
    for (int k0 = 0; k0 < ib; k0 += 4) {           // strictly below the diagonal block: no masks
      const int kk = k0 + q;
      const double a0 = p0[(size_t)kk * ks], a1 = p1[(size_t)kk * ks];
      const double b0 = tb[kk * TC], b1 = tb[kk * TC + 8];
      dmma884(c[0][0][0], c[0][0][1], a0, b0);
      dmma884(c[0][1][0], c[0][1][1], a0, b1);
      dmma884(c[1][0][0], c[1][0][1], a1, b0);
      dmma884(c[1][1][0], c[1][1][1], a1, b1);
    }
#pragma unroll
    for (int k0 = 0; k0 < 16; k0 += 4) {            // the diagonal block
      const int kk = ib + k0 + q;
      const int kc = (kk < n) ? kk : n - 1;
      const double a0 = (i0 < n && kk <= i0) ? p0[(size_t)kc * ks] : 0.0, a1 = (i1 < n && kk <= i1) ? p1[(size_t)kc * ks] : 0.0;
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
        if (g == 0) part[(mat * slots + I) * TC + jb + 8 * b + 2 * q + e] = sq;
      }
  }
}
__device__ __forceinline__ void var_tile_partials(const CtaTeam& tm, const VarSmem& s, const VarWork& w, int n,
                                                  const double* tile, double* part, int slots, bool packed, int& n_slots) {
  if (packed) var_tile_partials_mma<true>(tm, s, w, n, tile, part, slots, n_slots);
  else var_tile_partials_mma<false>(tm, s, w, n, tile, part, slots, n_slots);
}
#endif

// ------------------------------------------------------------------------------------------
// Fused pool pass of a variational model (after var_prepare).  `tile` and `part`: var_acq_smem_doubles(cap, TC)
// doubles of shared memory in all (part = tile + (cap16 + kMaxD) x TC).  MC models: `eps` (global, [N, S]) if non-null,
// else Philox(seed, stream = campaign id, counter = candidate * (S/2) + pair).
// Optional outputs (may be null): mean_out / var_out = q(f) moments (no likelihood noise), acq_out.
// ------------------------------------------------------------------------------------------
template <class Team>
STP_HD void var_acquire(const Team& tm, const VarSmem& s, const VarWork& w, int n, int d, int lik, double* tile,
                        double* part, const double* pool, int N, const uint8_t* mask, double best_f, int S,
                        const double* eps, uint64_t seed, uint64_t stream, int* best_idx, double* best_acq,
                        double* mean_out, double* var_out, double* acq_out, bool packed = false,
                        int acq_kind = kAcqLogEI, double acq_beta = 0.0) {
  const int cap = s.cap, TC = tm.warp_width(), nw = tm.nwarps();
  const int n16 = (n + 15) & ~15;                 // rows n .. n16 - 1 of the tile are zero (DMMA row tiles)
  const int slots = var_acq_slots(cap);           // partial-sum slots per candidate and quantity
  const double* hyp = s.small;
  const double* invl = s.small + 16;
  const double cst = hyp[0], os = hyp[1 + d], noise = hyp[2 + d], rho = hyp[3 + d];
  double* xc = tile + (size_t)((cap + 15) & ~15) * TC;
  const double lik_var = var_lik_variance(lik, noise, rho);
  const double sig_l = sqrt(lik_var > 1e-12 ? lik_var : 1e-12);
  const double log_sig_l = log(sig_l);
  const double sqrt_beta = sqrt(acq_beta > 0.0 ? acq_beta : 0.0);
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
            const double tt = (xc[k * TC + c] - s.Zt[k * cap + j]) * invl[k];
            r2 += tt * tt;
          }
        v = os * exp(-0.5 * r2);
      }
      tile[j * TC + c] = v;
    }
    tm.sync();
    int n_slots;
    var_tile_partials(tm, s, w, n, tile, part, slots, packed, n_slots);
    tm.sync();
    for (int c = tm.rank(); c < cnt; c += tm.size()) {
      double sa = 0.0, sb = 0.0, mu = 0.0;
      for (int ww = 0; ww < n_slots; ++ww) {
        sa += part[ww * TC + c];
        sb += part[(slots + ww) * TC + c];
      }
      const int nmu = nw < 8 ? nw : 8;
      for (int ww = 0; ww < nmu; ++ww) mu += part[(2 * slots + ww) * TC + c];
      const int gi = base + c;
      const double mean = cst + mu;
      const double var = os + kVarJitter + (sb - sa);
      if (mean_out) mean_out[gi] = mean;
      if (var_out) var_out[gi] = var;
      part[c] = mean;              // slot 0 of this column is dead now (only this thread read it): hand (mean, var) over
      part[slots * TC + c] = var;
    }
    tm.sync();
    if (lik == kLikGaussian) {
      for (int c = tm.rank(); c < cnt; c += tm.size()) {
        const int gi = base + c;
        const double a = acq_value(acq_kind, acq_beta, part[c], part[slots * TC + c] + noise, best_f);
        if (acq_out) acq_out[gi] = a;
        if ((!mask || mask[gi]) && acq_better(a, gi, run_v, run_i)) { run_v = a; run_i = gi; }
      }
    } else {
      // Monte-Carlo mean of LogEI over S latent samples; one warp per candidate, lanes split the samples
      for (int c = tm.warp(); c < cnt; c += nw) {
        const int gi = base + c;
        if (mask && !mask[gi] && !acq_out) continue;
        const double mean = part[c], sd = sqrt(part[slots * TC + c]);
        double a = 0.0;
        if (eps) {
          const double* e = eps + (size_t)gi * S;
          for (int q = tm.lane(); q < S; q += TC) a += acq_sample(acq_kind, sqrt_beta, mean + sd * e[q], best_f, sig_l);
        } else if (acq_kind == kAcqLogEI) {
          // u = ((mean + sd z) - best_f) / sigma_l as one FMA per sample: the division (and two more operations) per
          // sample are hoisted to the candidate (in-kernel stream only; the explicit-eps parity path keeps the
          // reference's operation order)
          const double inv = 1.0 / sig_l, ua = (mean - best_f) * inv, ub = sd * inv;
          for (int q = tm.lane(); 2 * q < S; q += TC) {
            double z0, z1;
            philox_normal2(seed, stream, (uint64_t)gi * (uint64_t)((S + 1) / 2) + q, z0, z1);
            a += log_h_mc(fma(ub, z0, ua));
            if (2 * q + 1 < S) a += log_h_mc(fma(ub, z1, ua));
          }
        } else {
          for (int q = tm.lane(); 2 * q < S; q += TC) {
            double z0, z1;
            philox_normal2(seed, stream, (uint64_t)gi * (uint64_t)((S + 1) / 2) + q, z0, z1);
            a += acq_sample(acq_kind, sqrt_beta, mean + sd * z0, best_f, sig_l);
            if (2 * q + 1 < S) a += acq_sample(acq_kind, sqrt_beta, mean + sd * z1, best_f, sig_l);
          }
        }
        a = tm.warp_sum(a) / S + (acq_kind == kAcqLogEI ? log_sig_l : 0.0);
        if (tm.lane() == 0) {
          if (acq_out) acq_out[gi] = a;
          if ((!mask || mask[gi]) && acq_better(a, gi, run_v, run_i)) { run_v = a; run_i = gi; }
        }
      }
    }
  }
  tm.argmax(run_v, run_i);
  *best_idx = run_i;
  *best_acq = run_v;
}

}  // namespace stp
