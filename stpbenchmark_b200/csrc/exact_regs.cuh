// Register-blocked variant of the exact-GP closure (device only).
//
// Same mathematics as exact_mll_fg (exact.cuh) -- bordered symmetric sweep of [K + s2 I, y - c; (y - c)^T, 0],
// then the gradient contraction 1/2 tr((alpha alpha^T - K^-1) dK) -- but the matrix never touches shared
// memory: thread t owns one BS x BS block (BS = STP_REG_BS, bi >= bj) of the lower triangle in REGISTERS for the whole sweep.
// Per pivot a thread loads the 4 + 4 pivot-column entries it needs (two LDS.128 each), issues 16 DFMA, and
// the owners of row / column k+1 publish the next pivot column; one barrier per pivot.  ~1.5 instructions per
// matrix element per pivot instead of ~10 for the shared-memory loop, and no bank traffic for the matrix.
// Needs blockDim.x >= number of 4 x 4 blocks = mb (mb + 1) / 2, mb = ceil((n + 1) / 4)  (276 for n = 89).
#pragma once
#include "exact.cuh"

#if defined(__CUDACC__)
namespace stp {

#ifndef STP_REG_BS
#define STP_REG_BS 6            // edge of the square register block owned by one thread
#endif
__host__ __device__ __forceinline__ int reg_blocks_for(int m) {
  const int mb = (m + STP_REG_BS - 1) / STP_REG_BS;
  return mb * (mb + 1) / 2;
}

// K + noise I, bordered by y - c, evaluated at (i, j); padding rows / columns are zero.
__device__ __forceinline__ double bordered_entry(const ExactSmem& s, int n, int d, const double* invl, double os,
                                                 double noise, double cst, int i, int j) {
  if (i > n || j > n) return 0.0;
  if (i == n) return (j == n) ? 0.0 : s.y[j] - cst;
  if (j == n) return s.y[i] - cst;
  if (i == j) return os + noise;
  return ard_rbf_tt(s.Xt, s.cap, d, invl, i, j, os);
}

template <int BS>
__device__ __forceinline__ int exact_mll_fg_regs(const CtaTeam& tm, const ExactSmem& s, int n, int d,
                                                 const double* theta, const ExactPrior& pr, double* f_out, double* g) {
  const int cap = s.cap;
  const double noise = theta[0], cst = theta[1], os = theta[2];
  double* invl = s.small + 32;
  tm.sync();
  for (int k = tm.rank(); k < d; k += tm.size()) invl[k] = 1.0 / theta[3 + k];
  int bad = !(noise > 0.0) || !(os > 0.0) || !(cst == cst);
  for (int k = 0; k < d; ++k) bad |= !(theta[3 + k] > 0.0);
  if (bad) { *f_out = NAN; return 2; }
  tm.sync();
  const int m = n + 1, mb = (m + BS - 1) / BS, nblk = mb * (mb + 1) / 2, mp = mb * BS;
  const int t = tm.rank();
  const bool live = t < nblk;
  // block coordinates: largest bi with bi (bi + 1) / 2 <= t
  int bi = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
  while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
  while (bi * (bi + 1) / 2 > t) --bi;
  const int bj = t - bi * (bi + 1) / 2;
  const int i0 = bi * BS, j0 = bj * BS;
  double A[BS][BS];
  if (live) {
#pragma unroll
    for (int a = 0; a < BS; ++a)
#pragma unroll
      for (int b = 0; b < BS; ++b) A[a][b] = bordered_entry(s, n, d, invl, os, noise, cst, i0 + a, j0 + b);
  }
  double* cur = s.c0;   // mp <= cap + 4 doubles each: c0 / c1 / piv are contiguous, see exact_smem_carve
  double* nxt = s.c1;
  if (live && bj == 0) {
#pragma unroll
    for (int a = 0; a < BS; ++a) cur[i0 + a] = A[a][0];
  }
  tm.sync();
  const int kb_i = i0, kb_j = j0;
  for (int k = 0; k < n; ++k) {
    const double p = cur[k];
    if (!(p > 0.0) || !(p <= 1.79e308)) { *f_out = NAN; return 1; }
    const double invp = 1.0 / p;
    if (t == 0) s.piv[k] = p;
    if (live) {
      double ci[BS], cj[BS], ti[BS];
#pragma unroll
      for (int a = 0; a < BS; ++a) { ci[a] = cur[kb_i + a]; cj[a] = cur[kb_j + a]; ti[a] = ci[a] * invp; }
#pragma unroll
      for (int a = 0; a < BS; ++a)
#pragma unroll
        for (int b = 0; b < BS; ++b) A[a][b] = fma(-ti[a], cj[b], A[a][b]);
      const int ka = k - i0, kc = k - j0;   // position of the pivot inside this block's rows / columns
      if ((unsigned)ka < (unsigned)BS) {              // pivot row: a_kj / p, diagonal -1 / p
#pragma unroll
        for (int a = 0; a < BS; ++a)
          if (a == ka) {
#pragma unroll
            for (int b = 0; b < BS; ++b) A[a][b] = (b == kc) ? -invp : cj[b] * invp;
          }
      }
      if ((unsigned)kc < (unsigned)BS) {              // pivot column: a_ik / p
#pragma unroll
        for (int b = 0; b < BS; ++b)
          if (b == kc) {
#pragma unroll
            for (int a = 0; a < BS; ++a)
              if (a != ka) A[a][b] = ti[a];
          }
      }
      const int na = k + 1 - i0, nc = k + 1 - j0;   // publish column k + 1 (rows >= k + 1) and row k + 1 (cols <= k + 1)
      if ((unsigned)nc < (unsigned)BS) {
#pragma unroll
        for (int b = 0; b < BS; ++b)
          if (b == nc) {
#pragma unroll
            for (int a = 0; a < BS; ++a)
              if (i0 + a >= k + 1) nxt[i0 + a] = A[a][b];
          }
      }
      if ((unsigned)na < (unsigned)BS) {
#pragma unroll
        for (int a = 0; a < BS; ++a)
          if (a == na) {
#pragma unroll
            for (int b = 0; b < BS; ++b)
              if (j0 + b <= k + 1) nxt[j0 + b] = A[a][b];
          }
      }
    }
    tm.sync();
    double* tt = cur; cur = nxt; nxt = tt;
  }
  // border row -> alpha (shared), corner -> -quad
  const int bn = n / BS, an = n - bn * BS;
  if (live && bi == bn) {
#pragma unroll
    for (int a = 0; a < BS; ++a)
      if (a == an) {
#pragma unroll
        for (int b = 0; b < BS; ++b) {
          const int j = j0 + b;
          if (j < n) s.alpha[j] = A[a][b];
          else if (j == n) s.small[40] = -A[a][b];
        }
      }
  }
  tm.sync();
  const double quad = s.small[40];
  double acc[kMaxD + 4];
#pragma unroll
  for (int q = 0; q < kMaxD + 4; ++q) acc[q] = 0.0;
  for (int i = tm.rank(); i < n; i += tm.size()) {
    acc[0] += log(s.piv[i]);
    acc[2] += s.alpha[i];
  }
  if (live) {
#pragma unroll
    for (int a = 0; a < BS; ++a) {
      const int i = i0 + a;
      if (i < n) {
        const double ai = s.alpha[i];
#pragma unroll
        for (int b = 0; b < BS; ++b) {
          const int j = j0 + b;
          if (j < i) {
            double r2 = 0.0, dx2[kMaxD];
#pragma unroll
            for (int k = 0; k < kMaxD; ++k)
              if (k < d) {
                const double dx = s.Xt[k * cap + i] - s.Xt[k * cap + j];
                dx2[k] = dx * dx;
                r2 += dx2[k] * invl[k] * invl[k];
              }
            const double tw = (ai * s.alpha[j] + A[a][b]) * (os * exp(-0.5 * r2));
            acc[3] += tw;
#pragma unroll
            for (int k = 0; k < kMaxD; ++k)
              if (k < d) acc[4 + k] += tw * dx2[k];
          } else if (j == i) {
            acc[1] += ai * ai + A[a][b];
          }
        }
      }
    }
  }
  tm.sum_vec(acc, 4 + d);
  const double logN = -0.5 * (quad + acc[0] + n * kLog2Pi);
  double total = logN + lognormal_logpdf(noise, pr.noise_loc, pr.noise_scale) +
                 lognormal_logpdf(os, pr.os_loc, pr.os_scale);
  for (int k = 0; k < d; ++k) total += lognormal_logpdf(theta[3 + k], pr.ls_loc, pr.ls_scale);
  const double scale = -1.0 / n;
  const double f = total * scale;
  tm.sync();
  if (tm.rank() == 0) {
    g[0] = scale * (0.5 * acc[1] + lognormal_dlogpdf(noise, pr.noise_loc, pr.noise_scale));
    g[1] = scale * acc[2];
    g[2] = scale * (acc[3] / os + 0.5 * acc[1] + lognormal_dlogpdf(os, pr.os_loc, pr.os_scale));
  }
  for (int k = tm.rank(); k < d; k += tm.size()) {
    const double l = theta[3 + k];
    g[3 + k] = scale * (acc[4 + k] / (l * l * l) + lognormal_dlogpdf(l, pr.ls_loc, pr.ls_scale));
  }
  tm.sync();
  *f_out = f;
  int st = !(f == f) || !(fabs(f) <= 1.79e308);
  for (int k = 0; k < d + 3; ++k) st |= !(g[k] == g[k]);
  (void)mp;
  return st ? 2 : 0;
}

// Closure dispatch: register-blocked when the CTA has one thread per 4 x 4 block, shared-memory sweep otherwise.
__device__ __forceinline__ int exact_closure(const CtaTeam& tm, const ExactSmem& s, int n, int d, const double* theta,
                                             const ExactPrior& pr, double* f_out, double* g) {
#if defined(STP_EXACT_REGS) && STP_EXACT_REGS
  if (reg_blocks_for(n + 1) <= tm.size()) return exact_mll_fg_regs<STP_REG_BS>(tm, s, n, d, theta, pr, f_out, g);
#endif
  return exact_mll_fg(tm, s, n, d, theta, pr, f_out, g);
}

}  // namespace stp
#endif

namespace stp {
STP_HD int exact_closure(const SoloTeam& tm, const ExactSmem& s, int n, int d, const double* theta,
                         const ExactPrior& pr, double* f_out, double* g) {
  return exact_mll_fg(tm, s, n, d, theta, pr, f_out, g);
}
}  // namespace stp
