// Device-resident L-BFGS-B (Byrd, Lu, Nocedal, Zhu 1995; v3.0 with the Morales-Nocedal 2011
// subspace projection) for the tiny box-constrained problems of the exact-GP fit
// (d + 3 <= 11 unknowns, history m = 10) -- replaces scipy's L-BFGS-B, which botorch's
// fit_gpytorch_mll drives at optim.py:31-32.
//
// Same algorithm, different linear algebra: with <= 11 unknowns the limited-memory matrix
//   B = theta I - W M W^T            (compact form of the last `col` BFGS pairs)
// is formed densely by replaying those pairs on theta I (Byrd-Nocedal-Schnabel 1994, Thm 2.3:
// the two are identical), so the generalised Cauchy point, the subspace minimiser
// (B_FF d = -r_F by a dense Cholesky) and every test are the same mathematical quantities
// scipy computes through the 2m x 2m middle matrix.  Control flow, constants and
// termination tests follow the published routine (mainlb / cauchy / subsm / lnsrlb /
// dcsrch / dcstep) and scipy's wrapper (_minimize_lbfgsb): m = 10, factr = ftol / eps,
// pgtol, maxls = 20, line-search ftol = 1e-3, gtol = 0.9, xtol = 0.1, maxiter checked at NEW_X.
//
// Ask/tell interface, run by ONE WARP of the CTA (the objective between two calls is evaluated by the
// whole CTA):
//     lbfgsb_init(L, S, ...);                       // S.x holds the first point to evaluate
//     while (lbfgsb_advance(L, S, f, g)) { f, g = objective(S.x); }
//     -> S.x, S.f, S.status (0 converged, 1 iteration/evaluation cap, 2 abnormal line search,
//        3 non-finite objective), S.iter, S.nfev
// `L` is a lane group (team.cuh): WarpLanes on the device -- lane i owns variable i / row i of the small
// matrices, dot products are shuffle reductions, the scalar bookkeeping (line search) is done by lane 0 -- or
// SoloLanes (one thread, every loop serial) in the host emulation and the C++ unit tests.  The state lives
// in shared memory; every routine leaves it consistent for all lanes (ends with L.sync()) and returns a
// lane-uniform value.
//
// Code size is a first-class constraint here.  One warp runs this branchy state machine between two closures while
// the other warps of the CTA wait, and the SM's instruction caches are small (L0 ~6 KB, L1.5 32 KB per SM, shared
// GPC-level cache behind them): with every routine force-inlined at each call site and every `for (i = lane; i < n;
// i += 32)` unrolled, the optimiser was 270 KB of SASS inside a 505 KB kernel, ran at ~32 cycles per instruction
// (117 k cycles per step, 50 - 75 % of a CTA's time) and the GPC instruction cache sat at 40 % of its peak request
// rate (profiles/r02_summary.md).  Hence: every routine exists ONCE (STP_HD_NOINLINE), lane-owned loops are
// STP_OWNED (at most one trip per lane), serial loops are STP_ROLLED, divisions / square roots off the dependent
// chains go through one out-of-line helper, reductions through warp_sum16 / warp_max16.
#pragma once
#include "exact.cuh"
#include "team.cuh"

namespace stp {

constexpr int kMaxP = kMaxD + 3;
constexpr int kHist = 10;
constexpr double kEpsMch = 2.220446049250313e-16;

struct LbfgsbOpts {
  int m;          // history length (<= kHist)
  int maxiter;
  int maxfun;
  int maxls;
  double factr;   // ftol / eps
  double pgtol;
};

struct LbfgsbState {
  int np, m, maxiter, maxfun, maxls;
  double factr, pgtol;
  double x[kMaxP], g[kMaxP], lb[kMaxP], ub[kMaxP];
  int nbd[kMaxP];                 // 0 free, 1 lower, 2 both, 3 upper
  int iwhere[kMaxP];
  double f, fold, sbgnrm, theta;
  double t[kMaxP], r[kMaxP], d[kMaxP], z[kMaxP], xp[kMaxP];
  double ws[kHist][kMaxP], wy[kHist][kMaxP];
  double ys[kHist], iys[kHist];   // s.y of every stored pair and its reciprocal (fixed once the pair is stored)
  double B[kMaxP][kMaxP];
  int col, head, itail, iupdat, updatd;
  int iter, nfev, nskip, iback, ifun, info;
  int cnstnd, boxed;
  // line search (lnsrlb + dcsrch) state
  double stp, stpmx, dnorm, dtd, gd, gdold;
  int ls_task;                    // 0 START, 1 FG, 2 CONVERGENCE, 3 WARNING, 4 ERROR
  int brackt, stage;
  double ginit, gtest, gx, gy, finit, fx, fy, stx, sty, stmin, stmax, width, width1;
  int phase;                      // 0 waiting for f(x0); 1 inside a line search
  int lsfail;                     // lane 0 -> all lanes: the line search could not start
  // scratch of the routines: kept in the (shared-memory) state on purpose -- with ~225 KB of the SM's
  // 228 KB carved out as shared memory there is next to no L1 left, and thread-local arrays would turn every
  // access into an L2 round trip
  double sc_a[kMaxP], sc_b[kMaxP], sc_c[kMaxP], sc_d[kMaxP], sc_e[kMaxP], sc_C[kMaxP][kMaxP];
  int sc_i[kMaxP], sc_j[kMaxP];
  int status;
};

// ------------------------------ small dense helpers -----------------------------------
// IEEE division / square root, one copy (the inlined sequences are ~30 instructions each plus a slow path).
STP_HD_NOINLINE double lb_div(double a, double b) { return a / b; }
STP_HD_NOINLINE double lb_sqrt(double a) { return sqrt(a); }

template <class LN>
STP_HD double lb_dot(const LN& L, const double* a, const double* b, int n) {
  double s = 0.0;
  STP_OWNED(L, i, n) s += a[i] * b[i];
  return L.sum(s);
}

template <class LN>
STP_HD void lb_projgr(const LN& L, LbfgsbState& S) {
  double m = 0.0;
  STP_OWNED(L, i, S.np) {
    double gi = S.g[i];
    const int nb = S.nbd[i];
    if (nb != 0) {
      if (gi < 0.0) { if (nb >= 2) gi = fmax(S.x[i] - S.ub[i], gi); }
      else { if (nb <= 2) gi = fmin(S.x[i] - S.lb[i], gi); }
    }
    m = fmax(m, fabs(gi));
  }
  m = L.max(m);
  L.sync();
  if (L.lane() == 0) S.sbgnrm = m;
  L.sync();
}

template <class LN>
STP_HD void lb_reset_memory(const LN& L, LbfgsbState& S) {
  L.sync();
  if (L.lane() == 0) { S.col = 0; S.head = 0; S.itail = -1; S.theta = 1.0; S.iupdat = 0; S.updatd = 0; }
  L.sync();
}

// B = theta I replayed through the stored (s, y) pairs, oldest first.  Returns 1 on breakdown.
STP_HD_NOINLINE int lb_build_B(const SoloLanes&, LbfgsbState& S) {
  const int n = S.np;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) S.B[i][j] = (i == j) ? S.theta : 0.0;
  for (int q = 0; q < S.col; ++q) {
    const int p = (S.head + q) % S.m;
    const double* s = S.ws[p];
    const double* y = S.wy[p];
    double* Bs = S.sc_a;
    double sBs = 0.0;
    for (int i = 0; i < n; ++i) {
      double a = 0.0;
      for (int j = 0; j < n; ++j) a += S.B[i][j] * s[j];
      Bs[i] = a;
    }
    for (int i = 0; i < n; ++i) sBs += s[i] * Bs[i];
    if (!(sBs > 0.0) || !(S.ys[p] > 0.0)) return 1;
    const double ia = 1.0 / sBs, ib = S.iys[p];
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) S.B[i][j] += y[i] * y[j] * ib - Bs[i] * Bs[j] * ia;
  }
  return 0;
}
#if defined(__CUDACC__)
// Warp version: lane r keeps row r of B in registers.  Per stored pair: Bs_r = B[r,:] . s (two DFMA chains), Bs is
// exchanged through shared memory (one store + warp barrier + broadcast loads: the 22 shuffles and the 5-level
// butterfly this replaces were 20 % of the optimiser's time), every lane forms s . Bs itself, rank-2 update.
static __device__ __noinline__ int lb_build_B(const WarpLanes& L, LbfgsbState& S) {
  const int n = S.np, r = L.lane();
  const bool live = r < n;
  double Brow[kMaxP];
#pragma unroll
  for (int c = 0; c < kMaxP; ++c) Brow[c] = (live && c == r) ? S.theta : 0.0;
  int broke = 0;
  const int col = S.col, head = S.head, mm = S.m;
  double* Bsv = S.sc_a;
  STP_ROLLED for (int q = 0; q < col; ++q) {
    int p = head + q;
    if (p >= mm) p -= mm;
    const double* s = S.ws[p];
    const double* y = S.wy[p];
    const double ysp = S.ys[p], ib = S.iys[p];
    double sv[kMaxP];
    double Bs0 = 0.0, Bs1 = 0.0;   // two chains: the dot product is latency-bound
#pragma unroll
    for (int c = 0; c < kMaxP; ++c) sv[c] = (c < n) ? s[c] : 0.0;
#pragma unroll
    for (int c = 0; c < kMaxP; c += 2) {
      Bs0 = fma(Brow[c], sv[c], Bs0);
      if (c + 1 < kMaxP) Bs1 = fma(Brow[(c + 1 < kMaxP) ? c + 1 : c], sv[(c + 1 < kMaxP) ? c + 1 : c], Bs1);
    }
    const double Bs = Bs0 + Bs1;
    L.sync();                       // the previous pair's readers are done with Bsv
    if (r < kMaxP) Bsv[r] = live ? Bs : 0.0;
    L.sync();
    double bsv[kMaxP];
    double a0 = 0.0, a1 = 0.0;
#pragma unroll
    for (int c = 0; c < kMaxP; ++c) bsv[c] = Bsv[c];
#pragma unroll
    for (int c = 0; c < kMaxP; c += 2) {
      a0 = fma(sv[c], bsv[c], a0);
      if (c + 1 < kMaxP) a1 = fma(sv[(c + 1 < kMaxP) ? c + 1 : c], bsv[(c + 1 < kMaxP) ? c + 1 : c], a1);
    }
    const double a = a0 + a1;
    if (!(a > 0.0) || !(ysp > 0.0)) { broke = 1; break; }
    const double ia = rcp_pos(a);
    const double ya = (live ? y[r] : 0.0) * ib, ba = Bs * ia;
#pragma unroll
    for (int c = 0; c < kMaxP; ++c)
      if (c < n) Brow[c] += ya * y[c] - ba * bsv[c];
  }
  L.sync();
  if (live) {
#pragma unroll
    for (int c = 0; c < kMaxP; ++c)
      if (c < n) S.B[r][c] = Brow[c];
  }
  L.sync();
  return broke;
}
#endif

// row i of B times a vector (serial over the <= 11 columns)
STP_HD double lb_Brow_dot(const LbfgsbState& S, int i, const double* v, int n) {
  double a = 0.0;
  STP_ROLLED for (int j = 0; j < n; ++j) a += S.B[i][j] * v[j];
  return a;
}

// ------------------------------ generalised Cauchy point -------------------------------
// On exit S.z = x^cp and S.iwhere marks the variables fixed at a bound.
template <class LN>
STP_HD_NOINLINE void lb_cauchy(const LN& L, LbfgsbState& S) {
  const int n = S.np, l0 = L.lane();
  double* xcp = S.z;
  double* dvec = S.sc_a; double* tbrk = S.sc_b; double* tall = S.sc_c;
  int* order = S.sc_i; int* hasb = S.sc_j;
  if (S.sbgnrm <= 0.0) {
    STP_OWNED(L, i, n) xcp[i] = S.x[i];
    L.sync();
    return;
  }
  int nb_own = 0;
  STP_OWNED(L, i, n) {
    const double neggi = -S.g[i];
    const int nbd = S.nbd[i];
    double tl = 0.0, tu = 0.0;
    int iw = S.iwhere[i];
    if (iw != 3 && iw != -1) {
      if (nbd <= 2) tl = S.x[i] - S.lb[i];
      if (nbd >= 2) tu = S.ub[i] - S.x[i];
      const int xlower = nbd <= 2 && tl <= 0.0;
      const int xupper = nbd >= 2 && tu <= 0.0;
      iw = 0;
      if (xlower) { if (neggi <= 0.0) iw = 1; }
      else if (xupper) { if (neggi >= 0.0) iw = 2; }
      else { if (fabs(neggi) <= 0.0) iw = -3; }
      S.iwhere[i] = iw;
    }
    xcp[i] = S.x[i];
    int hb = 0;
    double tb = 0.0, dv = 0.0;
    if (iw == 0 || iw == -1) {
      dv = neggi;
      if (nbd <= 2 && nbd != 0 && neggi < 0.0) { tb = lb_div(tl, -neggi); hb = 1; }
      else if (nbd >= 2 && neggi > 0.0) { tb = lb_div(tu, neggi); hb = 1; }
    }
    dvec[i] = dv; tall[i] = tb; hasb[i] = hb;
    nb_own += hb;
  }
  const int nbreak = L.isum(nb_own);
  L.sync();
  // breakpoints in ascending order (stable: ties keep the variable order, like the insertion sort it replaces)
  if (nbreak > 0) {
    STP_OWNED(L, i, n)
      if (hasb[i]) {
        const double ti = tall[i];
        int rk = 0;
        STP_ROLLED for (int j = 0; j < n; ++j) rk += (hasb[j] && (tall[j] < ti || (tall[j] == ti && j < i))) ? 1 : 0;
        tbrk[rk] = ti; order[rk] = i;
      }
  }
  // derivatives of the quadratic model along the projected-gradient path
  double f1p = 0.0, f2p = 0.0;
  STP_OWNED(L, i, n) {
    const double di = dvec[i], bd = lb_Brow_dot(S, i, dvec, n);
    f1p -= di * di;
    f2p += di * bd;
  }
  double f1 = L.sum(f1p), f2 = L.sum(f2p);
  L.sync();
  const double f2_org = -S.theta * f1;  // the col = 0 value, as in the published routine
  if (!(f2 > 0.0)) return;               // d = 0: x is its own Cauchy point
  double dtm = lb_div(-f1, f2);
  double tsum = 0.0, tj = 0.0;
  int nfixed_all = 0;
  STP_ROLLED for (int q = 0; q < nbreak; ++q) {
    const double tj0 = tj;
    tj = tbrk[q];
    const double dt = tj - tj0;
    if (dtm < dt) break;
    const int ibp = order[q];
    tsum += dt;
    const double dibp = dvec[ibp];
    L.sync();   // every lane holds dvec[ibp] before lane 0 clears it
    if (l0 == 0) {
      dvec[ibp] = 0.0;
      if (dibp > 0.0) { xcp[ibp] = S.ub[ibp]; S.iwhere[ibp] = 2; }
      else { xcp[ibp] = S.lb[ibp]; S.iwhere[ibp] = 1; }
    }
    L.sync();
    if (q == nbreak - 1 && nbreak == n) { dtm = dt; nfixed_all = 1; break; }
    // z = x(tsum) - x : fixed variables sit on their bounds, the others moved tsum * d
    f1p = 0.0; f2p = 0.0;
    STP_OWNED(L, i, n) {
      const double di = dvec[i];
      const double zi = (di != 0.0 || S.iwhere[i] <= 0) ? tsum * di : xcp[i] - S.x[i];
      const double bd = lb_Brow_dot(S, i, dvec, n);
      f1p += di * S.g[i];
      f2p += di * bd;
      f1p += zi * bd;
    }
    f1 = L.sum(f1p); f2 = L.sum(f2p);
    f2 = fmax(kEpsMch * f2_org, f2);
    dtm = lb_div(-f1, f2);
  }
  if (!nfixed_all) {
    if (dtm < 0.0) dtm = 0.0;
    tsum += dtm;
  }
  STP_OWNED(L, i, n)
    if (dvec[i] != 0.0) xcp[i] = S.x[i] + tsum * dvec[i];
  L.sync();
}

// ------------------------------ subspace minimisation ----------------------------------
// Minimise the quadratic model over the variables free at x^cp, then project (v3.0).
// On entry S.z = x^cp; on exit S.z = the subspace minimiser.  Returns 1 if B_FF is not PD.
template <class LN>
STP_HD_NOINLINE int lb_subsm(const LN& L, LbfgsbState& S) {
  const int n = S.np, l0 = L.lane();
  int* ind = S.sc_j;
  int nf_own = 0;
  STP_OWNED(L, i, n) nf_own += (S.iwhere[i] <= 0) ? 1 : 0;
  const int nf = L.isum(nf_own);
  if (nf == 0 || S.col == 0) return 0;
  L.sync();
  STP_OWNED(L, i, n)
    if (S.iwhere[i] <= 0) {
      int pos = 0;
      STP_ROLLED for (int j = 0; j < i; ++j) pos += (S.iwhere[j] <= 0) ? 1 : 0;
      ind[pos] = i;
    }
  // z - x, once (the rows of B_FF need it nf times)
  double* dz = S.sc_d;
  STP_OWNED(L, i, n) dz[i] = S.z[i] - S.x[i];
  L.sync();
  double* rr = S.sc_a; double* dd = S.sc_b; double* idiag = S.sc_c; double* sol = S.sc_e;
  double (*C)[kMaxP] = S.sc_C;
  STP_OWNED(L, a, nf) {
    const int i = ind[a];
    rr[a] = -(S.g[i] + lb_Brow_dot(S, i, dz, n));
    STP_ROLLED for (int b = 0; b <= a; ++b) C[a][b] = S.B[i][ind[b]];
  }
  L.sync();
  // Cholesky of B_FF, column by column; the rows of a column in parallel (C[j][j] keeps l_jj^2, 1 / l_jj in idiag)
  STP_ROLLED for (int j = 0; j < nf; ++j) {
    STP_OWNED(L, i, nf)
      if (i >= j) {
        double v = C[i][j];
        STP_ROLLED for (int k = 0; k < j; ++k) v -= C[i][k] * C[j][k];
        C[i][j] = v;
      }
    L.sync();
    const double s = C[j][j];
    if (!(s > 0.0)) return 1;
    const double il = rsqrt_pos(s);
    if (l0 == 0) idiag[j] = il;
    STP_OWNED(L, i, nf)
      if (i > j) C[i][j] *= il;
    L.sync();
  }
  STP_ROLLED for (int j = 0; j < nf; ++j) {          // forward substitution, column-oriented
    const double dj = rr[j] * idiag[j];
    if (l0 == 0) dd[j] = dj;
    STP_OWNED(L, i, nf)
      if (i > j) rr[i] -= C[i][j] * dj;
    L.sync();
  }
  STP_ROLLED for (int j = nf - 1; j >= 0; --j) {     // back substitution
    const double dj = dd[j] * idiag[j];
    if (l0 == 0) sol[j] = dj;
    STP_OWNED(L, i, j) dd[i] -= C[j][i] * dj;
    L.sync();
  }
  // projection of x^cp + d onto the box
  STP_OWNED(L, i, n) S.xp[i] = S.z[i];
  L.sync();
  int iw = 0;
  STP_OWNED(L, a, nf) {
    const int k = ind[a];
    const int nbd = S.nbd[k];
    const double zk0 = S.z[k] + sol[a];
    double zk = zk0;
    if (nbd != 0) {
      const double lo = S.lb[k], hi = S.ub[k];
      if (nbd <= 2) zk = fmax(lo, zk);
      if (nbd >= 2) zk = fmin(hi, zk);
      if ((nbd <= 2 && zk == lo) || (nbd >= 2 && zk == hi)) iw = 1;
    }
    S.z[k] = zk;
  }
  const int iword = L.any(iw);
  L.sync();
  if (!iword) return 0;
  double ddpp = 0.0;
  STP_OWNED(L, i, n) ddpp += (S.z[i] - S.x[i]) * S.g[i];
  const double ddp = L.sum(ddpp);
  L.sync();
  if (ddp > 0.0) {
    if (l0 == 0) {   // rare path, kept serial: the running `temp1` makes the published loop order-dependent
      STP_ROLLED for (int i = 0; i < n; ++i) S.z[i] = S.xp[i];
      double alpha = 1.0, temp1 = 1.0;
      int ibd = -1;
      STP_ROLLED for (int a = 0; a < nf; ++a) {
        const int k = ind[a];
        const double dk = sol[a];
        if (S.nbd[k] != 0) {
          if (dk < 0.0 && S.nbd[k] <= 2) {
            const double t2 = S.lb[k] - S.z[k];
            if (t2 >= 0.0) temp1 = 0.0; else if (dk * alpha < t2) temp1 = lb_div(t2, dk);
          } else if (dk > 0.0 && S.nbd[k] >= 2) {
            const double t2 = S.ub[k] - S.z[k];
            if (t2 <= 0.0) temp1 = 0.0; else if (dk * alpha > t2) temp1 = lb_div(t2, dk);
          }
          if (temp1 < alpha) { alpha = temp1; ibd = a; }
        }
      }
      if (alpha < 1.0 && ibd >= 0) {
        const double dk = sol[ibd];
        const int k = ind[ibd];
        if (dk > 0.0) { S.z[k] = S.ub[k]; sol[ibd] = 0.0; }
        else if (dk < 0.0) { S.z[k] = S.lb[k]; sol[ibd] = 0.0; }
      }
      STP_ROLLED for (int a = 0; a < nf; ++a) S.z[ind[a]] += alpha * sol[a];
    }
    L.sync();
  }
  return 0;
}

// ------------------------------ More-Thuente line search -------------------------------
// cubic-interpolation helper shared by the four cases of dcstep: gamma = s sqrt(max(0, (th/s)^2 - (a/s)(b/s)))
STP_HD_NOINLINE double lb_gamma(double th, double a, double b, int clamp) {
  const double s = fmax(fabs(th), fmax(fabs(a), fabs(b)));
  const double ts = lb_div(th, s);
  double r = ts * ts - lb_div(a, s) * lb_div(b, s);
  if (clamp) r = fmax(0.0, r);
  return s * lb_sqrt(r);
}

STP_HD_NOINLINE void lb_dcstep(double& stx, double& fx, double& dx, double& sty, double& fy, double& dy, double& stp,
                               double fp, double dp, int& brackt, double stpmin, double stpmax) {
  const double sgnd = dp * lb_div(dx, fabs(dx));
  double stpf, stpc, stpq, th, gamma, p, q, r;
  if (fp > fx) {
    th = lb_div(3.0 * (fx - fp), stp - stx) + dx + dp;
    gamma = lb_gamma(th, dx, dp, 0);
    if (stp < stx) gamma = -gamma;
    p = (gamma - dx) + th; q = ((gamma - dx) + gamma) + dp; r = lb_div(p, q);
    stpc = stx + r * (stp - stx);
    stpq = stx + (lb_div(dx, lb_div(fx - fp, stp - stx) + dx) / 2.0) * (stp - stx);
    stpf = (fabs(stpc - stx) < fabs(stpq - stx)) ? stpc : stpc + (stpq - stpc) / 2.0;
    brackt = 1;
  } else if (sgnd < 0.0) {
    th = lb_div(3.0 * (fx - fp), stp - stx) + dx + dp;
    gamma = lb_gamma(th, dx, dp, 0);
    if (stp > stx) gamma = -gamma;
    p = (gamma - dp) + th; q = ((gamma - dp) + gamma) + dx; r = lb_div(p, q);
    stpc = stp + r * (stx - stp);
    stpq = stp + lb_div(dp, dp - dx) * (stx - stp);
    stpf = (fabs(stpc - stp) > fabs(stpq - stp)) ? stpc : stpq;
    brackt = 1;
  } else if (fabs(dp) < fabs(dx)) {
    th = lb_div(3.0 * (fx - fp), stp - stx) + dx + dp;
    gamma = lb_gamma(th, dx, dp, 1);
    if (stp > stx) gamma = -gamma;
    p = (gamma - dp) + th; q = (gamma + (dx - dp)) + gamma; r = lb_div(p, q);
    if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
    else if (stp > stx) stpc = stpmax;
    else stpc = stpmin;
    stpq = stp + lb_div(dp, dp - dx) * (stx - stp);
    if (brackt) {
      stpf = (fabs(stpc - stp) < fabs(stpq - stp)) ? stpc : stpq;
      if (stp > stx) stpf = fmin(stp + 0.66 * (sty - stp), stpf);
      else stpf = fmax(stp + 0.66 * (sty - stp), stpf);
    } else {
      stpf = (fabs(stpc - stp) > fabs(stpq - stp)) ? stpc : stpq;
      stpf = fmin(stpmax, stpf);
      stpf = fmax(stpmin, stpf);
    }
  } else {
    if (brackt) {
      th = lb_div(3.0 * (fp - fy), sty - stp) + dy + dp;
      gamma = lb_gamma(th, dy, dp, 0);
      if (stp > sty) gamma = -gamma;
      p = (gamma - dp) + th; q = ((gamma - dp) + gamma) + dy; r = lb_div(p, q);
      stpc = stp + r * (sty - stp);
      stpf = stpc;
    } else if (stp > stx) stpf = stpmax;
    else stpf = stpmin;
  }
  if (fp > fx) { sty = stp; fy = fp; dy = dp; }
  else {
    if (sgnd < 0.0) { sty = stx; fy = fx; dy = dx; }
    stx = stp; fx = fp; dx = dp;
  }
  stp = stpf;
}

// One call of dcsrch with (f, g) at S.stp; updates S.stp and S.ls_task.
STP_HD_NOINLINE void lb_dcsrch(LbfgsbState& S, double f, double g) {
  const double ftol = 1e-3, gtol = 0.9, xtol = 0.1, stpmin = 0.0, stpmax = S.stpmx;
  const double xtrapl = 1.1, xtrapu = 4.0;
  if (S.ls_task == 0) {
    if (S.stp < stpmin || S.stp > stpmax || g >= 0.0) { S.ls_task = 4; return; }
    S.brackt = 0; S.stage = 1; S.finit = f; S.ginit = g; S.gtest = ftol * g;
    S.width = stpmax - stpmin; S.width1 = S.width / 0.5;
    S.stx = 0.0; S.fx = f; S.gx = g; S.sty = 0.0; S.fy = f; S.gy = g;
    S.stmin = 0.0; S.stmax = S.stp + xtrapu * S.stp;
    S.ls_task = 1;
    return;
  }
  const double ftest = S.finit + S.stp * S.gtest;
  if (S.stage == 1 && f <= ftest && g >= 0.0) S.stage = 2;
  int task = 1;
  if (S.brackt && (S.stp <= S.stmin || S.stp >= S.stmax)) task = 3;
  if (S.brackt && S.stmax - S.stmin <= xtol * S.stmax) task = 3;
  if (S.stp == stpmax && f <= ftest && g <= S.gtest) task = 3;
  if (S.stp == stpmin && (f > ftest || g >= S.gtest)) task = 3;
  if (f <= ftest && fabs(g) <= gtol * (-S.ginit)) task = 2;
  if (task != 1) { S.ls_task = task; return; }
  if (S.stage == 1 && f <= S.fx && f > ftest) {
    const double fm = f - S.stp * S.gtest;
    double fxm = S.fx - S.stx * S.gtest, fym = S.fy - S.sty * S.gtest;
    const double gm = g - S.gtest;
    double gxm = S.gx - S.gtest, gym = S.gy - S.gtest;
    lb_dcstep(S.stx, fxm, gxm, S.sty, fym, gym, S.stp, fm, gm, S.brackt, S.stmin, S.stmax);
    S.fx = fxm + S.stx * S.gtest; S.fy = fym + S.sty * S.gtest;
    S.gx = gxm + S.gtest; S.gy = gym + S.gtest;
  } else {
    lb_dcstep(S.stx, S.fx, S.gx, S.sty, S.fy, S.gy, S.stp, f, g, S.brackt, S.stmin, S.stmax);
  }
  if (S.brackt) {
    if (fabs(S.sty - S.stx) >= 0.66 * S.width1) S.stp = S.stx + 0.5 * (S.sty - S.stx);
    S.width1 = S.width;
    S.width = fabs(S.sty - S.stx);
  }
  if (S.brackt) { S.stmin = fmin(S.stx, S.sty); S.stmax = fmax(S.stx, S.sty); }
  else { S.stmin = S.stp + xtrapl * (S.stp - S.stx); S.stmax = S.stp + xtrapu * (S.stp - S.stx); }
  S.stp = fmax(S.stp, stpmin);
  S.stp = fmin(S.stp, stpmax);
  if ((S.brackt && (S.stp <= S.stmin || S.stp >= S.stmax)) ||
      (S.brackt && S.stmax - S.stmin <= xtol * S.stmax))
    S.stp = S.stx;
  S.ls_task = 1;
}

// ------------------------------ driver ---------------------------------------------------
template <class LN>
STP_HD void lbfgsb_init(const LN& L, LbfgsbState& S, int np, const double* x0, const double* lb, const double* ub,
                        const LbfgsbOpts& o) {
  int not_boxed = 0, bounded = 0;
  STP_OWNED(L, i, np) {
    const int hl = lb[i] > -1.79e308, hu = ub[i] < 1.79e308;
    S.lb[i] = hl ? lb[i] : 0.0; S.ub[i] = hu ? ub[i] : 0.0;
    S.nbd[i] = hl ? (hu ? 2 : 1) : (hu ? 3 : 0);
    double xi = x0[i];
    if (hl && xi < lb[i]) xi = lb[i];   // scipy clips x0 into the box; `active` projects again
    if (hu && xi > ub[i]) xi = ub[i];
    S.x[i] = xi;
    if (S.nbd[i] != 2) not_boxed = 1;
    if (S.nbd[i] != 0) bounded = 1;
    S.iwhere[i] = (S.nbd[i] == 0) ? -1 : ((S.nbd[i] == 2 && S.ub[i] - S.lb[i] <= 0.0) ? 3 : 0);
  }
  not_boxed = L.any(not_boxed);
  bounded = L.any(bounded);
  if (L.lane() == 0) {
    S.np = np; S.m = o.m < kHist ? o.m : kHist; S.maxiter = o.maxiter; S.maxfun = o.maxfun; S.maxls = o.maxls;
    S.factr = o.factr; S.pgtol = o.pgtol;
    S.cnstnd = bounded; S.boxed = !not_boxed;
    S.iter = 0; S.nfev = 0; S.nskip = 0; S.iback = 0; S.ifun = 0; S.info = 0;
    S.phase = 0; S.status = -1; S.f = 0.0; S.fold = 0.0; S.sbgnrm = 0.0; S.stp = 0.0; S.lsfail = 0;
  }
  lb_reset_memory(L, S);
}

// Start (or restart) an iteration at the current x: Cauchy point, subspace step, direction,
// first trial point of the line search.  Returns 1 when S.x holds a point to evaluate,
// 0 when the run ended (S.status set).
template <class LN>
STP_HD_NOINLINE int lb_begin_iteration(const LN& L, LbfgsbState& S) {
  const int n = S.np, l0 = L.lane();
  STP_ROLLED for (int guard = 0; guard < 4; ++guard) {
    STP_PHASE(15);
    if (S.col > 0) {
      if (lb_build_B(L, S)) { lb_reset_memory(L, S); continue; }   // breakdown while forming B
    } else {
      const double th = S.theta;
      STP_OWNED(L, i, n) {
        STP_ROLLED for (int j = 0; j < n; ++j) S.B[i][j] = (i == j) ? th : 0.0;
      }
      L.sync();
    }
    STP_PHASE(11);
    if (!S.cnstnd && S.col > 0) {
      STP_OWNED(L, i, n) S.z[i] = S.x[i];
      L.sync();
    } else {
      lb_cauchy(L, S);
    }
    STP_PHASE(12);
    if (S.col > 0 && lb_subsm(L, S)) { lb_reset_memory(L, S); continue; }
    STP_PHASE(13);
    // search direction and the line-search set-up (lnsrlb, first entry)
    STP_OWNED(L, i, n) { S.d[i] = S.z[i] - S.x[i]; S.t[i] = S.x[i]; S.r[i] = S.g[i]; }
    L.sync();
    const double dtd = lb_dot(L, S.d, S.d, n);
    const double gd = lb_dot(L, S.g, S.d, n);
    if (l0 == 0) {
      S.dtd = dtd;
      S.dnorm = lb_sqrt(dtd);
      double stpmx = 1e10;
      if (S.cnstnd) {
        if (S.iter == 0) stpmx = 1.0;
        else {
          STP_ROLLED for (int i = 0; i < n; ++i) {
            const double a1 = S.d[i];
            const int nbd = S.nbd[i];
            if (nbd != 0) {
              if (a1 < 0.0 && nbd <= 2) {
                const double a2 = S.lb[i] - S.x[i];
                if (a2 >= 0.0) stpmx = 0.0; else if (a1 * stpmx < a2) stpmx = lb_div(a2, a1);
              } else if (a1 > 0.0 && nbd >= 2) {
                const double a2 = S.ub[i] - S.x[i];
                if (a2 <= 0.0) stpmx = 0.0; else if (a1 * stpmx > a2) stpmx = lb_div(a2, a1);
              }
            }
          }
        }
      }
      S.stpmx = stpmx;
      S.stp = (S.iter == 0 && !S.boxed) ? fmin(lb_div(1.0, S.dnorm), stpmx) : 1.0;
      S.fold = S.f;
      S.ifun = 0; S.iback = 0; S.ls_task = 0;
      S.gd = gd;
      S.gdold = gd;
      int lsfail = 0;
      if (gd >= 0.0) lsfail = 1;               // not a descent direction: info = -4
      if (!lsfail) {
        lb_dcsrch(S, S.f, gd);
        if (S.ls_task == 4) lsfail = 1;
      }
      S.lsfail = lsfail;
    }
    L.sync();
    if (S.lsfail) {
      if (S.col == 0) {   // abnormal termination in lnsrch
        L.sync();
        if (l0 == 0) { S.status = 2; S.iter += 1; }
        L.sync();
        return 0;
      }
      lb_reset_memory(L, S);
      continue;
    }
    const double stp = S.stp;
    STP_OWNED(L, i, n) S.x[i] = (stp == 1.0) ? S.z[i] : stp * S.d[i] + S.t[i];
    if (l0 == 0) { S.ifun = 1; S.nfev += 1; S.iback = 0; S.phase = 1; }
    L.sync();
    STP_PHASE(14);
    return 1;
  }
  L.sync();
  if (l0 == 0) S.status = 2;
  L.sync();
  return 0;
}

// Give up the current line search: back to the iterate it started from.
template <class LN>
STP_HD void lb_restore_iterate(const LN& L, LbfgsbState& S) {
  STP_OWNED(L, i, S.np) { S.x[i] = S.t[i]; S.g[i] = S.r[i]; }
  if (L.lane() == 0) S.f = S.fold;
  L.sync();
}

template <class LN>
STP_HD int lb_finish(const LN& L, LbfgsbState& S, int status, int bump_iter) {
  L.sync();
  if (L.lane() == 0) { S.status = status; if (bump_iter) S.iter += 1; }
  L.sync();
  return 0;
}

// One transition of the state machine.  Returns 0 finished, 1 evaluate S.x, 2 S.x is the point
// (f, g) already belong to (the caller feeds them again without evaluating).
template <class LN>
STP_HD_NOINLINE int lb_advance_once(const LN& L, LbfgsbState& S, double f, const double* g) {
  const int n = S.np, l0 = L.lane();
  STP_PHASE(19);
  int badg = 0;
  STP_OWNED(L, i, n) badg |= !(g[i] == g[i]);
  const int finite = (f == f) && (fabs(f) <= 1.79e308) && !L.any(badg);
  const int phase = S.phase;
  L.sync();
  STP_OWNED(L, i, n) S.g[i] = g[i];
  if (l0 == 0) { S.f = f; if (phase == 0) S.nfev = 1; }
  L.sync();
  if (phase == 0) {
    if (!finite) return lb_finish(L, S, 3, 0);
    lb_projgr(L, S);
    if (S.sbgnrm <= S.pgtol) return lb_finish(L, S, 0, 0);
    return lb_begin_iteration(L, S);
  }
  // inside a line search: (f, g) belong to x = t + stp d
  if (!finite) {
    // A non-finite objective cannot be handled by the cubic interpolation: the fit ends with status 3 and the caller
    // freezes the campaign.  Upstream differs here ("Fit failures"): botorch's closure turns a failed
    // evaluation into a NaN loss that scipy's dcsrch keeps stepping on, and _fit_fallback then retries the whole fit
    // from resampled hyper-parameters (up to 5 attempts) before run_single_loop's except-branch returns the partial
    // trace.  With psd_safe_cholesky's jitter retries inside the closure this branch was not reached in any of the 250
    // golden ExactGP campaigns nor in the bench workloads (frozen_campaigns = 0).
    lb_restore_iterate(L, S);
    return lb_finish(L, S, 3, 0);
  }
  const double gd = lb_dot(L, S.g, S.d, n);
  STP_PHASE(20);
  if (l0 == 0) { S.gd = gd; lb_dcsrch(S, f, gd); }
  L.sync();
  STP_PHASE(21);
  const int task = S.ls_task;
  if (task == 1) {
    if (S.iback + 1 >= S.maxls) {
      // too many trial points: restore and restart from scratch (or stop if memory is empty)
      lb_restore_iterate(L, S);
      if (S.col == 0) return lb_finish(L, S, 2, 1);
      lb_reset_memory(L, S);
      return lb_begin_iteration(L, S);
    }
    const double stp = S.stp;
    int mv = 0;
    STP_OWNED(L, i, n) {
      const double xi = (stp == 1.0) ? S.z[i] : stp * S.d[i] + S.t[i];
      mv |= (xi != S.x[i]);
      S.x[i] = xi;
    }
    const int moved = L.any(mv);
    L.sync();
    if (l0 == 0) { S.ifun += 1; S.iback = S.ifun - 1; if (moved) S.nfev += 1; }
    L.sync();
    return moved ? 1 : 2;   // 2: same point again (stp fell back to stx): reuse f, g like scipy's cache
  }
  if (task == 4) {
    lb_restore_iterate(L, S);
    if (S.col == 0) return lb_finish(L, S, 2, 1);
    lb_reset_memory(L, S);
    return lb_begin_iteration(L, S);
  }
  // line search finished (CONVERGENCE or WARNING): new iterate
  L.sync();
  if (l0 == 0) S.iter += 1;
  lb_projgr(L, S);
  // scipy's wrapper sees NEW_X here
  if (S.iter >= S.maxiter) return lb_finish(L, S, 1, 0);
  if (S.nfev > S.maxfun) return lb_finish(L, S, 1, 0);
  if (S.sbgnrm <= S.pgtol) return lb_finish(L, S, 0, 0);
  const double ddum = fmax(fmax(fabs(S.fold), fabs(S.f)), 1.0);
  if (S.fold - S.f <= kEpsMch * S.factr * ddum) return lb_finish(L, S, 0, 0);
  STP_PHASE(22);
  // BFGS pair
  const double stp = S.stp, gdn = S.gd, gdold = S.gdold;
  double rrp = 0.0, dr, dd;
  STP_OWNED(L, i, n) {
    const double ri = S.g[i] - S.r[i];
    S.r[i] = ri;
    rrp += ri * ri;
    if (stp != 1.0) S.d[i] *= stp;
  }
  const double rr = L.sum(rrp);
  L.sync();
  const double ysn = lb_dot(L, S.d, S.r, n);   // s.y of the new pair
  if (stp == 1.0) { dr = gdn - gdold; dd = -gdold; }
  else { dr = (gdn - gdold) * stp; dd = -gdold * stp; }
  int iupdat = S.iupdat, col = S.col, itail = S.itail, head = S.head;
  const int m = S.m;
  L.sync();
  if (dr <= kEpsMch * dd) {
    if (l0 == 0) { S.nskip += 1; S.updatd = 0; }
  } else {
    iupdat += 1;
    if (iupdat <= m) { col = iupdat; itail = (head + iupdat - 1) % m; }
    else { itail = (itail + 1) % m; head = (head + 1) % m; }
    STP_OWNED(L, i, n) { S.ws[itail][i] = S.d[i]; S.wy[itail][i] = S.r[i]; }
    if (l0 == 0) { S.updatd = 1; S.iupdat = iupdat; S.col = col; S.itail = itail; S.head = head; S.theta = lb_div(rr, dr);
                   S.ys[itail] = ysn; S.iys[itail] = lb_div(1.0, ysn); }
  }
  L.sync();
  return lb_begin_iteration(L, S);
}

// Feed the objective at S.x.  Returns 1 if S.x now holds a new point to evaluate, else 0 (finished).
// scipy's ScalarFunction does not re-evaluate (or count) a point equal to the previous one;
// neither do we.
template <class LN>
STP_HD int lbfgsb_advance(const LN& L, LbfgsbState& S, double f, const double* g) {
  for (;;) {
    const int r = lb_advance_once(L, S, f, g);
    if (r != 2) return r;
  }
}

// single-thread convenience forms (host emulation, unit tests)
STP_HD void lbfgsb_init(LbfgsbState& S, int np, const double* x0, const double* lb, const double* ub,
                        const LbfgsbOpts& o) {
  lbfgsb_init(SoloLanes(), S, np, x0, lb, ub, o);
}
STP_HD int lbfgsb_advance(LbfgsbState& S, double f, const double* g) { return lbfgsb_advance(SoloLanes(), S, f, g); }

}  // namespace stp
