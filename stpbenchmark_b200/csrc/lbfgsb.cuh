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
// Ask/tell interface, serial code (run by ONE thread of the CTA; the objective between two
// calls is evaluated by the whole CTA):
//     lbfgsb_init(S, ...);                       // S.x holds the first point to evaluate
//     while (lbfgsb_advance(S, f, g)) { f, g = objective(S.x); }
//     -> S.x, S.f, S.status (0 converged, 1 iteration/evaluation cap, 2 abnormal line search,
//        3 non-finite objective), S.iter, S.nfev
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
  int B_ready;                    // the dense B of the current memory has been formed (possibly by the team)
  // scratch of the serial routines: kept in the (shared-memory) state on purpose -- with ~220 KB of the SM's
  // 228 KB carved out as shared memory there is next to no L1 left, and thread-local arrays would turn every
  // access of the optimiser thread into an L2 round trip
  double sc_a[kMaxP], sc_b[kMaxP], sc_c[kMaxP], sc_d[kMaxP], sc_C[kMaxP][kMaxP];
  int sc_i[kMaxP], sc_j[kMaxP];
  int status;
};

// ------------------------------ small dense helpers -----------------------------------
STP_HD double lb_dot(const double* a, const double* b, int n) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

STP_HD void lb_projgr(LbfgsbState& S) {
  double m = 0.0;
  for (int i = 0; i < S.np; ++i) {
    double gi = S.g[i];
    if (S.nbd[i] != 0) {
      if (gi < 0.0) { if (S.nbd[i] >= 2) gi = fmax(S.x[i] - S.ub[i], gi); }
      else { if (S.nbd[i] <= 2) gi = fmin(S.x[i] - S.lb[i], gi); }
    }
    m = fmax(m, fabs(gi));
  }
  S.sbgnrm = m;
}

STP_HD void lb_reset_memory(LbfgsbState& S) {
  S.col = 0; S.head = 0; S.itail = -1; S.theta = 1.0; S.iupdat = 0; S.updatd = 0;
}

// B = theta I replayed through the stored (s, y) pairs, oldest first.  Returns 1 on breakdown.
STP_HD int lb_build_B(LbfgsbState& S) {
  const int n = S.np;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) S.B[i][j] = (i == j) ? S.theta : 0.0;
  for (int q = 0; q < S.col; ++q) {
    const int p = (S.head + q) % S.m;
    const double* s = S.ws[p];
    const double* y = S.wy[p];
    double* Bs = S.sc_a;
    double sBs = 0.0, ys = 0.0;
    for (int i = 0; i < n; ++i) {
      double a = 0.0;
      for (int j = 0; j < n; ++j) a += S.B[i][j] * s[j];
      Bs[i] = a;
    }
    for (int i = 0; i < n; ++i) { sBs += s[i] * Bs[i]; ys += s[i] * y[i]; }
    if (!(sBs > 0.0) || !(ys > 0.0)) return 1;
    const double ia = 1.0 / sBs, ib = 1.0 / ys;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) S.B[i][j] += y[i] * y[j] * ib - Bs[i] * Bs[j] * ia;
  }
  return 0;
}

// ------------------------------ generalised Cauchy point -------------------------------
// On exit S.z = x^cp and S.iwhere marks the variables fixed at a bound.
STP_HD void lb_cauchy(LbfgsbState& S) {
  const int n = S.np;
  double* xcp = S.z;
  double* dvec = S.sc_a; double* tbrk = S.sc_b; double* zz = S.sc_c;
  int* order = S.sc_i;
  int nbreak = 0;
  if (S.sbgnrm <= 0.0) { for (int i = 0; i < n; ++i) xcp[i] = S.x[i]; return; }
  for (int i = 0; i < n; ++i) {
    const double neggi = -S.g[i];
    double tl = 0.0, tu = 0.0;
    if (S.iwhere[i] != 3 && S.iwhere[i] != -1) {
      if (S.nbd[i] <= 2) tl = S.x[i] - S.lb[i];
      if (S.nbd[i] >= 2) tu = S.ub[i] - S.x[i];
      const int xlower = S.nbd[i] <= 2 && tl <= 0.0;
      const int xupper = S.nbd[i] >= 2 && tu <= 0.0;
      S.iwhere[i] = 0;
      if (xlower) { if (neggi <= 0.0) S.iwhere[i] = 1; }
      else if (xupper) { if (neggi >= 0.0) S.iwhere[i] = 2; }
      else { if (fabs(neggi) <= 0.0) S.iwhere[i] = -3; }
    }
    xcp[i] = S.x[i];
    zz[i] = 0.0;
    if (S.iwhere[i] != 0 && S.iwhere[i] != -1) {
      dvec[i] = 0.0;
    } else {
      dvec[i] = neggi;
      if (S.nbd[i] <= 2 && S.nbd[i] != 0 && neggi < 0.0) { tbrk[nbreak] = tl / (-neggi); order[nbreak++] = i; }
      else if (S.nbd[i] >= 2 && neggi > 0.0) { tbrk[nbreak] = tu / neggi; order[nbreak++] = i; }
    }
  }
  // sort breakpoints ascending (insertion sort on <= 11 entries)
  for (int a = 1; a < nbreak; ++a) {
    const double tv = tbrk[a]; const int ov = order[a];
    int b = a - 1;
    while (b >= 0 && tbrk[b] > tv) { tbrk[b + 1] = tbrk[b]; order[b + 1] = order[b]; --b; }
    tbrk[b + 1] = tv; order[b + 1] = ov;
  }
  // derivatives of the quadratic model along the projected-gradient path
  double f1 = 0.0, f2 = 0.0;
  for (int i = 0; i < n; ++i) {
    double bd = 0.0;
    for (int j = 0; j < n; ++j) bd += S.B[i][j] * dvec[j];
    f1 -= dvec[i] * dvec[i];
    f2 += dvec[i] * bd;
  }
  const double f2_org = -S.theta * f1;  // the col = 0 value, as in the published routine
  if (!(f2 > 0.0)) return;               // d = 0: x is its own Cauchy point
  double dtm = -f1 / f2;
  double tsum = 0.0, tj = 0.0;
  int nfixed_all = 0;
  for (int q = 0; q < nbreak; ++q) {
    const double tj0 = tj;
    tj = tbrk[q];
    const double dt = tj - tj0;
    if (dtm < dt) break;
    const int ibp = order[q];
    tsum += dt;
    const double dibp = dvec[ibp];
    dvec[ibp] = 0.0;
    if (dibp > 0.0) { xcp[ibp] = S.ub[ibp]; S.iwhere[ibp] = 2; }
    else { xcp[ibp] = S.lb[ibp]; S.iwhere[ibp] = 1; }
    if (q == nbreak - 1 && nbreak == n) { dtm = dt; nfixed_all = 1; break; }
    // z = x(tsum) - x : fixed variables sit on their bounds, the others moved tsum * d
    for (int i = 0; i < n; ++i) zz[i] = (dvec[i] != 0.0 || S.iwhere[i] <= 0) ? tsum * dvec[i] : xcp[i] - S.x[i];
    f1 = 0.0; f2 = 0.0;
    for (int i = 0; i < n; ++i) {
      double bd = 0.0;
      for (int j = 0; j < n; ++j) bd += S.B[i][j] * dvec[j];
      f1 += dvec[i] * S.g[i];
      f2 += dvec[i] * bd;
      f1 += zz[i] * bd;
    }
    f2 = fmax(kEpsMch * f2_org, f2);
    dtm = -f1 / f2;
  }
  if (!nfixed_all) {
    if (dtm < 0.0) dtm = 0.0;
    tsum += dtm;
  }
  for (int i = 0; i < n; ++i)
    if (dvec[i] != 0.0) xcp[i] = S.x[i] + tsum * dvec[i];
}

// ------------------------------ subspace minimisation ----------------------------------
// Minimise the quadratic model over the variables free at x^cp, then project (v3.0).
// On entry S.z = x^cp; on exit S.z = the subspace minimiser.  Returns 1 if B_FF is not PD.
STP_HD int lb_subsm(LbfgsbState& S) {
  const int n = S.np;
  int* ind = S.sc_j;
  int nf = 0;
  for (int i = 0; i < n; ++i) if (S.iwhere[i] <= 0) ind[nf++] = i;
  if (nf == 0 || S.col == 0) return 0;
  double* rr = S.sc_a; double* dd = S.sc_b;
  double (*C)[kMaxP] = S.sc_C;
  for (int a = 0; a < nf; ++a) {
    const int i = ind[a];
    double bz = 0.0;
    for (int j = 0; j < n; ++j) bz += S.B[i][j] * (S.z[j] - S.x[j]);
    rr[a] = -(S.g[i] + bz);
    for (int b = 0; b <= a; ++b) C[a][b] = S.B[i][ind[b]];
  }
  double* idiag = S.sc_c;           // reciprocals of the Cholesky diagonal: one division per column, not per entry
  for (int j = 0; j < nf; ++j) {  // Cholesky of B_FF
    double s = C[j][j];
    for (int k = 0; k < j; ++k) s -= C[j][k] * C[j][k];
    if (!(s > 0.0)) return 1;
    const double l = sqrt(s);
    const double il = 1.0 / l;
    C[j][j] = l;
    idiag[j] = il;
    for (int i = j + 1; i < nf; ++i) {
      double v = C[i][j];
      for (int k = 0; k < j; ++k) v -= C[i][k] * C[j][k];
      C[i][j] = v * il;
    }
  }
  for (int i = 0; i < nf; ++i) {
    double v = rr[i];
    for (int k = 0; k < i; ++k) v -= C[i][k] * dd[k];
    dd[i] = v * idiag[i];
  }
  for (int i = nf - 1; i >= 0; --i) {
    double v = dd[i];
    for (int k = i + 1; k < nf; ++k) v -= C[k][i] * dd[k];
    dd[i] = v * idiag[i];
  }
  // projection of x^cp + d onto the box
  int iword = 0;
  for (int i = 0; i < n; ++i) S.xp[i] = S.z[i];
  for (int a = 0; a < nf; ++a) {
    const int k = ind[a];
    const double xk = S.z[k], dk = dd[a];
    if (S.nbd[k] != 0) {
      if (S.nbd[k] == 1) { S.z[k] = fmax(S.lb[k], xk + dk); if (S.z[k] == S.lb[k]) iword = 1; }
      else if (S.nbd[k] == 2) {
        const double v = fmax(S.lb[k], xk + dk);
        S.z[k] = fmin(S.ub[k], v);
        if (S.z[k] == S.lb[k] || S.z[k] == S.ub[k]) iword = 1;
      } else { S.z[k] = fmin(S.ub[k], xk + dk); if (S.z[k] == S.ub[k]) iword = 1; }
    } else {
      S.z[k] = xk + dk;
    }
  }
  if (!iword) return 0;
  double ddp = 0.0;
  for (int i = 0; i < n; ++i) ddp += (S.z[i] - S.x[i]) * S.g[i];
  if (ddp > 0.0) {
    for (int i = 0; i < n; ++i) S.z[i] = S.xp[i];
    double alpha = 1.0, temp1 = 1.0;
    int ibd = -1;
    for (int a = 0; a < nf; ++a) {
      const int k = ind[a];
      const double dk = dd[a];
      if (S.nbd[k] != 0) {
        if (dk < 0.0 && S.nbd[k] <= 2) {
          const double t2 = S.lb[k] - S.z[k];
          if (t2 >= 0.0) temp1 = 0.0; else if (dk * alpha < t2) temp1 = t2 / dk;
        } else if (dk > 0.0 && S.nbd[k] >= 2) {
          const double t2 = S.ub[k] - S.z[k];
          if (t2 <= 0.0) temp1 = 0.0; else if (dk * alpha > t2) temp1 = t2 / dk;
        }
        if (temp1 < alpha) { alpha = temp1; ibd = a; }
      }
    }
    if (alpha < 1.0 && ibd >= 0) {
      const double dk = dd[ibd];
      const int k = ind[ibd];
      if (dk > 0.0) { S.z[k] = S.ub[k]; dd[ibd] = 0.0; }
      else if (dk < 0.0) { S.z[k] = S.lb[k]; dd[ibd] = 0.0; }
    }
    for (int a = 0; a < nf; ++a) S.z[ind[a]] += alpha * dd[a];
  }
  return 0;
}

// ------------------------------ More-Thuente line search -------------------------------
STP_HD void lb_dcstep(double& stx, double& fx, double& dx, double& sty, double& fy, double& dy, double& stp,
                      double fp, double dp, int& brackt, double stpmin, double stpmax) {
  const double sgnd = dp * (dx / fabs(dx));
  double stpf, stpc, stpq, th, s, gamma, p, q, r;
  if (fp > fx) {
    th = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    s = fmax(fabs(th), fmax(fabs(dx), fabs(dp)));
    gamma = s * sqrt((th / s) * (th / s) - (dx / s) * (dp / s));
    if (stp < stx) gamma = -gamma;
    p = (gamma - dx) + th; q = ((gamma - dx) + gamma) + dp; r = p / q;
    stpc = stx + r * (stp - stx);
    stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
    stpf = (fabs(stpc - stx) < fabs(stpq - stx)) ? stpc : stpc + (stpq - stpc) / 2.0;
    brackt = 1;
  } else if (sgnd < 0.0) {
    th = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    s = fmax(fabs(th), fmax(fabs(dx), fabs(dp)));
    gamma = s * sqrt((th / s) * (th / s) - (dx / s) * (dp / s));
    if (stp > stx) gamma = -gamma;
    p = (gamma - dp) + th; q = ((gamma - dp) + gamma) + dx; r = p / q;
    stpc = stp + r * (stx - stp);
    stpq = stp + (dp / (dp - dx)) * (stx - stp);
    stpf = (fabs(stpc - stp) > fabs(stpq - stp)) ? stpc : stpq;
    brackt = 1;
  } else if (fabs(dp) < fabs(dx)) {
    th = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    s = fmax(fabs(th), fmax(fabs(dx), fabs(dp)));
    gamma = s * sqrt(fmax(0.0, (th / s) * (th / s) - (dx / s) * (dp / s)));
    if (stp > stx) gamma = -gamma;
    p = (gamma - dp) + th; q = (gamma + (dx - dp)) + gamma; r = p / q;
    if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
    else if (stp > stx) stpc = stpmax;
    else stpc = stpmin;
    stpq = stp + (dp / (dp - dx)) * (stx - stp);
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
      th = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
      s = fmax(fabs(th), fmax(fabs(dy), fabs(dp)));
      gamma = s * sqrt((th / s) * (th / s) - (dy / s) * (dp / s));
      if (stp > sty) gamma = -gamma;
      p = (gamma - dp) + th; q = ((gamma - dp) + gamma) + dy; r = p / q;
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
STP_HD void lb_dcsrch(LbfgsbState& S, double f, double g) {
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
STP_HD void lbfgsb_init(LbfgsbState& S, int np, const double* x0, const double* lb, const double* ub,
                        const LbfgsbOpts& o) {
  S.np = np; S.m = o.m < kHist ? o.m : kHist; S.maxiter = o.maxiter; S.maxfun = o.maxfun; S.maxls = o.maxls;
  S.factr = o.factr; S.pgtol = o.pgtol;
  S.cnstnd = 0; S.boxed = 1;
  for (int i = 0; i < np; ++i) {
    const int hl = lb[i] > -1.79e308, hu = ub[i] < 1.79e308;
    S.lb[i] = hl ? lb[i] : 0.0; S.ub[i] = hu ? ub[i] : 0.0;
    S.nbd[i] = hl ? (hu ? 2 : 1) : (hu ? 3 : 0);
    double xi = x0[i];
    if (hl && xi < lb[i]) xi = lb[i];   // scipy clips x0 into the box; `active` projects again
    if (hu && xi > ub[i]) xi = ub[i];
    S.x[i] = xi;
    if (S.nbd[i] != 2) S.boxed = 0;
    if (S.nbd[i] != 0) S.cnstnd = 1;
    S.iwhere[i] = (S.nbd[i] == 0) ? -1 : ((S.nbd[i] == 2 && S.ub[i] - S.lb[i] <= 0.0) ? 3 : 0);
  }
  lb_reset_memory(S);
  S.iter = 0; S.nfev = 0; S.nskip = 0; S.iback = 0; S.ifun = 0; S.info = 0;
  S.phase = 0; S.status = -1; S.f = 0.0; S.fold = 0.0; S.sbgnrm = 0.0; S.stp = 0.0; S.B_ready = 0;
}

// Start (or restart) an iteration at the current x: Cauchy point, subspace step, direction,
// first trial point of the line search.  Returns 1 when S.x holds a point to evaluate,
// 0 when the run ended (S.status set).
STP_HD int lb_begin_iteration(LbfgsbState& S) {
  const int n = S.np;
  for (int guard = 0; guard < 4; ++guard) {
    if (S.col > 0 && !S.B_ready) return 3;      // caller forms B (lb_build_B or its team version), then re-enters
    if (S.col > 0 && S.B_ready == 2) { lb_reset_memory(S); S.B_ready = 0; continue; }   // breakdown while forming B
    S.B_ready = 0;
    if (S.col == 0)
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) S.B[i][j] = (i == j) ? S.theta : 0.0;
    if (!S.cnstnd && S.col > 0) {
      for (int i = 0; i < n; ++i) S.z[i] = S.x[i];
    } else {
      lb_cauchy(S);
    }
    if (S.col > 0 && lb_subsm(S)) { lb_reset_memory(S); continue; }
    // search direction and the line-search set-up (lnsrlb, first entry)
    for (int i = 0; i < n; ++i) S.d[i] = S.z[i] - S.x[i];
    S.dtd = lb_dot(S.d, S.d, n);
    S.dnorm = sqrt(S.dtd);
    S.stpmx = 1e10;
    if (S.cnstnd) {
      if (S.iter == 0) S.stpmx = 1.0;
      else {
        for (int i = 0; i < n; ++i) {
          const double a1 = S.d[i];
          if (S.nbd[i] != 0) {
            if (a1 < 0.0 && S.nbd[i] <= 2) {
              const double a2 = S.lb[i] - S.x[i];
              if (a2 >= 0.0) S.stpmx = 0.0; else if (a1 * S.stpmx < a2) S.stpmx = a2 / a1;
            } else if (a1 > 0.0 && S.nbd[i] >= 2) {
              const double a2 = S.ub[i] - S.x[i];
              if (a2 <= 0.0) S.stpmx = 0.0; else if (a1 * S.stpmx > a2) S.stpmx = a2 / a1;
            }
          }
        }
      }
    }
    S.stp = (S.iter == 0 && !S.boxed) ? fmin(1.0 / S.dnorm, S.stpmx) : 1.0;
    for (int i = 0; i < n; ++i) { S.t[i] = S.x[i]; S.r[i] = S.g[i]; }
    S.fold = S.f;
    S.ifun = 0; S.iback = 0; S.ls_task = 0;
    S.gd = lb_dot(S.g, S.d, n);
    S.gdold = S.gd;
    int lsfail = 0;
    if (S.gd >= 0.0) lsfail = 1;               // not a descent direction: info = -4
    if (!lsfail) {
      lb_dcsrch(S, S.f, S.gd);
      if (S.ls_task == 4) lsfail = 1;
    }
    if (lsfail) {
      if (S.col == 0) { S.status = 2; S.iter += 1; return 0; }  // abnormal termination in lnsrch
      lb_reset_memory(S);
      continue;
    }
    S.ifun = 1; S.nfev += 1; S.iback = 0;
    if (S.stp == 1.0) for (int i = 0; i < n; ++i) S.x[i] = S.z[i];
    else for (int i = 0; i < n; ++i) S.x[i] = S.stp * S.d[i] + S.t[i];
    S.phase = 1;
    return 1;
  }
  S.status = 2;
  return 0;
}

// One transition of the state machine.  Returns 0 finished, 1 evaluate S.x, 2 S.x is the point
// (f, g) already belong to (the caller feeds them again without evaluating).
STP_HD int lb_advance_once(LbfgsbState& S, double f, const double* g) {
  const int n = S.np;
  int finite = (f == f) && fabs(f) <= 1.79e308;
  for (int i = 0; i < n; ++i) finite &= (g[i] == g[i]);
  if (S.phase == 0) {
    S.nfev = 1;
    S.f = f;
    for (int i = 0; i < n; ++i) S.g[i] = g[i];
    if (!finite) { S.status = 3; return 0; }
    lb_projgr(S);
    if (S.sbgnrm <= S.pgtol) { S.status = 0; return 0; }
    return lb_begin_iteration(S);
  }
  // inside a line search: (f, g) belong to x = t + stp d
  S.f = f;
  for (int i = 0; i < n; ++i) S.g[i] = g[i];
  if (!finite) {
    // a non-finite objective cannot be handled by the cubic interpolation: give up on this
    // campaign (the reference's closure would raise and the trace would be NaN-padded)
    for (int i = 0; i < n; ++i) { S.x[i] = S.t[i]; S.g[i] = S.r[i]; }
    S.f = S.fold; S.status = 3; return 0;
  }
  S.gd = lb_dot(S.g, S.d, n);
  lb_dcsrch(S, f, S.gd);
  if (S.ls_task == 1) {
    if (S.iback + 1 >= S.maxls) {
      // too many trial points: restore and restart from scratch (or stop if memory is empty)
      for (int i = 0; i < n; ++i) { S.x[i] = S.t[i]; S.g[i] = S.r[i]; }
      S.f = S.fold;
      if (S.col == 0) { S.status = 2; S.iter += 1; return 0; }
      lb_reset_memory(S);
      return lb_begin_iteration(S);
    }
    S.ifun += 1; S.iback = S.ifun - 1;
    int moved = 0;
    for (int i = 0; i < n; ++i) {
      const double xi = (S.stp == 1.0) ? S.z[i] : S.stp * S.d[i] + S.t[i];
      moved |= (xi != S.x[i]);
      S.x[i] = xi;
    }
    if (!moved) return 2;   // same point again (stp fell back to stx): reuse f, g like scipy's cache
    S.nfev += 1;
    return 1;
  }
  if (S.ls_task == 4) {
    for (int i = 0; i < n; ++i) { S.x[i] = S.t[i]; S.g[i] = S.r[i]; }
    S.f = S.fold;
    if (S.col == 0) { S.status = 2; S.iter += 1; return 0; }
    lb_reset_memory(S);
    return lb_begin_iteration(S);
  }
  // line search finished (CONVERGENCE or WARNING): new iterate
  S.iter += 1;
  lb_projgr(S);
  // scipy's wrapper sees NEW_X here
  if (S.iter >= S.maxiter) { S.status = 1; return 0; }
  if (S.nfev > S.maxfun) { S.status = 1; return 0; }
  if (S.sbgnrm <= S.pgtol) { S.status = 0; return 0; }
  const double ddum = fmax(fmax(fabs(S.fold), fabs(S.f)), 1.0);
  if (S.fold - S.f <= kEpsMch * S.factr * ddum) { S.status = 0; return 0; }
  // BFGS pair
  double rr = 0.0, dr, dd;
  for (int i = 0; i < n; ++i) { S.r[i] = S.g[i] - S.r[i]; rr += S.r[i] * S.r[i]; }
  if (S.stp == 1.0) { dr = S.gd - S.gdold; dd = -S.gdold; }
  else {
    dr = (S.gd - S.gdold) * S.stp;
    for (int i = 0; i < n; ++i) S.d[i] *= S.stp;
    dd = -S.gdold * S.stp;
  }
  if (dr <= kEpsMch * dd) {
    S.nskip += 1; S.updatd = 0;
  } else {
    S.updatd = 1; S.iupdat += 1;
    if (S.iupdat <= S.m) { S.col = S.iupdat; S.itail = (S.head + S.iupdat - 1) % S.m; }
    else { S.itail = (S.itail + 1) % S.m; S.head = (S.head + 1) % S.m; }
    for (int i = 0; i < n; ++i) { S.ws[S.itail][i] = S.d[i]; S.wy[S.itail][i] = S.r[i]; }
    S.theta = rr / dr;
  }
  return lb_begin_iteration(S);
}

// Feed the objective at S.x.  Returns 1 if S.x now holds a new point to evaluate, else 0.
// scipy's ScalarFunction does not re-evaluate (or count) a point equal to the previous one;
// neither do we.
// Return codes: 0 finished, 1 evaluate S.x, 3 form the dense B (S.B_ready = lb_build_B(S) ? 2 : 1, or the
// team version) and call lbfgsb_resume(S).
STP_HD int lbfgsb_advance_split(LbfgsbState& S, double f, const double* g) {
  double* gc = S.sc_d;
  for (int i = 0; i < S.np; ++i) gc[i] = g[i];
  for (;;) {
    const int r = lb_advance_once(S, f, gc);
    if (r != 2) return r;
  }
}
STP_HD int lbfgsb_resume(LbfgsbState& S) {
  for (;;) {
    const int r = lb_begin_iteration(S);
    if (r != 2) return r;
  }
}
STP_HD int lbfgsb_advance(LbfgsbState& S, double f, const double* g) {
  int r = lbfgsb_advance_split(S, f, g);
  while (r == 3) {
    S.B_ready = lb_build_B(S) ? 2 : 1;
    r = lbfgsb_resume(S);
  }
  return r;
}

// Team version of lb_build_B: one lane per row of B (np <= 11 rows), pairs replayed with warp shuffles.
STP_HD void lb_build_B_team(const SoloTeam&, LbfgsbState& S) { S.B_ready = lb_build_B(S) ? 2 : 1; }
#if defined(__CUDACC__)
__device__ __forceinline__ void lb_build_B_team(const CtaTeam& tm, LbfgsbState& S) {
  if (tm.warp() != 0) return;
  const int n = S.np, r = tm.lane();
  const bool live = r < n;
  double Brow[kMaxP];
#pragma unroll
  for (int c = 0; c < kMaxP; ++c) Brow[c] = (live && c == r) ? S.theta : 0.0;
  int broke = 0;
  for (int q = 0; q < S.col; ++q) {
    const int p = (S.head + q) % S.m;
    const double* s = S.ws[p];
    const double* y = S.wy[p];
    double Bs = 0.0;
#pragma unroll
    for (int c = 0; c < kMaxP; ++c)
      if (c < n) Bs += Brow[c] * s[c];
    const double sr = live ? s[r] : 0.0, yr = live ? y[r] : 0.0;
    const double sBs = tm.warp_sum(sr * Bs), ys = tm.warp_sum(sr * yr);
    if (!(sBs > 0.0) || !(ys > 0.0)) { broke = 1; break; }
    const double ia = 1.0 / sBs, ib = 1.0 / ys;
#pragma unroll
    for (int c = 0; c < kMaxP; ++c)
      if (c < n) {
        const double Bsc = __shfl_sync(0xffffffffu, Bs, c);
        Brow[c] += yr * y[c] * ib - Bs * Bsc * ia;
      }
  }
  if (live) {
#pragma unroll
    for (int c = 0; c < kMaxP; ++c)
      if (c < n) S.B[r][c] = Brow[c];
  }
  __syncwarp();
  if (r == 0) S.B_ready = broke ? 2 : 1;
}
#endif

}  // namespace stp
